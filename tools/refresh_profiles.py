"""Regenerate the committed profile artefacts from a `tools/profile_bench.sh <tag>` run merged into gpurun_out/:
profiles/<round>_body_quat_ncu_full_summary.csv, profiles/traffic.json, profiles/<round>_launches_bench.csv.

    python tools/refresh_profiles.py <tag> [round, default r02]
"""
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
rep = f"gpurun_out/{tag}_bq_full.ncu-rep"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, launches = rows[0], rows[1], rows[2:]
keep = {r[0] for r in csv.reader(open("profiles/r01_body_quat_ncu_full_summary.csv"))}
with open(f"profiles/{rnd}_body_quat_ncu_full_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
    for i, h in enumerate(hdr):
        if h in keep:
            w.writerow([h, units[i]] + [l[i] for l in launches])
col = {h: i for i, h in enumerate(hdr)}
g = lambda n: float(launches[0][col[n]].replace(",", ""))  # noqa: E731
rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
t = json.load(open("profiles/traffic.json"))
t["body_quat_kernel"].update({
    "dram_bytes_per_launch": (rd + wr) * 1e6, "dram_bytes_read": rd * 1e6, "dram_bytes_write": wr * 1e6,
    "warp_instructions_per_launch": g("smsp__inst_executed.sum"),
    "issue_slots_busy_pct_ncu": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "source": "ncu --set full --clock-control none, python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras "
              f"(profiles/{rnd}_body_quat_ncu_full_summary.csv, capture {tag})"})
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
shutil.copy(f"gpurun_out/{tag}_launches.csv", f"profiles/{rnd}_launches_bench.csv")
print(launches[0][col["Kernel Name"]], "time", g("gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]],
      "inst", g("smsp__inst_executed.sum"), "issue", g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
      "fma", g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), "regs", g("launch__registers_per_thread"),
      "dram MB", rd, wr)
for l in open(f"gpurun_out/{tag}_bench_full.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("bench:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "hbm", d["roofline"]["frac"], "issue",
              d["roofline"]["issue"]["frac"], "lat", d["latency_us"]["resident_server"]["back_to_back"],
              d["latency_us"]["reference_class_call"], "fk", d["fk_65536"]["ms"], d["fk_65536"]["jacobian_2_links_ms"],
              "pos", d["pos_path_2p18"])
