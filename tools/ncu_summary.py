"""Condense an .ncu-rep into the CSV that is committed under profiles/: one column per captured launch, the metrics the
notes quote (time, instructions, issue / pipe utilisation, stall reasons per issue, DRAM bytes, launch geometry).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x_ncu_summary.csv"""
import csv
import subprocess
import sys

rep, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, launches = rows[0], rows[1], rows[2:]
want_prefix = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
               "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
               "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
               "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_",
               "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
               "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [l[hdr.index("Kernel Name")][:60] for l in launches])
    for i, h in enumerate(hdr):
        if any(h == p or (p.endswith("_") and h.startswith(p)) or (p.startswith("launch__occupancy") and h.startswith(p)) for p in want_prefix) \
                or h in ("dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"):
            w.writerow([h, units[i]] + [l[i] for l in launches])
print("wrote", dst, len(launches), "launches")
