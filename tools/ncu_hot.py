"""Summarise an ncu report's source page: per CUDA source line, executed warp instructions and
stall samples.   python tools/ncu_hot.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find header row
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
print("columns:", [h for h in hdr[:8]])
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
cur = None
tot_inst = tot_samp = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    addr = r[col.get("Address", 0)] if "Address" in col else ""
    try:
        inst = int(float(r[col["Instructions Executed"]] or 0))
        samp = int(float(r[col["# Samples"]] or 0))
    except ValueError:
        continue
    key = src.strip()[:110]
    # in cuda,sass view, cuda lines carry aggregated numbers; sass lines follow. keep cuda-line rows only
    agg[key][0] += inst
    agg[key][1] += samp
    for s in stall_cols:
        try:
            agg[key][3][s] += int(float(r[col[s]] or 0))
        except ValueError:
            pass
    tot_inst += inst
    tot_samp += samp
print("total inst", tot_inst, "samples", tot_samp)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(v[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"{v[0]:>12d} inst {v[1]:>7d} samp  {k}   {[(a.replace('stall_',''), b) for a, b in st if b]}")
