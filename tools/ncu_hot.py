"""Summarise an ncu report's source page per CUDA source line: executed warp instructions, stall samples and the top
stall reasons (the report must have been captured with --import-source on and the kernel compiled with -lineinfo).
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [top_n] [file-substring:first-last ...]
Extra arguments sum a line range of one file, e.g.  hrt_retarget.cuh:330:470  (instructions / samples of that range)."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ranges = [a.split(":") for a in sys.argv[3:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, col = "?", None, {}
lines = {}                       # (file, line) -> [source text, inst, samples, {stall: n}]
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].rsplit("/", 1)[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue                 # SASS rows have an empty line number: the CUDA row above carries their sum
    try:
        inst = int(float(r[col["Instructions Executed"]] or 0))
        samp = int(float(r[col["# Samples"]] or 0))
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    e = lines.setdefault(key, [r[1].strip()[:100], 0, 0, defaultdict(int)])
    e[1] += inst
    e[2] += samp
    for h, i in col.items():
        if h.startswith("stall_"):
            try:
                e[3][h[6:]] += int(float(r[i] or 0))
            except ValueError:
                pass
tot_i = sum(e[1] for e in lines.values())
tot_s = sum(e[2] for e in lines.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
byfile = defaultdict(lambda: [0, 0])
for (f, _), e in lines.items():
    byfile[f][0] += e[1]
    byfile[f][1] += e[2]
for f, (i, s) in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print(f"  {f:28s} {i:>13d} inst ({100 * i / max(tot_i, 1):5.1f} %) {s:>8d} samples ({100 * s / max(tot_s, 1):5.1f} %)")
for f, a, b in ranges:
    i = sum(e[1] for (ff, l), e in lines.items() if f in ff and int(a) <= l <= int(b))
    s = sum(e[2] for (ff, l), e in lines.items() if f in ff and int(a) <= l <= int(b))
    print(f"range {f}:{a}-{b}: {i} inst ({100 * i / max(tot_i, 1):.2f} %), {s} samples ({100 * s / max(tot_s, 1):.2f} %)")
for (f, l), e in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    st = sorted(e[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"{e[1]:>12d} inst {e[2]:>7d} samp  {f}:{l}  {e[0]}   {[(a, b) for a, b in st if b]}")
