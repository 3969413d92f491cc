"""Static SASS instruction count of one kernel per source line / function range (nvdisasm -g -c output on stdin or a file).
    cuobjdump -xelf all libhrt_b200.so && nvdisasm -g -c hrt_api.sm_100a.cubin > all.sass
    python tools/sass_by_line.py all.sass '_ZN3hrt19pos_retarget_kernelILi0ELi16E' [top]
Counts are static (a loop body counts once); inlined-at chains are ignored: a line is charged where its code sits."""
import collections
import re
import sys

path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
in_k = False
cur = ("?", 0)
cnt = collections.Counter()
ops = collections.Counter()
total = 0
for ln in open(path, errors="replace"):
    if ln.startswith(".text."):
        in_k = key in ln
        continue
    if not in_k:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m:
        cnt[cur] += 1
        ops[m.group(2).split(".")[0]] += 1
        total += 1
print("total instructions", total)
print("top opcodes", ops.most_common(25))
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print("by file", byfile.most_common())
for (f, l), c in cnt.most_common(top):
    print(f"{c:6d}  {f}:{l}")
