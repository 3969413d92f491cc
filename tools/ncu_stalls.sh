#!/bin/bash
# usage: tools/ncu_stalls.sh <skip> <count> <command...>   -- warp-state metrics of the body_quat kernels of a command
S=$1; C=$2; shift 2
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for r in barrier long_scoreboard short_scoreboard lg_throttle mio_throttle no_instruction wait math_pipe_throttle dispatch_stall branch_resolving membar sleeping not_selected; do M=$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio; done
ncu --metrics $M --clock-control none -k regex:body_quat -s $S -c $C --csv "$@" 2>/dev/null | python3 -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>8]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi].replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')]=r[vi]
for k,v in d.items(): print(k, v)
"
