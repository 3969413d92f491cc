#!/bin/bash
# Run on the GPU box (gpurun): plain bench, then the ncu launch list and one --set full capture of the
# fused kernel for the SAME command line (B200_PROFILING.md recipe).  Outputs under gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_full.log 2>&1
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:body_quat -s 4 -c 2 -f -o gpurun_out/${TAG}_bq_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_bench_full.log
tail -n 2 gpurun_out/${TAG}_ncu1.log; tail -n 2 gpurun_out/${TAG}_ncu2.log
