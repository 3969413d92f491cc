#!/bin/bash
# One single-GPU box session: GPU tests, bit-exactness of the slow-path-free division / square root, the N = 1 bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -I humanoid_real_time_retarget_b200/csrc -o /tmp/exact_ops_check tools/ubench/exact_ops_check.cu \
  && /tmp/exact_ops_check > gpurun_out/r2_exact_ops.json 2>&1; echo "exact rc=$?" >> gpurun_out/r2_exact_ops.json
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?" >> gpurun_out/r2_bench1.err
tail -25 gpurun_out/r2_gputests.log; cat gpurun_out/r2_exact_ops.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d.get('e2e_dof_only',{}).get('value'))
print(d['roofline']['frac'], d['fk_65536'], d['pos_path_2p18'])
print(d['latency_us']['resident_server']['back_to_back'], d['latency_us']['reference_class_call'])
PY
