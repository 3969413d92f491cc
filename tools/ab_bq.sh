# A/B of library variants on the headline kernel: GPU tests on the in-tree library, then the body_quat microbench per variant
set -u
mkdir -p gpurun_out
L=humanoid_real_time_retarget_b200/libhrt_b200.so
TAG=${1:-ab}; shift
cp $L /tmp/intree.so
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
cp gpurun_out/parity_r02.json gpurun_out/${TAG}_parity.json 2>/dev/null
for v in "$@"; do cp variants/$v.so $L; echo "== $v" | tee -a gpurun_out/${TAG}_bq.log
  for c in ik10 ik1_; do python tools/microbench.py --cases bq --iters 20 --bq-only $c 2>&1 | grep -E "case|rror" | cut -c1-160 | tee -a gpurun_out/${TAG}_bq.log; done; done
cp /tmp/intree.so $L
