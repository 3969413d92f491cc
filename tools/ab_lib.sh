# A/B of library variants (variants/<name>.so, tools/build_variant.sh) on one microbench case filter; the in-tree library is restored
#   bash tools/ab_lib.sh <log tag> <microbench --cases> <--bq-only filter or ''> <variant> [<variant> ...]
set -u
mkdir -p gpurun_out
L=humanoid_real_time_retarget_b200/libhrt_b200.so
TAG=$1; CASES=$2; FILT=$3; shift 3
cp $L /tmp/intree.so
for v in "$@"; do cp variants/$v.so $L; echo "== $v" | tee -a gpurun_out/${TAG}.log
  python tools/microbench.py --cases $CASES --iters 20 ${FILT:+--bq-only $FILT} 2>&1 | grep -E "case|rror" | cut -c1-130 | tee -a gpurun_out/${TAG}.log; done
cp /tmp/intree.so $L
