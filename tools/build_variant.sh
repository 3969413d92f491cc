#!/bin/bash
# Build a tuning variant of the library here (no GPU needed) into variants/<name>.so; on the GPU box copy it over the
# in-tree library before a run:   cp variants/<name>.so humanoid_real_time_retarget_b200/libhrt_b200.so
#   tools/build_variant.sh <name> "<extra nvcc flags>"
set -e
mkdir -p variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -cudart shared \
  -Xlinker -rpath=/usr/local/cuda/lib64 $2 -o variants/$1.so humanoid_real_time_retarget_b200/csrc/hrt_api.cu
echo built variants/$1.so
