#!/bin/bash
# Build named variants of libfdlbm.so HERE (no GPU needed) into gpurun_in/variants/ for A/B runs on the box:
#   gpurun_in/build_variants.sh base "" steal4 "-DFDLBM_STEAL=1 -DFDLBM_STEAL_EVERY=4 -DFDLBM_STEAL_MIN=28"
#   gpurun -- 'bash gpurun_in/ab.sh "--steps 200 --warmup 5" base steal4'
# and, for the correctness of a variant:  FDLBM_LIB=$PWD/gpurun_in/variants/lib_steal4.so python -m pytest tests -m gpu
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_in/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc $F $flags -o gpurun_in/variants/lib_$name.so fingering_dynamics_b200/csrc/fdlbm.cu 2>&1 | grep -E " error|spill stores, [1-9]" || true
  echo "built gpurun_in/variants/lib_$name.so ($flags)"
done
