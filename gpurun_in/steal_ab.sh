#!/bin/bash
# A/B of the column-range stealing knobs (fp64, 8192x2048).  Build the variants HERE first:
#   gpurun_in/build_variants.sh base "" steal1 "-DFDLBM_STEAL=1" \
#       steal4 "-DFDLBM_STEAL=1 -DFDLBM_STEAL_EVERY=4 -DFDLBM_STEAL_MIN=28" \
#       steal8 "-DFDLBM_STEAL=1 -DFDLBM_STEAL_EVERY=8 -DFDLBM_STEAL_MIN=48"
# then:  gpurun --timeout 400 -- 'bash gpurun_in/steal_ab.sh'
mkdir -p gpurun_out
for v in steal4 steal8; do   # results must stay bit-identical (duplicated columns are written with the same values)
  FDLBM_LIB=$PWD/gpurun_in/variants/lib_$v.so timeout 120 python -m pytest tests/test_gpu_fullsize.py tests/test_slab.py -m gpu -x -q \
      -k "8192 or translation or 258 or porous" 2>&1 | tail -2 | sed "s/^/$v: /"
done | tee gpurun_out/steal_tests.log
timeout 250 bash gpurun_in/ab.sh "--steps 200 --warmup 5" base steal1 steal4 steal8 2>&1 | tee gpurun_out/ab_steal_every.txt
