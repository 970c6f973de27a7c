#!/bin/bash
# correctness of the stealing build on the cases where ranges are long enough to be taken over, then A/B
export FDLBM_LIB=$PWD/gpurun_in/variants/lib_steal.so
(timeout 150 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_slab.py -m gpu -x -q 2>&1 | tail -5) | tee gpurun_out/steal_tests.log
unset FDLBM_LIB
timeout 200 bash gpurun_in/ab.sh "--steps 200 --warmup 5" base steal 2>&1 | tee gpurun_out/ab_steal.txt
