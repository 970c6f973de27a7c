#!/bin/bash
# compute-sanitizer over the GPU suite: memcheck on everything, racecheck (shared-memory hazards of the stage rings)
# on the parity trajectories, the fp32 kernels and the random-geometry cases
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
(time timeout 900 $S --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_fullsize.py::test_fused_equals_twopass_at_8192x2048 --deselect tests/test_gpu_fullsize.py::test_fused_equals_twopass_at_32768x8192 --deselect tests/test_gpu_fullsize.py::test_fp32_packed_kernel_at_8192x2048) > gpurun_out/sanitizer_memcheck_all.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck_all.log
(time timeout 600 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x -k "trajectory_matches or fp32_variant or (random_geometry and (75 or 66 or 130 or 4-33))") > gpurun_out/sanitizer_racecheck_all.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck_all.log
