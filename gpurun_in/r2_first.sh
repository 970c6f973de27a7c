#!/bin/bash
# round 2, first GPU session: the new tests, then the reworked bench at the driver's K=20
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests/test_init_device.py tests/test_slab.py tests/test_dropin_gpu.py \
  "tests/test_gpu_fullsize.py::test_benchmark_generator_matches_the_oracle_at_1024x512" \
  "tests/test_gpu_fullsize.py::test_fp32_packed_kernel_at_8192x2048" -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k20.json 2> gpurun_out/r2_bench_n1_k20.err
tail -5 gpurun_out/r2_tests1.log
head -c 3000 gpurun_out/r2_bench_n1_k20.json
tail -5 gpurun_out/r2_bench_n1_k20.err
