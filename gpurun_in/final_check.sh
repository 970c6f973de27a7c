#!/bin/bash
# round-end verification on one B200: GPU tests, smoke, bench lines (fp64 with the CPU leg, fp32, reference arm)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/final_gpu.txt
(time timeout 420 python -m pytest tests -m gpu -x -q) > gpurun_out/final_pytest.log 2>&1
tail -3 gpurun_out/final_pytest.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1
tail -1 gpurun_out/final_smoke.log
timeout 150 python bench.py --steps 300 --warmup 5 > gpurun_out/final_bench_f64.json 2> gpurun_out/final_bench_f64.err
cut -c1-400 gpurun_out/final_bench_f64.json
timeout 100 python bench.py --steps 300 --warmup 5 --dtype f32 --no-cpu > gpurun_out/final_bench_f32.json 2> gpurun_out/final_bench_f32.err
cut -c1-300 gpurun_out/final_bench_f32.json
timeout 60 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/final_bench_ref.json 2>&1
cut -c1-300 gpurun_out/final_bench_ref.json
