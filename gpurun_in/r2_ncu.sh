#!/bin/bash
# round 2 ncu evidence: launch lists (fp64, fp32) and one --set full capture of each fused step kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --reps 1 --no-cpu --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/launches_r2.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f32_r2.csv $B --dtype f32 > gpurun_out/launches_f32_r2.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 6 -c 1 -f -o gpurun_out/ncu_fused_f64_r2 $B --no-e2e > gpurun_out/ncu_f64.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_f32p -s 6 -c 1 -f -o gpurun_out/ncu_fused_f32_r2 $B --no-e2e --dtype f32 > gpurun_out/ncu_f32.out 2>&1
# the per-GPU slab of configs[4] at 8 GPUs (4096 x 8192), fp64: traffic figure for the strong-scaling leg
ncu --set full --clock-control none -k regex:k_fused -s 6 -c 1 -f -o gpurun_out/ncu_fused_f64_c5slab_r2 $B --no-e2e --H 8192 --W 4096 > gpurun_out/ncu_c5.out 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_f64.out gpurun_out/ncu_f32.out gpurun_out/ncu_c5.out
