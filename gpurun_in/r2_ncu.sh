#!/bin/bash
# round 2 ncu evidence: launch lists (fp64, fp32) and one --set full capture of each fused step kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2}
B="python bench.py --steps 5 --warmup 3 --reps 1 --no-cpu --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $B > gpurun_out/launches_$T.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f32_$T.csv $B --dtype f32 > gpurun_out/launches_f32_$T.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 6 -c 1 -f -o gpurun_out/ncu_fused_f64_$T $B --no-e2e > gpurun_out/ncu_f64.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_f32p -s 6 -c 1 -f -o gpurun_out/ncu_fused_f32_$T $B --no-e2e --dtype f32 > gpurun_out/ncu_f32.out 2>&1
# the per-GPU slab of configs[4] at 8 GPUs (4096 x 8192), fp64: traffic figure for the strong-scaling leg
ncu --set full --clock-control none -k regex:k_fused -s 6 -c 1 -f -o gpurun_out/ncu_fused_f64_c5slab_$T $B --no-e2e --H 8192 --W 4096 > gpurun_out/ncu_c5.out 2>&1
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/ncu_f64.out gpurun_out/ncu_f32.out gpurun_out/ncu_c5.out; do tail -n 2 $f; done
