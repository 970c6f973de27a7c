#!/bin/bash
# sizing experiment for two steps per pass (FDLBM_F32_MOCK2 / FDLBM_F32_PADSMEM, lbm_fused_f32.cuh): A/B of the four
# variants + one ncu capture of the mock at 2 CTAs/SM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
bash gpurun_in/ab.sh "--dtype f32 --steps 100 --warmup 5 --no-extras" f32base f32pad f32mock f32mockpad 2>&1 | tee gpurun_out/r2_ab_mock2.txt
B="python bench.py --steps 5 --warmup 3 --reps 1 --no-cpu --no-extras --no-e2e --dtype f32"
for v in f32mockpad f32mock; do
FDLBM_LIB=$PWD/gpurun_in/variants/lib_$v.so ncu --set full --clock-control none --import-source on -k regex:k_fused_f32p -s 6 -c 1 -f -o gpurun_out/ncu_$v $B > gpurun_out/ncu_$v.out 2>&1
tail -n 2 gpurun_out/ncu_$v.out
done
