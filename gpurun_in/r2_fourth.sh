#!/bin/bash
# placement A/B (same box, interleaved), then the slab / fullsize tests with placement on
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do for pl in 0 1; do
  for K in 20 200; do
  FDLBM_PLACEMENT=$pl python bench.py --steps $K --warmup 5 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('placement=$pl K=$K rep$rep', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'], [round(w,2) for w in d['timing']['windows_ms']])"
  done
done; done 2>&1 | tee gpurun_out/r2_ab_placement.txt
FDLBM_PLACEMENT=1 python gpurun_in/placement_dump.py 2>&1 | tail -4 | tee gpurun_out/r2_placement_dump.txt
FDLBM_PLACEMENT=0 python gpurun_in/placement_dump.py 2>&1 | tail -2 | tee -a gpurun_out/r2_placement_dump.txt
BD_TAG=c5slab BD_H=8192 BD_W=4096 python gpurun_in/placement_dump.py 2>&1 | tail -2 | tee -a gpurun_out/r2_placement_dump.txt
BD_TAG=c5slab_off FDLBM_PLACEMENT=0 BD_H=8192 BD_W=4096 python gpurun_in/placement_dump.py 2>&1 | tail -2 | tee -a gpurun_out/r2_placement_dump.txt
timeout 1200 python -m pytest tests/test_slab.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
