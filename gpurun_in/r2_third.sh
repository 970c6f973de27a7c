#!/bin/bash
# per-SM balancer A/B (same box, interleaved)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do for bal in 0 1; do
  for K in 20 200; do
  FDLBM_BALANCE=$bal python bench.py --steps $K --warmup 5 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('persm bal=$bal K=$K rep$rep', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'], [round(w,2) for w in d['timing']['windows_ms']])"
  done
done; done 2>&1 | tee gpurun_out/r2_ab_balance_persm.txt
BD_TAG=persm FDLBM_BALANCE=1 python gpurun_in/balance_dump.py 2>&1 | tail -14 | tee gpurun_out/r2_balance_dump_persm.txt
