#!/bin/bash
# usage: ab2.sh "<bench args>" variant...   -- like ab.sh, for fp64 AND fp32 in one go (one repetition each, interleaved twice)
args=$1; shift
for rep in 1 2; do
for v in "$@"; do
for dt in f64 f32; do
  FDLBM_LIB=$PWD/gpurun_in/variants/lib_$v.so python bench.py $args --dtype $dt --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $dt rep$rep', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done; done
