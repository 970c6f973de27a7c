#!/bin/bash
# usage: ab.sh "<bench args>" variant...   -- two interleaved repetitions of each variant
args=$1; shift
for rep in 1 2; do
for v in "$@"; do
  FDLBM_LIB=$PWD/gpurun_in/variants/lib_$v.so python bench.py $args --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v rep$rep', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
