#!/bin/bash
# the driver's round-end sequence on one GPU: GPU suite, smoke, reference arm, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-check}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2>gpurun_out/${T}_bench_ref.err; cut -c1-400 gpurun_out/${T}_bench_ref.json
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'sust',round(d['roofline']['frac_sustained'],4),'e2e',round(d['e2e']['value']),'fp32',round(d['fp32']['value']),round(d['fp32']['roofline']['frac'],4),'cpu',d['cpu_baseline']['value'], d['clocks'])
print(d['timing']['windows_ms'], d['fp32']['windows_ms'])
P
tail -4 gpurun_out/${T}_bench_n1.err
