#!/bin/bash
# N=2: the driver's launch line (K=20), with all extra legs (multi_gpu_check, strong_c5) and a per-step timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 --timeline gpurun_out/r2_timeline_n$N.json ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 2500 gpurun_out/r2_bench_n$N.json
grep -v "^W\|^$" gpurun_out/r2_bench_n$N.err | tail -12
