#!/bin/bash
# round 2, call I (8 GPUs): BASELINE configs[3]/[4] with oracle-sampled parity, group/sort exchange rates, bench.py weak scaling point
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2i_bench_8gpu.log 2>&1; tail -1 gpurun_out/r2i_bench_8gpu.log | cut -c1-300
timeout 900 $T --master-port 29532 tools/bench_dist.py --rows 1e9 --ops zipf,group,join,sort --reps 2 > gpurun_out/r2i_dist_nccl.log 2>&1; grep '^{' gpurun_out/r2i_dist_nccl.log | cut -c1-500
timeout 600 $T --master-port 29533 tools/bench_dist.py --rows 1e9 --ops group,join --reps 2 --p2p > gpurun_out/r2i_dist_p2p.log 2>&1; grep '^{' gpurun_out/r2i_dist_p2p.log | cut -c1-500
