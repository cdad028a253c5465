#!/bin/bash
# round 2, call G (2 GPUs): where does the per-step exchange overhead go? (host wall-clock breakdown in vb_xstats)
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2g_bench_2gpu.log 2>&1; tail -1 gpurun_out/r2g_bench_2gpu.log | cut -c1-300
