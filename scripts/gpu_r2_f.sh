#!/bin/bash
# round 2, call F (2 GPUs): the exchange inside libvega_b200 (NCCL grouped send/recv, fused P2P, multi-rank sort) — parity + timings
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r2f_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_dist.log; tail -15 gpurun_out/r2f_pytest_dist.log | cut -c1-300
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_2gpu.log 2>&1; tail -1 gpurun_out/r2f_bench_2gpu.log | cut -c1-600
timeout 600 $T --master-port 29512 tools/bench_dist.py --rows 2.5e8 --ops zipf,group,join --reps 2 > gpurun_out/r2f_dist_nccl.log 2>&1; grep '^{' gpurun_out/r2f_dist_nccl.log | cut -c1-700
timeout 600 $T --master-port 29513 tools/bench_dist.py --rows 2.5e8 --ops group,join --reps 2 --p2p > gpurun_out/r2f_dist_p2p.log 2>&1; grep '^{' gpurun_out/r2f_dist_p2p.log | cut -c1-700
