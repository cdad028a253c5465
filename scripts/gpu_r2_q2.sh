#!/bin/bash
# round 2, call Q2 (2 GPUs): multi-rank parity on the gather-sweep build (NCCL exchange, fused P2P, join/cogroup, multi-rank sort)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2q2_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q2_pytest_dist.log; tail -6 gpurun_out/r2q2_pytest_dist.log | cut -c1-300
