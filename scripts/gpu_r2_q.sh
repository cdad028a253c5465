#!/bin/bash
# round 2, call Q (2 GPUs): multi-rank parity on the final code (NCCL exchange, fused P2P, join/cogroup, multi-rank sort incl. a 3-row case)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2q_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_dist.log; tail -8 gpurun_out/r2q_pytest_dist.log | cut -c1-300
