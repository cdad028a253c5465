#!/bin/bash
# round 2, call M: sweep scatter with 256-thread CTAs x 4 per SM (variant build) vs the default 512 x 2
set -x
mkdir -p gpurun_out
for v in "" _t256; do echo "== variant '$v'" >> gpurun_out/r2m_ops.log; VEGA_B200_LIB=$PWD/vega_b200/libvega_b200$v.so timeout 300 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sortkv,sort --reps 2 >> gpurun_out/r2m_ops.log 2>&1; done; cat gpurun_out/r2m_ops.log | cut -c1-300
VEGA_B200_LIB=$PWD/vega_b200/libvega_b200_t256.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sort or group or partition_first" > gpurun_out/r2m_pytest_t256.log 2>&1; tail -3 gpurun_out/r2m_pytest_t256.log
