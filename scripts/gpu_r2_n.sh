#!/bin/bash
# round 2, call N: sweep scatter CTA shapes: default (512x2 with values, 256x4 key-only), 1024x1 with values, 256 x 16 keys
set -x
mkdir -p gpurun_out
for v in "" _v1024 _k16; do echo "== variant '$v'" >> gpurun_out/r2n_ops.log; VEGA_B200_LIB=$PWD/vega_b200/libvega_b200$v.so timeout 300 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sortkv,sort --reps 2 >> gpurun_out/r2n_ops.log 2>&1; done; cat gpurun_out/r2n_ops.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sort or group or partition_first or sweep" > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
