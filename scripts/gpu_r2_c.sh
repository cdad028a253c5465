#!/bin/bash
# round 2, call C: sweep pass (one-kernel radix pass) parity + per-operator timings + CUB yardstick
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -5 gpurun_out/r2c_pytest.log
timeout 900 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sortkv,sort,join --reps 3 > gpurun_out/r2c_ops.log 2>&1; echo "ops rc=$?"; cat gpurun_out/r2c_ops.log | cut -c1-400
VEGA_B200_NO_SWEEP=1 timeout 900 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sort --reps 2 > gpurun_out/r2c_ops_nosweep.log 2>&1; cat gpurun_out/r2c_ops_nosweep.log | cut -c1-400
timeout 600 ./bench_micro/cub_yardstick 1e9 > gpurun_out/r2c_cub.log 2>&1; cat gpurun_out/r2c_cub.log
