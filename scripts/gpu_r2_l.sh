#!/bin/bash
# round 2, call L: the sweep tile pipeline fed by per-part offsets (rp_sweep_kernel<STATIC>) as the default scatter for row streams:
# full parity + A/B against rp_scatter_kernel (VEGA_B200_NO_SWEEP_STATIC=1)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -4 gpurun_out/r2l_pytest.log | cut -c1-300
timeout 600 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sortkv,sort,join --reps 3 > gpurun_out/r2l_ops.log 2>&1; cat gpurun_out/r2l_ops.log | cut -c1-300
VEGA_B200_NO_SWEEP_STATIC=1 timeout 600 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sort --reps 2 > gpurun_out/r2l_ops_old.log 2>&1; cat gpurun_out/r2l_ops_old.log | cut -c1-300
