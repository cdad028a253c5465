#!/bin/bash
# round 2, call K: after reverting the hot-key promotion — full parity, Zipf with bulk- vs register-staged hash_agg, bench + launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log | cut -c1-300
timeout 600 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops reduce,count,zipf,group,sortkv,sort,join --reps 3 > gpurun_out/r2k_ops.log 2>&1; cat gpurun_out/r2k_ops.log | cut -c1-300
VEGA_B200_NO_BULK=1 timeout 600 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops zipf --reps 3 > gpurun_out/r2k_ops_zipf_nobulk.log 2>&1; cat gpurun_out/r2k_ops_zipf_nobulk.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.log 2>&1; tail -1 gpurun_out/r2k_bench.log | cut -c1-200
