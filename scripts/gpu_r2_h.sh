#!/bin/bash
# round 2, call H: full GPU parity after the join / N2 / N3 / table-split changes + bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2h_bench.log | cut -c1-300
timeout 600 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops reduce,zipf,group --reps 3 > gpurun_out/r2h_ops.log 2>&1; cat gpurun_out/r2h_ops.log | cut -c1-400
