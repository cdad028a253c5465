#!/bin/bash
# round 2, call J: warp pre-aggregation of the hottest keys (Zipf) — parity + timings, and no regression on the uniform headline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "zipf or skew or hot or reduce_by_key or sentinel or count" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -4 gpurun_out/r2j_pytest.log | cut -c1-300
timeout 600 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops reduce,zipf --reps 3 > gpurun_out/r2j_ops.log 2>&1; cat gpurun_out/r2j_ops.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2j_bench.log 2>&1; tail -1 gpurun_out/r2j_bench.log | cut -c1-200
