#!/bin/bash
# round 2, call X: full GPU suite on the gather sweep (1024x1 for rows with values, 512x2 key-only, double-buffered keys),
# translation fused into the first histogram, vectorised csr_bounds; operator timings; ncu of the three row types
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log; tail -5 gpurun_out/r2x_pytest.log | cut -c1-400
timeout 600 python tools/bench_ops.py --ops group,sort,sortkv,join --reps 3 > gpurun_out/r2x_ops.jsonl 2> gpurun_out/r2x_ops.err; tail -4 gpurun_out/r2x_ops.jsonl | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rp_gsweep|rp_hist" -s 6 -c 6 -f -o gpurun_out/r2x_gs_group python tools/bench_ops.py --rows 2.5e8 --ops group --reps 1 > gpurun_out/r2x_ncu1.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_gsweep -s 8 -c 1 -f -o gpurun_out/r2x_gs_sort python tools/bench_ops.py --rows 2.5e8 --ops sort --reps 2 > gpurun_out/r2x_ncu2.log 2>&1; echo "rc=$?"
