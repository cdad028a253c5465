#!/bin/bash
# round 2, call B: full GPU parity (bulk-staged hash_agg with the proxy fence), bench both arms, launch list, ncu --set full of the hot kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -3 gpurun_out/r2b_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2b_bench.log | cut -c1-300
VEGA_B200_NO_BULK=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_bench_nobulk.log 2>&1; tail -1 gpurun_out/r2b_bench_nobulk.log | cut -c1-300
timeout 900 python bench.py --steps 2 --warmup 1 --impl reference > gpurun_out/r2b_bench_ref.log 2>&1; tail -1 gpurun_out/r2b_bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2b_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hash_agg_bulk -s 9 -c 2 -f -o gpurun_out/r2b_prof_hash_agg python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"
