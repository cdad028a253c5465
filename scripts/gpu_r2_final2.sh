#!/bin/bash
# round 2, final 1-GPU validation of the gather-sweep build: full parity, smoke, bench (driver's flags), operator timings, launch lists
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2f_bench.log | cut -c1-200
timeout 900 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops reduce,count,zipf,group,sortkv,sort,join --reps 3 > gpurun_out/r2f_ops.log 2>&1; cat gpurun_out/r2f_ops.log | cut -c1-250
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_group.csv python tools/bench_ops.py --rows 1e9 --ops group,sort --reps 1 > gpurun_out/r2f_ncu_group.log 2>&1; echo "ncu group rc=$?"
