#!/bin/bash
# round 2, call Y: translation fused into the first histogram with all look-ups in flight; group parity + timings + launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group or golden or gather_sweep or join or cogroup or config" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log; tail -4 gpurun_out/r2y_pytest.log | cut -c1-300
timeout 600 python tools/bench_ops.py --ops group,zipf,zipfgroup --reps 3 > gpurun_out/r2y_ops.jsonl 2> gpurun_out/r2y_ops.err; tail -3 gpurun_out/r2y_ops.jsonl | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_launches_group.csv python tools/bench_ops.py --rows 1e9 --ops group --reps 1 > gpurun_out/r2y_ncu.log 2>&1; echo "rc=$?"
