#!/bin/bash
# round 2, call Z2: group_by_key over AoS rows — the dictionary kernel emits the values as a column (first sort pass reads 12 B/row, not 20)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group or golden or gather_sweep or join or cogroup or config or reduce_by_key_matches" > gpurun_out/r2z2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z2_pytest.log; tail -4 gpurun_out/r2z2_pytest.log | cut -c1-300
timeout 600 python tools/bench_ops.py --ops reduce,group,join --reps 3 > gpurun_out/r2z2_ops.jsonl 2> gpurun_out/r2z2_ops.err; tail -3 gpurun_out/r2z2_ops.jsonl | cut -c1-330
VEGA_B200_NO_DICT_VALS=1 timeout 600 python tools/bench_ops.py --ops group --reps 3 > gpurun_out/r2z2_ops_nocol.jsonl 2> gpurun_out/r2z2_ops_nocol.err; tail -1 gpurun_out/r2z2_ops_nocol.jsonl | cut -c1-330
