#!/bin/bash
# round 2, call W: gather sweep in global tile order (tile histogram + column scan) — parity, then CTA shape / buffering A/B at 1e9 rows
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_sweep or sort_by_key_matches or group_by_key_matches or golden or config1 or sort_skips or join_matches" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log; tail -5 gpurun_out/r2w_pytest.log | cut -c1-400
for v in "" _gs_v512 _gs_v512n _gs_k1024; do
  VEGA_B200_LIB=$PWD/vega_b200/libvega_b200$v.so timeout 600 python tools/bench_ops.py --ops group,sort,sortkv --reps 2 > gpurun_out/r2w_ops$v.jsonl 2> gpurun_out/r2w_ops$v.err; tail -3 gpurun_out/r2w_ops$v.jsonl | cut -c1-330
done
