#!/bin/bash
# round 2, call U: gather sweep with double-buffered keys — parity, then A/B of the buffering choices at 1e9 rows
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_sweep or sort_by_key_matches or group_by_key_matches or config1" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log; tail -5 gpurun_out/r2u_pytest.log | cut -c1-400
for v in "" _gs_dbnarrow _gs_nodb; do
  VEGA_B200_LIB=$PWD/vega_b200/libvega_b200$v.so timeout 600 python tools/bench_ops.py --ops group,sort,sortkv --reps 2 > gpurun_out/r2u_ops$v.jsonl 2> gpurun_out/r2u_ops$v.err; tail -3 gpurun_out/r2u_ops$v.jsonl | cut -c1-330
done
