#!/bin/bash
# round 2, call V: gather sweep CTA shape A/B (512x2 vs 1024x1, which row types keep two key buffers) at 1e9 rows
set -x
mkdir -p gpurun_out
for v in "" _gs_1024 _gs_1024db; do
  VEGA_B200_LIB=$PWD/vega_b200/libvega_b200$v.so timeout 600 python tools/bench_ops.py --ops group,sort,sortkv --reps 2 > gpurun_out/r2v_ops$v.jsonl 2> gpurun_out/r2v_ops$v.err; tail -3 gpurun_out/r2v_ops$v.jsonl | cut -c1-330
done
VEGA_B200_LIB=$PWD/vega_b200/libvega_b200_gs_1024.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_sweep" > gpurun_out/r2v_pytest_1024.log 2>&1; tail -3 gpurun_out/r2v_pytest_1024.log
