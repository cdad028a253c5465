#!/bin/bash
# round 2, call Z1: CSR offsets recorded by the last sort pass, vectorised translate, branch-free histogram fast path — parity + timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group or golden or gather_sweep or join or cogroup or config or sort" > gpurun_out/r2z1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z1_pytest.log; tail -4 gpurun_out/r2z1_pytest.log | cut -c1-300
timeout 600 python tools/bench_ops.py --ops group,sort,sortkv --reps 3 > gpurun_out/r2z1_ops.jsonl 2> gpurun_out/r2z1_ops.err; tail -3 gpurun_out/r2z1_ops.jsonl | cut -c1-330
VEGA_B200_NO_CSR_FUSION=1 timeout 600 python tools/bench_ops.py --ops group --reps 3 > gpurun_out/r2z1_ops_nocsr.jsonl 2> gpurun_out/r2z1_ops_nocsr.err; tail -1 gpurun_out/r2z1_ops_nocsr.jsonl | cut -c1-330
