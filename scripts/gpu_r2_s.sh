#!/bin/bash
# round 2, call S: gather sweep (csrc/gsweep.cuh) — parity on sort/group cases, then A/B against rp_sweep_kernel<STATIC> at 1e9 rows
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_sweep or sort_by_key_matches or group_by_key_matches or golden or config1 or sort_skips or fallback" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log; tail -15 gpurun_out/r2s_pytest.log | cut -c1-400
timeout 600 python tools/bench_ops.py --ops group,sort,sortkv --reps 3 > gpurun_out/r2s_ops_gsweep.jsonl 2> gpurun_out/r2s_ops_gsweep.err; tail -3 gpurun_out/r2s_ops_gsweep.jsonl | cut -c1-400
VEGA_B200_GS_STAGGER=2500 timeout 600 python tools/bench_ops.py --ops group,sort,sortkv --reps 3 > gpurun_out/r2s_ops_gsweep_stagger.jsonl 2> gpurun_out/r2s_ops_gsweep_stagger.err; tail -3 gpurun_out/r2s_ops_gsweep_stagger.jsonl | cut -c1-400
VEGA_B200_NO_GSWEEP=1 timeout 600 python tools/bench_ops.py --ops group,sort,sortkv --reps 3 > gpurun_out/r2s_ops_static.jsonl 2> gpurun_out/r2s_ops_static.err; tail -3 gpurun_out/r2s_ops_static.jsonl | cut -c1-400
