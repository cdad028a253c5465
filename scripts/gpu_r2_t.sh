#!/bin/bash
# round 2, call T: rerun the gather-sweep parity test, then ncu --set full (with source) of rp_gsweep_kernel in group_by_key and the sorts at 2.5e8 rows
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_sweep or fallback" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log; tail -5 gpurun_out/r2t_pytest.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_gsweep -s 3 -c 3 -f -o gpurun_out/r2t_gs_group python tools/bench_ops.py --rows 2.5e8 --ops group --reps 1 > gpurun_out/r2t_ncu1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2t_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_gsweep -s 8 -c 1 -f -o gpurun_out/r2t_gs_sort python tools/bench_ops.py --rows 2.5e8 --ops sort --reps 2 > gpurun_out/r2t_ncu2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2t_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_gsweep -s 8 -c 1 -f -o gpurun_out/r2t_gs_sortkv python tools/bench_ops.py --rows 2.5e8 --ops sortkv --reps 2 > gpurun_out/r2t_ncu3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2t_ncu3.log
