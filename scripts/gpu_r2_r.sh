#!/bin/bash
# round 2, call R: ncu --set full (with source) of rp_sweep_kernel<STATIC> in group_by_key and key-only sort at 2.5e8 rows,
# and a per-kernel launch list of group_by_key at 1e9 rows
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_sweep -s 3 -c 3 -f -o gpurun_out/r2r_sw_group python tools/bench_ops.py --rows 2.5e8 --ops group --reps 1 > gpurun_out/r2r_ncu1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2r_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_sweep -s 8 -c 2 -f -o gpurun_out/r2r_sw_sort python tools/bench_ops.py --rows 2.5e8 --ops sort --reps 2 > gpurun_out/r2r_ncu2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2r_ncu2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r_launches_group.csv python tools/bench_ops.py --rows 1e9 --ops group --reps 1 > gpurun_out/r2r_ncu3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2r_ncu3.log
