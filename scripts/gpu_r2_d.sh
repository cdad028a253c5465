#!/bin/bash
# round 2, call D: sweep pass with the windowed look-back (parity of the paths that use it + timings), N2/N3 tests, evict_last A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sort or group or partition_first or concurrent or staging or range or intersection or subtract or set_ops or join or cogroup" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; tail -5 gpurun_out/r2d_pytest.log
timeout 600 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sort --reps 2 > gpurun_out/r2d_ops.log 2>&1; echo "ops rc=$?"; cat gpurun_out/r2d_ops.log | cut -c1-400
cd bench_micro; for v in base evl; do echo "== $v" >> ../gpurun_out/r2d_micro_evl.log; timeout 120 ./micro_r2_$v 2.5e8 1e6 1 >> ../gpurun_out/r2d_micro_evl.log 2>&1; done; cd ..
cat gpurun_out/r2d_micro_evl.log
for v in lb8 lb32; do echo "== $v" >> gpurun_out/r2d_ops_variants.log; VEGA_B200_LIB=$PWD/vega_b200/libvega_b200_$v.so timeout 300 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sort --reps 2 >> gpurun_out/r2d_ops_variants.log 2>&1; done; cat gpurun_out/r2d_ops_variants.log | cut -c1-300
timeout 300 python tools/bench_ops.py --rows 1.25e8 --distinct 1e6 --ops join --reps 3 > gpurun_out/r2d_join.log 2>&1; cat gpurun_out/r2d_join.log | cut -c1-500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_sweep -s 3 -c 3 -f -o gpurun_out/r2d_prof_sweep python tools/bench_ops.py --rows 2.5e8 --distinct 1e6 --ops group --reps 1 > gpurun_out/r2d_ncu_sweep.log 2>&1; echo "ncu rc=$?"
