#!/bin/bash
# round 2, call E: where does the sweep pass spend its time? (timing-bisection builds: results of those are wrong by design) + ncu capture + join
set -x
mkdir -p gpurun_out
for v in "" _nolb _nowrite _norank _lb8 _lb32; do echo "== variant '$v'" >> gpurun_out/r2e_ops_variants.log; VEGA_B200_LIB=$PWD/vega_b200/libvega_b200$v.so timeout 300 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group --reps 2 >> gpurun_out/r2e_ops_variants.log 2>&1; done; cat gpurun_out/r2e_ops_variants.log | cut -c1-330
timeout 300 python tools/bench_ops.py --rows 5e8 --distinct 1e6 --ops join --reps 3 > gpurun_out/r2e_join.log 2>&1; cat gpurun_out/r2e_join.log | cut -c1-500
VEGA_B200_EAGER_COGROUP=1 timeout 300 python tools/bench_ops.py --rows 5e8 --distinct 1e6 --ops join --reps 2 > gpurun_out/r2e_join_eager.log 2>&1; cat gpurun_out/r2e_join_eager.log | cut -c1-500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rp_sweep -s 3 -c 3 -f -o gpurun_out/r2e_prof_sweep python tools/bench_ops.py --rows 2.5e8 --distinct 1e6 --ops group --reps 1 > gpurun_out/r2e_ncu_sweep.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "join or cogroup or intersection or group" > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
