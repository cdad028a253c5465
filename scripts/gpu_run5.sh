set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rp_" -s 7 -c 3 -f -o gpurun_out/prof_rp_r1 python tools/bench_ops.py --rows 2.5e8 --ops group --reps 1 > gpurun_out/ncu_rp.log 2>&1; echo "ncu rp rc=$?"; tail -3 gpurun_out/ncu_rp.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1e9.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_1e9.log
