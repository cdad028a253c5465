set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/bench_ops.py --rows 1e9 --distinct 1e6 > gpurun_out/ops_1e9.log 2>&1; echo "ops rc=$?"; cat gpurun_out/ops_1e9.log | tail -20
timeout 900 ncu --set full --clock-control none --import-source on -k hash_agg_kernel -s 16 -c 2 -f -o gpurun_out/prof_hash_agg_r1b python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
