set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rp_(hist|scatter)" -s 4 -c 2 -f -o gpurun_out/prof_rp_group_ballot python tools/bench_ops.py --rows 2.5e8 --ops group --reps 1 > gpurun_out/ncu_rp3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_rp3.log
timeout 900 python tools/bench_ops.py --rows 1e9 --distinct 1e6 --ops group,sort --reps 2 2>&1 | tail -3
