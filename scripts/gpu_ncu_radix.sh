set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rp_(hist|scatter)" -s 4 -c 4 -f -o gpurun_out/prof_rp_sortkeys python tools/bench_ops.py --rows 2.5e8 --ops sort --reps 1 > gpurun_out/ncu_rp1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_rp1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rp_(hist|scatter)" -s 4 -c 4 -f -o gpurun_out/prof_rp_sortkv python tools/bench_ops.py --rows 2.5e8 --ops sortkv --reps 1 > gpurun_out/ncu_rp2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_rp2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rp_(hist|scatter)" -s 2 -c 4 -f -o gpurun_out/prof_rp_group python tools/bench_ops.py --rows 2.5e8 --ops group --reps 1 > gpurun_out/ncu_rp3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_rp3.log
