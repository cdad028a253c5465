set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_ops.py --rows 1e9 --distinct 1e6 > gpurun_out/ops_1e9.log 2>&1; echo "ops rc=$?"; cat gpurun_out/ops_1e9.log | tail -20
