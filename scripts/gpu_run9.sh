set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_ops.py --rows 1e9 --ops reduce,zipf,count --reps 3 2>&1 | tail -4
timeout 600 ./bench_micro/micro 2.5e8 1e6 2>&1 | grep -E "ctas/SM (3|6)|table"
