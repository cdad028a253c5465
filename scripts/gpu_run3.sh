set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ./bench_micro/micro 2.5e8 1e6 > gpurun_out/micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/micro.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_1e9.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_1e9.log
