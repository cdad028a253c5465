set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 60 ./examples/_build/group_by; echo "group_by rc=$?"
timeout 60 ./examples/_build/join; echo "join rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-e2e --no-cpu > gpurun_out/bench_1e8.log 2>&1; echo "bench1e8 rc=$?"; tail -3 gpurun_out/bench_1e8.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1e9.log 2>&1; echo "bench1e9 rc=$?"; tail -3 gpurun_out/bench_1e9.log
