set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -15 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 rc=$?"; tail -4 gpurun_out/bench_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_2gpu_ref.log 2>&1; echo "ref rc=$?"; tail -2 gpurun_out/bench_2gpu_ref.log
