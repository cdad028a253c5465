set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -12 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_dist.py --rows 1e9 --ops group,join --p2p > gpurun_out/dist_2gpu_p2p.log 2>&1; echo "rc=$?"; grep -E "^\{|Error|error" gpurun_out/dist_2gpu_p2p.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/bench_dist.py --rows 1e9 --ops group,join > gpurun_out/dist_2gpu_nccl.log 2>&1; echo "rc=$?"; grep -E "^\{|Error|error" gpurun_out/dist_2gpu_nccl.log | tail -6
