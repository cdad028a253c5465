set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/bench_dist.py --rows 1e9 --ops group,join --p2p > gpurun_out/dist_8gpu_p2p.log 2>&1; echo "rc=$?"; grep -E "^\{|Error|error" gpurun_out/dist_8gpu_p2p.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_8gpu.log 2>&1; echo "bench8 rc=$?"; tail -1 gpurun_out/bench_8gpu.log | cut -c1-240
