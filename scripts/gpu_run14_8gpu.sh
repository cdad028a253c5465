set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/bench_dist.py --rows 1e9 > gpurun_out/dist_8gpu.log 2>&1; echo "rc=$?"; grep -E "^\{|Error|error" gpurun_out/dist_8gpu.log | tail -8
