set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/cardinality.log
for D in 1e7 1e8; do
  timeout 300 python tools/bench_ops.py --rows 1e9 --distinct $D --ops reduce --reps 2 2>&1 | grep "^{" | sed "s/^{/{\"distinct\": $D, /" >> gpurun_out/cardinality.log
  VEGA_B200_NO_PARTITION=1 timeout 300 python tools/bench_ops.py --rows 1e9 --distinct $D --ops reduce --reps 2 2>&1 | grep "^{" | sed "s/^{/{\"distinct\": $D, \"no_partition\": true, /" >> gpurun_out/cardinality.log
done
cat gpurun_out/cardinality.log | cut -c1-360
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
