set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 86 --log-file gpurun_out/memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or sentinel or empty_and_tiny or join_matches_oracle or sort_skips or partition_by_key or bincode or stage_resubmission or shared_table" > gpurun_out/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck.log | head -10
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 86 --log-file gpurun_out/racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_group_by_key_golden or test_sort_skips or test_count_by_value_golden" > gpurun_out/racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
tail -2 gpurun_out/racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/racecheck.log | head -10
