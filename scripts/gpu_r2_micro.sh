#!/bin/bash
# round 2, call A: parity of the bulk-staged hash_agg + request-port micro-benchmarks + counters
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
cd bench_micro
for v in r4c3s3 r4c2s4 r2c4s3 r8c2s2 r8c1s3; do
  echo "== $v" >> ../gpurun_out/r2a_micro.log
  timeout 300 ./micro_r2_$v 2.5e8 1e6 $([ $v = r4c3s3 ] && echo 0 || echo 1) >> ../gpurun_out/r2a_micro.log 2>&1
done
M="gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_requests_srcunit_tex.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_atom.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file ../gpurun_out/r2a_ncu_micro.csv ./micro_r2_r4c3s3 1.25e8 1e6 0 > ../gpurun_out/r2a_ncu_micro.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file ../gpurun_out/r2a_ncu_micro_r4c2s4.csv -k regex:hash_agg ./micro_r2_r4c2s4 1.25e8 1e6 1 > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file ../gpurun_out/r2a_ncu_micro_r2c4s3.csv -k regex:hash_agg ./micro_r2_r2c4s3 1.25e8 1e6 1 > /dev/null 2>&1
cd ..
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.log 2>&1
