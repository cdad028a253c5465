#!/usr/bin/env python
"""ncu_summary.py <report.ncu-rep> [...] — the handful of counters DESIGN.md quotes, per captured launch
(reads the report with `ncu -i ... --page raw --csv`; no GPU needed)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    stall = [(x, i) for i, x in enumerate(h) if "smsp__average_warp" in x and "issue_stalled" in x and x.endswith(".ratio") and "not_issued" not in x]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[ki]}   [{rep}]")
        for w in WANT:
            if w in h:
                i = h.index(w)
                print(f"  {w} = {r[i]} {rows[1][i]}")
        s = sorted([(float(r[i].replace(',', '') or 0), x.split("issue_stalled_")[1].split("_per")[0]) for x, i in stall], reverse=True)[:8]
        print("  top stall reasons (warps per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in s))
