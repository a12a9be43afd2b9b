#!/usr/bin/env python3
"""Print the key metrics of an .ncu-rep (first kernel) -- used to write the profiles/*.md summaries."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum","sm__inst_executed_pipe_fma.sum","sm__inst_executed_pipe_lsu.sum",
        "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct", "smsp__warps_eligible.avg.per_cycle_active", "local_load", "lmem")
for h, u, v in zip(hdr, units, vals):
    if h in want or ("issue_stalled" in h and h.endswith("per_warp_active.pct")) or "local" in h and "sum" in h and "inst" in h:
        print(f"{h} [{u}] = {v}")
