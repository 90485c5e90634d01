#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: one block per profiled launch with the metrics the roofline uses.
usage: ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_summary.py [substring ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "sm__cycles_active.avg"]
extra = sys.argv[1:]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        base = h.split(".", 2)[-1] if h.count(".") >= 3 and h.split(".")[1].startswith("Triage") else h
        if h in KEYS or base in KEYS or any(e in h for e in extra):
            print(f"   {h} = {r[i]} {units[i]}")
