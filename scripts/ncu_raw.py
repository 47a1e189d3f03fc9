#!/usr/bin/env python
"""Key metrics from `ncu -i X.ncu-rep --page raw --csv`.  usage: ncu_raw.py file.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
        "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu",
        "sm__inst_executed_pipe_fma", "sm__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_pipe_xu", "sm__inst_executed_pipe_fmaheavy", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active"]
for i, h in enumerate(hdr):
    if any(h == w or (h.startswith(w) and "pct_of_peak_sustained_active" in h and h.count(".") <= 3) for w in want):
        print(f"{h} [{units[i]}]: {[r[i] for r in rows[2:]]}")
