#!/usr/bin/env python
"""One block of key ncu metrics per kernel launch of an .ncu-rep (`ncu --set full`), the layout of
profiles/round2/ncu_main_kernels*.txt.  Only the LAST launch of each distinct (kernel, grid) pair is printed, so a capture
of a whole bench run reports the steady state.   python tools/ncu_kernel_table.py gpurun_out/x.ncu-rep [min_grid] > out.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

min_grid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
last = {}
for r in rows[2:]:
    try:
        grid = int(r[col["launch__grid_size"]].replace(",", ""))
    except (KeyError, ValueError):
        grid = 0
    if grid >= min_grid:
        last[(r[col["Kernel Name"]], grid)] = r
for (name, grid), r in last.items():
    print(f"  {'Kernel Name':<76}{name}")
    for w in WANT:
        if w in col:
            print(f"  {w:<76}{r[col[w]]} {units[col[w]]}")
    print()
