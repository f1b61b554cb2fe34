#!/usr/bin/env python
"""Markdown summary of one `ncu --set full` capture: per launch the metrics DESIGN.md quotes.

  python profiles/ncu_step_summary.py gpurun_out/r02aa_step.ncu-rep > profiles/r02aa_step_full.md
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
           "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    print(f"# ncu --set full summary ({rep})\n")
    print("`ncu --set full --clock-control none`: cold-cache, serialised launches (absolute times are not bench values).\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        block = r[col["Block Size"]] if "Block Size" in col else ""
        print(f"## {name}  grid {grid} block {block}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in col:
                print(f"| {m} | {r[col[m]]} | {units[col[m]]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
