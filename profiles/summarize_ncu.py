#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches_summary.md
  python profiles/summarize_ncu.py full gpurun_out/prof_r01_X.ncu-rep profiles/r01_X_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict, defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__sass_inst_executed_op_shared_ld.sum",
]


PIPELINE = ("fast_blur_kernel", "select_kernel", "brief_kernel", "hamming_pipeline_kernel", "lr_filter_kernel", "track_filter_kernel",
            "ransac_prepare_kernel", "fundmat_kernel", "consistency_kernel", "gn_kernel", "begin_step_kernel", "scatter_cams_kernel",
            "rbr_match_kernel", "win_track_kernel", "upload_meta_kernel", "pyramid_kernel", "flatten_octaves_kernel",
            "reset_ids_octaves_kernel")


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    per = defaultdict(list)
    order = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        if k not in per:
            order.append(k)
        per[k].append(float(r["Metric Value"]) * (1e-3 if r["Metric Unit"] == "ns" else 1.0))
    total = sum(sum(v) for v in per.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- cold-cache, serialised launches: compare SHARES.\n\n")
        f.write("| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for k in sorted(per, key=lambda k: -sum(per[k])):
            v = per[k]
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {sum(v)/total:.3f} |\n")
        # the kernels of one pipeline step only (the microbenches of bench.py launch others)
        step = [k for k in per if any(t in k for t in PIPELINE)]
        tot2 = sum(sum(per[k]) for k in step) or 1.0
        f.write("\n## pipeline kernels only (share of a step)\n\n| kernel | launches | mean us | share of step |\n|---|---:|---:|---:|\n")
        for k in sorted(step, key=lambda k: -sum(per[k])):
            v = per[k]
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v)/tot2:.3f} |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for v in vals:
            d = OrderedDict(zip(hdr, v))
            u = dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u.get(k, '')} |\n")
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
