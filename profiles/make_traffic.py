#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of ONE serialised step of bench.py (--chunks 1 --batch 512):
per kernel the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), the warp-instructions executed
(smsp__inst_executed.sum) and ncu's SM throughput of that launch.  bench.py reads these as STATIC companions of the live
timings (roofline.traffic, roofline.binding_resource).

  python profiles/make_traffic.py gpurun_out/r02g_step.ncu-rep 512 "<how it was captured>" > profiles/traffic.json
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1, "%": 1, "us": 1, "ns": 1e-3, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}


def main(rep, pairs, how):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(hdr)}

    def val(r, key):
        if key not in col:
            return None
        try:
            return float(r[col[key]].replace(",", "")) * UNIT.get(units[col[key]], 1)
        except ValueError:
            return None
    res = {"source": f"{rep.split('/')[-1]}: {how}", "pairs_per_launch": int(pairs), "dram_bytes_per_launch": {},
           "warp_instructions_per_launch": {}, "sm_throughput_pct": {}, "duration_us": {}}
    # the capture window may hold more than one step and start in the middle of one: cut it at begin_step launches and
    # keep the complete step with the longest total duration (a tracking step; the first frame of a sequence skips stages 3-5)
    steps, cur = [], None
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        if not short:
            continue
        if short == "begin_step_kernel":
            cur = []
            steps.append(cur)
        if cur is not None:
            cur.append((short, r))
    n_kern = max(len(s) for s in steps)
    best = max((s for s in steps if len(s) == n_kern), key=lambda s: sum(val(r, "gpu__time_duration.sum") or 0 for _, r in s))
    seen_hamming = 0
    for short, r in best:
        if short in ("hamming_pipeline_kernel", "hamming_kernel"):
            short = "hamming_kernel(lr)" if seen_hamming == 0 else "hamming_kernel(track)"
            seen_hamming += 1
        if short == "begin_step_kernel":
            short = "begin_step"
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        res["dram_bytes_per_launch"][short] = int((rd or 0) + (wr or 0))
        res["warp_instructions_per_launch"][short] = int(val(r, "smsp__inst_executed.sum") or 0)
        res["sm_throughput_pct"][short] = round(val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0, 1)
        res["duration_us"][short] = round(val(r, "gpu__time_duration.sum") or 0, 2)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
