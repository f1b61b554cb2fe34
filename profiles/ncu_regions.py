#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on).  usage: ncu_regions.py REP KERNEL [min_pct]"""
import csv
import io
import subprocess
import sys


def main(rep, kernel, min_pct=1.0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    files = [(i, r[1]) for i, r in enumerate(rows) if r and r[0] == "File Path"]
    hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    tot_s = tot_i = 0
    per_file = []
    for fi, (start, fname) in enumerate(files):
        hi = [i for i in hdrs if i > start][0]
        end = files[fi + 1][0] if fi + 1 < len(files) else len(rows)
        h = rows[hi]
        if "# Samples" not in h:
            continue
        cs, ci = h.index("# Samples"), h.index("Instructions Executed")
        src = []
        for r in rows[hi + 1:end]:
            if len(r) <= ci or not r[0].isdigit():
                continue
            try:
                src.append((int(r[0]), r[1], int(r[cs] or 0), int(r[ci] or 0)))
            except ValueError:
                pass
        per_file.append((fname, src))
        tot_s += sum(x[2] for x in src)
        tot_i += sum(x[3] for x in src)
    print(f"kernel {kernel}: samples {tot_s}, warp instructions {tot_i}")
    for fname, src in per_file:
        fs, fi_ = sum(x[2] for x in src), sum(x[3] for x in src)
        print(f"== {fname.split('/')[-1]}: {100*fs/max(tot_s,1):.1f}% samples, {100*fi_/max(tot_i,1):.1f}% instructions")
        for ln, code, s, i in src:
            if 100 * s >= min_pct * tot_s or 100 * i >= min_pct * tot_i:
                print("%5d %5.1f%% smp %5.1f%% inst | %s" % (ln, 100 * s / max(tot_s, 1), 100 * i / max(tot_i, 1), code.strip()[:120]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)


def regions(rep, kernel, spec):
    """spec: 'name:a-b,name:a-b' line ranges of the first file (the kernel's own source)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    files = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
    h = rows[hdrs[0]]
    cs, ci = h.index("# Samples"), h.index("Instructions Executed")
    end = files[1] if len(files) > 1 else len(rows)
    src = {}
    for r in rows[hdrs[0] + 1:end]:
        if len(r) > ci and r[0].isdigit():
            src[int(r[0])] = (int(r[cs] or 0), int(r[ci] or 0))
    ts, ti = sum(v[0] for v in src.values()), sum(v[1] for v in src.values())
    for part in spec.split(","):
        name, rng = part.split(":")
        a, b = map(int, rng.split("-"))
        s = sum(v[0] for k, v in src.items() if a <= k <= b)
        i = sum(v[1] for k, v in src.items() if a <= k <= b)
        print("%-14s lines %4d-%4d  %5.1f%% smp %5.1f%% inst  %12d warp-inst" % (name, a, b, 100 * s / ts, 100 * i / ti, i))
