#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import io
import subprocess
import sys


def main(rep, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur = None
    agg = {}
    inst = {}
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 6 and r[0].isdigit() and r[2] == "-":
            key = (cur, int(r[0]), r[1].strip()[:110])
            agg[key] = agg.get(key, 0) + (int(r[4]) if r[4].isdigit() else 0)
            inst[key] = inst.get(key, 0) + (int(r[7]) if len(r) > 7 and r[7].isdigit() else 0)
    tot = sum(agg.values()) or 1
    ti = sum(inst.values()) or 1
    print(f"total samples {tot}, warp instructions {ti}")
    for (f, l, src), s in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * s / tot, 100 * inst[(f, l, src)] / ti, f, l, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
