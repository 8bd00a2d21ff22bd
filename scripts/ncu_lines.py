#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (compile with -lineinfo).

  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep regex:extend_query_warp [top]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
fname = ""
data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            data.append((fname, int(r[0]), r[1], int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]),
                         int(r[hdr.index("stall_long_sb")]), int(r[hdr.index("stall_short_sb")]), int(r[hdr.index("stall_wait")]),
                         int(r[hdr.index("stall_math")]), int(r[hdr.index("stall_barrier")])))
        except ValueError:
            pass
ts = sum(d[3] for d in data) or 1
ti = sum(d[4] for d in data) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
print("%6s %6s | long short wait math barr | line" % ("samp%", "inst%"))
for d in sorted(data, key=lambda x: -x[3])[:top]:
    print("%5.1f%% %5.1f%% | %4d %4d %4d %4d %4d | %s:%d  %s" % (100.0 * d[3] / ts, 100.0 * d[4] / ti, d[5], d[6], d[7], d[8], d[9], d[0], d[1], d[2].strip()[:110]))
