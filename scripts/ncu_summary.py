#!/usr/bin/env python3
"""Condenses an `ncu --set full` report into what the repository keeps under profiles/:

  python scripts/ncu_summary.py <report.ncu-rep> <out prefix> [--roofline <kernel regex> <records per launch>]

  <prefix>_raw.csv      one row per captured launch, the metrics the summaries quote
  <prefix>_table.md     the same as a Markdown table (time, DRAM bytes and rate, occupancy, issue slots, top stall reasons)
  profiles/roofline_capture.json (with --roofline): dram__bytes_read/write.sum of the first launch matching the regex and
                        the number of records that launch processed -- bench.py scales it to the run's launch
"""
import csv
import json
import re
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
keep = [k for k in keep if k in idx]
stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def to_unit(v, u, want):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    return v * scale.get(u, 1.0)


with open(prefix + "_raw.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep + ["top stalls (cycles per issued instruction)"])
    w.writerow([units[idx[k]] for k in keep] + [""])
    table = []
    for r in rows[2:]:
        st = sorted([(num(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall], reverse=True)[:4]
        sts = ", ".join("%s %.2f" % (n, v) for v, n in st)
        w.writerow([r[idx[k]] for k in keep] + [sts])
        ms = to_unit(num(r[idx["gpu__time_duration.sum"]]), units[idx["gpu__time_duration.sum"]], "ms")
        rd = to_unit(num(r[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]], "byte")
        wr = to_unit(num(r[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]], "byte")
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("pg::", "").replace("<unnamed>::", "")
        table.append((name, ms, rd, wr, r[idx["launch__grid_size"]], r[idx["launch__block_size"]], r[idx["launch__registers_per_thread"]],
                      num(r[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]]), num(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]), sts))
with open(prefix + "_table.md", "w") as f:
    f.write("| kernel | ms | DRAM read GB | DRAM write GB | DRAM GB/s | grid x block | regs | warps active % | issue slots % | top stalls (cycles / issued instruction) |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for t in table:
        f.write("| `%s` | %.3f | %.3f | %.3f | %.0f | %s x %s | %s | %.1f | %.1f | %s |\n" % (t[0], t[1], t[2] / 1e9, t[3] / 1e9, (t[2] + t[3]) / 1e9 / (t[1] / 1e3) if t[1] else 0, t[4], t[5], t[6], t[7], t[8], t[9]))
if "--roofline" in sys.argv:
    i = sys.argv.index("--roofline")
    rx, nrec = re.compile(sys.argv[i + 1]), int(sys.argv[i + 2])
    for r in rows[2:]:
        if rx.search(r[idx["Kernel Name"]]):
            rd = to_unit(num(r[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]], "byte")
            wr = to_unit(num(r[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]], "byte")
            ms = to_unit(num(r[idx["gpu__time_duration.sum"]]), units[idx["gpu__time_duration.sum"]], "ms")
            json.dump({"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]), "source": rep.split("/")[-1] + " (ncu --set full --clock-control none)", "records": nrec,
                       "dram_bytes_read": rd, "dram_bytes_write": wr, "launch_ms_under_ncu": ms}, open("profiles/roofline_capture.json", "w"), indent=1)
            break
print("wrote", prefix + "_raw.csv", prefix + "_table.md")
