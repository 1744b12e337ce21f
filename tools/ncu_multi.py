#!/usr/bin/env python3
"""Headline metrics of every kernel in an ncu report, one column per kernel.
  python tools/ncu_multi.py report.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
names = [r[col["Kernel Name"]].split("(")[0].replace("void ", "")[:22] for r in body]
print("%-66s" % "metric" + "".join("%24s" % n for n in names))
for k in KEYS:
    if k in col:
        print("%-66s" % (k + " [" + units[col[k]] + "]") + "".join("%24s" % r[col[k]][:22] for r in body))
st = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("\nstall reasons (cycles per issued instruction)")
tot = {h: sum(float(r[col[h]] or 0) for r in body) for h in st}
for h in sorted(st, key=lambda h: -tot[h])[:9]:
    print("%-66s" % h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") + "".join("%24.3f" % float(r[col[h]] or 0) for r in body))
