#!/usr/bin/env python3
"""Summarise an `ncu --set full --import-source on` capture of one kernel into a text file:
headline metrics, pipe utilisation, stall mix, opcode mix and a per-function attribution
(static SASS size, dynamic instructions, SIMT efficiency, stall samples).

  python tools/ncu_summary.py gpurun_out/prof_render.ncu-rep raytracinginrust_b200/lib/librtb200.so render_kernel > profiles/....txt
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, lib, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = dict(zip(hdr, zip(vals, units)))
    print("# ncu summary of %s (%s)" % (kernel, os.path.basename(rep)))
    print("\n## headline metrics")
    for k in KEYS:
        if k in m:
            print("%-70s %18s %s" % (k, m[k][0], m[k][1]))
    # achieved bandwidth per level against the B200 peaks (BASELINE: "achieved L2/HBM GB/s ... against B200 peak")
    def val(k):
        try:
            return float(m[k][0])
        except (KeyError, ValueError):
            return None
    dur_ns = val("gpu__time_duration.sum")
    if dur_ns is not None and m["gpu__time_duration.sum"][1] in ("us", "usecond"):
        dur_ns *= 1e3
    elif dur_ns is not None and m["gpu__time_duration.sum"][1] in ("ms", "msecond"):
        dur_ns *= 1e6
    elif dur_ns is not None and m["gpu__time_duration.sum"][1] in ("s", "second"):
        dur_ns *= 1e9
    print("\n## achieved bandwidth (whole launch)")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    if rd is not None and wr is not None and dur_ns:
        b = rd * scale.get(m["dram__bytes_read.sum"][1], 1.0) + wr * scale.get(m["dram__bytes_write.sum"][1], 1.0)
        peak = 6536.4  # MEASURED_PEAKS.json hbm_gbs of this pool's B200s
        print("HBM   %10.1f GB/s   %5.2f %% of the measured copy peak (%.0f GB/s); ncu: %s %% of its own peak" % (
            b / dur_ns, 100.0 * b / dur_ns / peak, peak, m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("?",))[0]))
    sec = val("lts__t_sectors.sum")
    if sec is not None and dur_ns:
        print("L2    %10.1f GB/s   (lts__t_sectors x 32 B; %s %% of ncu's L2 sector peak, lts__throughput %s %%)" % (
            sec * 32.0 / dur_ns, m.get("lts__t_sectors.sum.pct_of_peak_sustained_elapsed", ("?",))[0],
            m.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", ("?",))[0]))
    if "l1tex__throughput.avg.pct_of_peak_sustained_elapsed" in m:
        print("L1    l1tex__throughput %s %% of peak, hit rate %s %%" % (
            m["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"][0], m.get("l1tex__t_sector_hit_rate.pct", ("?",))[0]))
    print("FP32  pipe_fma %s %% (FFMA/FMUL/FADD), ALU %s %% (FMNMX, integer), FP64 %s %%, issue slots %s %%, SIMT %s of 32 threads" % (
        m.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", ("?",))[0],
        m.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", ("?",))[0],
        m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", ("?",))[0],
        m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", ("?",))[0],
        m.get("smsp__thread_inst_executed_per_inst_executed.ratio", ("?",))[0]))

    print("\n## warp stall reasons (cycles per issued instruction)")
    st = [(k, float(v[0])) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for k, v in sorted(st, key=lambda x: -x[1])[:8]:
        print("%-40s %8.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))

    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
    shdr = src[1]
    col = {h: k for k, h in enumerate(shdr)}
    body = src[2:]
    dyn, thr = collections.Counter(), collections.Counter()
    tot = 0
    for r in body:
        mm = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]].strip())
        op = mm.group(1) if mm else "?"
        n = int(r[col["Instructions Executed"]])
        dyn[op] += n
        thr[op] += int(r[col["Thread Instructions Executed"]])
        tot += n
    print("\n## dynamic opcode mix (warp instructions: %.4g; static SASS instructions: %d = %.0f KB)" % (tot, len(body), len(body) * 16 / 1024))
    for k, v in dyn.most_common(16):
        print("%-10s %5.1f%%   avg active threads %.1f" % (k, 100.0 * v / tot, thr[k] / max(v, 1)))

    # per-function attribution through nvdisasm line info (lib: .so or a comma-separated list of .o)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from ncu_lines import sass_lines
    order = sass_lines(lib, kernel, len(body))
    if order is None or len(order) != len(body):
        print("\n(per-function attribution skipped: SASS listing and capture differ: %s vs %d)" % (None if order is None else len(order), len(body)))
        return
    cache = {}

    def func_of(loc):
        if not loc:
            return "?"
        path, ln = loc
        if path not in cache:
            try:
                cache[path] = open(path).read().split("\n")
            except OSError:
                cache[path] = None
        s = cache[path]
        if s is None:
            return os.path.basename(path)
        for k in range(min(ln, len(s)) - 1, -1, -1):
            mm = re.match(r"^(?:RT_DEV|__global__|__device__)[^(]*?(\w+)\s*\(", s[k]) or re.match(r"^(\w+)\s*\(const __grid_constant__", s[k])
            if mm:
                return mm.group(1)
        return os.path.basename(path)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    tots = 0
    for loc, r in zip(order, body):
        a = agg[func_of(loc)]
        a[0] += 1
        a[1] += int(r[col["Instructions Executed"]])
        a[2] += int(r[col["Thread Instructions Executed"]])
        a[3] += int(r[col["# Samples"]])
        a[4] += int(r[col["stall_no_inst"]])
        tots += int(r[col["# Samples"]])
    print("\n## per-function attribution (inlined device functions by source line)")
    print("%-26s %7s %7s %8s %12s %9s" % ("function", "static", "dyn %", "samples%", "avg threads", "no_inst%"))
    for f, a in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
        print("%-26s %7d %6.1f%% %7.1f%% %12.1f %8.1f%%" % (f, a[0], 100.0 * a[1] / tot, 100.0 * a[3] / max(tots, 1), a[2] / max(a[1], 1), 100.0 * a[4] / max(a[3], 1)))


if __name__ == "__main__":
    main()
