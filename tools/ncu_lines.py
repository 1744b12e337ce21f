#!/usr/bin/env python3
"""Per-source-line and per-function attribution of one kernel of an ncu report (needs the .so the
capture ran, built with -lineinfo).
  python tools/ncu_lines.py report.ncu-rep lib.so kernel_regex [top_n]"""
import collections, csv, io, os, re, subprocess, sys, tempfile


def run(cmd, cwd=None):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=cwd).stdout


def sass_lines(lib, kernel_sub, n_expected):
    """(file, line) of every SASS instruction of the kernel section whose size matches the capture."""
    # `lib` may be a comma-separated list (.so or .o): objects compiled from the same source give
    # cubins of the same name, so each file is extracted into a directory of its own
    cubins = []
    for one in lib.split(","):
        tmp = tempfile.mkdtemp()
        run(["cuobjdump", "-xelf", "all", os.path.abspath(one)], cwd=tmp)
        cubins += [os.path.join(tmp, f) for f in sorted(os.listdir(tmp)) if f.endswith(".cubin")]
    best = None
    for f in cubins:
        lines = run(["nvdisasm", "-gi", "-c", f]).split("\n")
        starts = [i for i, l in enumerate(lines) if l.strip().startswith(".section") and ".text." in l and kernel_sub in l]
        for start in starts:
            order, cur, open_ = [], None, False
            for l in lines[start + 1:]:
                if l.strip().startswith(".section"):
                    break
                mm = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
                if mm:
                    if not open_:  # the first marker before an instruction is the innermost (inlined) location
                        cur, open_ = (mm.group(1), int(mm.group(2))), True
                elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S+", l):
                    order.append(cur)
                    open_ = False
            if len(order) == n_expected:
                return order
            best = order
    return best


def main():
    rep, lib, kre = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre]))))
    kname = src[0][1]
    col = {h: k for k, h in enumerate(src[1])}
    body = [r for r in src[2:] if len(r) > 10 and r[0].startswith("0x")]
    # several launches of the same kernel are concatenated: keep the first
    first_addr = body[0][0]
    for k in range(1, len(body)):
        if body[k][0] == first_addr:
            body = body[:k]
            break
    mangled = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
    order = sass_lines(lib, mangled, len(body))
    print("# %s: %d SASS instructions" % (kname.split("(")[0], len(body)))
    if order is None or len(order) != len(body):
        print("SASS listing and capture differ: %s vs %d" % (None if order is None else len(order), len(body)))
        return
    tot_i = sum(int(r[col["Instructions Executed"]]) for r in body)
    tot_s = sum(int(r[col["# Samples"]]) for r in body)
    cache = {}

    def text(loc):
        path, ln = loc
        if path not in cache:
            try:
                cache[path] = open(path).read().split("\n")
            except OSError:
                cache[path] = None
        s = cache[path]
        return s[ln - 1].strip()[:90] if s and ln <= len(s) else ""

    def func_of(loc):
        path, ln = loc
        text(loc)
        s = cache[path]
        if s is None:
            return os.path.basename(path)
        for k in range(min(ln, len(s)) - 1, -1, -1):
            mm = re.match(r"^(?:RT_DEV_COLD|RT_DEV|__global__|__device__|template)[^(]*?(\w+)\s*\(", s[k]) or re.match(r"^(\w+)\s*\(const __grid_constant__", s[k])
            if mm:
                return mm.group(1)
        return os.path.basename(path)

    per_line = collections.defaultdict(lambda: [0, 0, 0, 0])
    per_fn = collections.defaultdict(lambda: [0, 0, 0, 0])
    for loc, r in zip(order, body):
        loc = loc or ("?", 0)
        vals = (1, int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]), int(r[col["# Samples"]]))
        for agg, key in ((per_line, (os.path.basename(loc[0]), loc[1])), (per_fn, func_of(loc) if loc[0] != "?" else "?")):
            a = agg[key]
            for k in range(4):
                a[k] += vals[k]
    print("total warp instructions %.4g, avg threads %.2f" % (tot_i, sum(a[2] for a in per_fn.values()) / max(tot_i, 1)))
    print("\n## per function\n%-26s %7s %7s %8s %11s" % ("function", "static", "dyn %", "samples%", "avg threads"))
    for f, a in sorted(per_fn.items(), key=lambda x: -x[1][1])[:30]:
        print("%-26s %7d %6.1f%% %7.1f%% %11.1f" % (f, a[0], 100.0 * a[1] / tot_i, 100.0 * a[3] / max(tot_s, 1), a[2] / max(a[1], 1)))
    print("\n## per source line\n%-22s %6s %8s %7s  %s" % ("file:line", "dyn %", "samples%", "threads", "source"))
    files = {os.path.basename(l[0]): l[0] for l in order if l}
    for (f, ln), a in sorted(per_line.items(), key=lambda x: -x[1][1])[:top]:
        print("%-22s %5.1f%% %7.1f%% %7.1f  %s" % ("%s:%d" % (f, ln), 100.0 * a[1] / tot_i, 100.0 * a[3] / max(tot_s, 1), a[2] / max(a[1], 1), text((files.get(f, f), ln))))


if __name__ == "__main__":
    main()
