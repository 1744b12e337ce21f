#!/usr/bin/env python3
"""profiles/ncu_traffic.json from an `ncu --set full` capture of the render kernel under the bench command:
measured dram__bytes_read + dram__bytes_write of ONE launch, with the configuration it belongs to
(bench.py only uses it when scene, spp, GPU count and chunk count equal its own run).

  python tools/ncu_traffic.py gpurun_out/b_render_kernel_bench.ncu-rep cornell 1000 1 <chunks> "<source note>"
"""
import csv
import io
import json
import os
import subprocess
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, scene, spp, n_gpus, chunks, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                                                     stderr=subprocess.DEVNULL, text=True).stdout)))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = dict(zip(hdr, zip(vals, units)))

    def b(k):
        return float(m[k][0]) * SCALE.get(m[k][1], 1.0)
    rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
    out = {"render_kernel": {"scene": scene, "spp": spp, "n_gpus": n_gpus, "chunks": chunks,
                             "dram_bytes_per_launch": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                             "kernel": m["Kernel Name"][0] if "Kernel Name" in m else "render_kernel",
                             "duration_under_ncu_ms": float(m["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(
                                 m["gpu__time_duration.sum"][1].replace("second", "s").replace("nsecond", "ns"), 1.0),
                             "source": note}}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    json.dump(out, open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=2)
    print(json.dumps(out, indent=2))


if __name__ == "__main__":
    main()
