#!/usr/bin/env python3
"""bench.py — throughput of the path-tracing hot path on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full render of the workload: BASELINE.json configs[1], the Cornell box at
600x600, 1000 spp, max depth 100, HEAD integrator — the configuration the ">= 1 Grays/s per
B200" target is quoted on.  With N > 1 the SAME image is rendered (strong scaling): rank g
renders a contiguous block of the 1000 sample indices for all pixels, and the fp32 sum buffers
are combined by one NCCL reduce to rank 0 inside the timed region.

The JSON line (rank 0):
  value      Mpaths/s, whole job, scene resident in HBM, device time (CUDA events, max over ranks)
  e2e        Mpaths/s through the public entry with HOST buffers: scene description in host
             memory -> compile + upload -> kernels (-> reduce) -> fp32 sums back in host memory
  roofline   the render kernel against measured HBM copy bandwidth, with SURVEY §8(d)'s
             algorithmic bytes per segment (this path is NOT HBM-bound; see roofline_fp64)
  roofline_fp64  the bound that applies: FP64 issue, against a live-measured DFMA peak
  cpu_baseline   the f64 oracle (a restatement of the reference: no Rust toolchain here) on the
             box's host cores, on a bounded sample of the same workload
`--impl reference` times that CPU oracle as the reference arm (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, description)
    "cornell": "Cornell box (BASELINE configs[1]): 600x600, 1000 spp, depth 100, HEAD integrator",
    "cornell_smoke": "Cornell smoke (configs[2]): 600x600, 1000 spp",
    "random": "RTiOW random spheres (configs[0]): 500x500, 800 spp, legacy integrator",
    "final": "Next Week final scene (configs[3]): 800x800, 10000 spp",
    "mesh": "Triangle-mesh scene (configs[4]): 3840x2160, 1024 spp (Venus stand-in + teapot)",
}

# f64 bytes a test has to read (the reference's own parameters), SURVEY §8(d) restated for f64:
# AABB 6 doubles; sphere c+r; moving sphere c0,c1,t0,t1,r; rect a0,a1,b0,b1,k; triangle 3 vertices;
# translate 3 doubles / rotate sin,cos (one "xform" is one wrapper).
BYTES = {"box_tests": 48, "sphere_tests": 32, "msphere_tests": 72, "rect_tests": 40, "tri_tests": 72, "xform": 24,
         "medium_tests": 8}
# f64 flops per test, read off the reference's expressions (DESIGN.md "Rooflines")
FLOPS = {"box_tests": 18, "sphere_tests": 45, "msphere_tests": 60, "rect_tests": 14, "tri_tests": 60, "xform": 18,
         "medium_tests": 25}
FLOPS_SHADE = 160  # ONB + cosine/light sample + pdfs + throughput update, per segment


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons for one GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for ts, r in self.rows if (t_begin is None or ts >= t_begin) and (t_end is None or ts <= t_end + 0.15)]
        if not rows:  # window shorter than the sampling period: use everything collected while the GPU was busy
            rows = [r for _, r in self.rows]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # "under load": the upper half of the samples (idle gaps between steps pull the plain median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                # the whole distribution inside the timed window (idle gaps between steps included)
                "sm_mhz_min": sm[0] if sm else None, "sm_mhz_p10": sm[len(sm) // 10] if sm else None,
                "sm_mhz_median_all": sm[len(sm) // 2] if sm else None}


def algorithmic_work(counters):
    seg = max(counters["segments"], 1)
    per_seg = {k: counters[k] / seg for k in BYTES}
    bytes_per_seg = sum(BYTES[k] * per_seg[k] for k in BYTES)
    flops_per_seg = sum(FLOPS[k] * per_seg[k] for k in FLOPS) + FLOPS_SHADE
    return per_seg, bytes_per_seg, flops_per_seg


def cpu_oracle_run(scene_name, width, height, spp_sample, max_depth, integrator, threads=0):
    """The CPU arm: the oracle's sample loop (a restatement of src/main.rs:772-834)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    import raytracinginrust_b200 as rt
    hs = rt.HostScene(scene_name)
    osc = orc.OracleScene(hs.scene_desc)
    opts = rt.render_opts(seed=1, integrator=integrator, sample_begin=0, sample_count=spp_sample)
    t0 = time.perf_counter()
    _, rays, cnt = osc.render(hs.camera, width, height, spp_sample, max_depth, opts, threads=threads, counters=True)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "paths": width * height * spp_sample, "rays": rays, "counters": cnt,
            "threads": orc.num_threads()}


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port; the Rust crate cannot be built here)."""
    if rank != 0:
        return
    import raytracinginrust_b200 as rt
    hs = rt.HostScene(args.workload)
    spp_sample = args.ref_spp
    for _ in range(args.warmup):
        cpu_oracle_run(args.workload, hs.width, hs.height, max(1, spp_sample // 4), hs.max_depth, hs.integrator)
    tot_t, tot_paths, tot_rays, cores = 0.0, 0, 0, 1
    for _ in range(args.steps):
        r = cpu_oracle_run(args.workload, hs.width, hs.height, spp_sample, hs.max_depth, hs.integrator)
        tot_t += r["seconds"]
        tot_paths += r["paths"]
        tot_rays += r["rays"]
        cores = r["threads"]
    value = tot_paths / tot_t / 1e6
    sample = "%dx%d, %d of %d spp per step, all %d host threads (OpenMP over scanlines)" % (hs.width, hs.height, spp_sample, hs.spp, cores)
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mrays_per_s": tot_rays / tot_t / 1e6,
        "config": {"workload": WORKLOADS[args.workload], "scene": args.workload, "width": hs.width, "height": hs.height,
                   "spp": hs.spp, "max_depth": hs.max_depth, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU f64 restatement of the reference (oracle/oracle.cpp); cargo/rustc are absent so `cargo run --release` cannot be timed",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cornell", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp (diagnostics only)")
    ap.add_argument("--cpu-spp", type=int, default=400, help="spp of the bounded cpu_baseline sample")
    ap.add_argument("--ref-spp", type=int, default=200, help="spp per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))

    import numpy as np
    import torch
    import raytracinginrust_b200 as rt
    from raytracinginrust_b200.multi_gpu import sample_partition

    if rt.device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line (no "NCCL version" banner)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    hs = rt.HostScene(args.workload)
    W, H, depth = hs.width, hs.height, hs.max_depth
    spp = args.spp or hs.spp
    begin, count = sample_partition(spp, rank, world)
    opts = rt.render_opts(seed=1, integrator=hs.integrator, sample_begin=begin, sample_count=count)
    scene = rt.DeviceScene(hs.scene_desc, device=local_rank)
    out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if count > 0:
            scene.render_device(hs.camera, W, H, spp, depth, opts, out.data_ptr(), stream.cuda_stream)
        else:
            out.zero_()
        if dist is not None:
            dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)

    sampler = ClockSampler(local_rank) if rank == 0 else None  # started early: nvidia-smi takes ~1 s to come up
    for _ in range(args.warmup):
        flush.fill_(1)
        step()
        if count > 0:
            scene.render_wait()
    barrier()

    t_begin = time.time()
    step_ms, kern_ms, paths, rays, launches = [], [], 0, 0, 0
    for _ in range(args.steps):
        flush.fill_(0)  # L2 flush between timed steps
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        if count > 0:
            st = scene.render_wait()  # the library's own events around its two kernels, same stream
            kern_ms.append(st.render_ms)
            paths += st.paths
            rays += st.rays
            launches += st.kernel_launches
    barrier()
    t_end = time.time()
    if sampler and t_end - t_begin < 1.0:
        time.sleep(0.3)
    clocks = sampler.stop(t_begin, t_end) if sampler else None

    t = torch.tensor(step_ms, dtype=torch.float64, device="cuda")
    cnt = torch.tensor([paths, rays], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # per step: the slowest rank
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms = float(t.sum().item())
    total_paths, total_rays = float(cnt[0].item()), float(cnt[1].item())
    value = total_paths / (total_ms * 1e-3) / 1e6
    mrays = total_rays / (total_ms * 1e-3) / 1e6
    checksum = float(out.double().sum().item()) if rank == 0 else 0.0

    # ---- e2e: host scene description -> compile/upload -> render (-> reduce) -> host pixels ----
    host_out = torch.empty((H, W, 3), dtype=torch.float32).pin_memory() if rank == 0 else None
    e2e_ms = []
    h2d = d2h = 0
    for it in range(2):
        barrier()
        t0 = time.perf_counter()
        sc2 = rt.DeviceScene(hs.scene_desc, device=local_rank)  # flatten is already in hs; compile + H2D here
        if count > 0:
            sc2.render_device(hs.camera, W, H, spp, depth, opts, out.data_ptr(), stream.cuda_stream)
        else:
            out.zero_()
        if dist is not None:
            dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            host_out.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        h2d, d2h = sc2.device_bytes, W * H * 3 * 4
        if count > 0:
            sc2.render_wait()
        sc2.close()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if it > 0:
            e2e_ms.append(float(tt.item()))
    e2e_value = (W * H * spp) / (sum(e2e_ms) / len(e2e_ms) * 1e-3) / 1e6

    # ---- the whole `cargo run --release > image.ppm` job at N=1: scene graph -> flatten -> compile/upload ->
    # render -> format_color + P3 text on the GPU -> the file's bytes in host memory (rtb200_render's path) ----
    e2e_ppm = None
    if world == 1:
        what = ("host scene graph -> flatten -> rt_scene_group_create -> rt_render_multi -> rt_encode_ppm "
                "(format_color + P3 text on the GPU) -> the PPM file in host memory")
        # Measured in a child process that holds nothing but the library (tools/ppm_phase_probe.py: one warm-up run,
        # then three): inside this process the same calls take 2-3x longer for a 130 ms render - every run creates
        # and destroys its device scene, and next to torch's context the driver's allocation calls vary from 10 to a
        # few 100 ms (profiles/r1_h_whole_job_phases_cornell.txt: 150 ms wall for 126 ms on the device).
        try:
            child = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ppm_phase_probe.py"), args.workload,
                                    str(W), str(H), str(spp), str(depth), "3"], capture_output=True, text=True, timeout=600)
            res = json.loads(child.stdout.strip().splitlines()[-1])
            runs = sorted(float(x) for x in res["ms"])
            med = runs[len(runs) // 2]
            e2e_ppm = {"value": (W * H * spp) / (med * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": med,
                       "d2h_bytes_per_step": int(res["bytes"]), "first_call_ms": res.get("first_ms"),
                       "what": what + "; child process without torch, median of %d run(s) after one warm-up run" % len(runs)}
        except Exception as exc:  # the leg must not take the bench line down: fall back to this process
            sys.stderr.write("e2e_ppm child failed (%s); measuring in-process\n" % (exc,))
            ppm_ms, ppm_len = [], 0
            it = 0
            while True:  # one warm-up, then the median of three runs (of one when a run takes seconds)
                t0 = time.perf_counter()
                ppm, _ = hs.render_ppm(W, H, spp, depth, opts, n_gpus=1)
                dt = (time.perf_counter() - t0) * 1e3
                ppm_len = len(ppm)
                if it > 0:
                    ppm_ms.append(dt)
                it += 1
                if it >= 4 or (it >= 2 and dt > 2000.0):
                    break
            ppm_ms.sort()
            med = ppm_ms[len(ppm_ms) // 2]
            e2e_ppm = {"value": (W * H * spp) / (med * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": med,
                       "d2h_bytes_per_step": ppm_len, "what": what + "; in-process, median of %d run(s)" % len(ppm_ms)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample + the algorithmic work per segment (oracle counters) ----
    cpu = None
    per_seg, bytes_seg, flops_seg = {}, None, None
    if not args.no_cpu_baseline:
        r = cpu_oracle_run(args.workload, W, H, args.cpu_spp, depth, hs.integrator)
        per_seg, bytes_seg, flops_seg = algorithmic_work(r["counters"])
        cpu = {"value": r["paths"] / r["seconds"] / 1e6, "unit": "Mpaths/s", "cores": r["threads"], "kind": "port",
               "mrays_per_s": r["rays"] / r["seconds"] / 1e6,
               "sample": "%dx%d, %d of %d spp, %.1f s, all %d host threads (OpenMP over scanlines)"
                         % (W, H, args.cpu_spp, spp, r["seconds"], r["threads"])}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    kern_s = (sum(kern_ms) / len(kern_ms)) * 1e-3 if kern_ms else None
    segs_per_launch = (rays / len(kern_ms)) if kern_ms else 0.0
    info = scene.render_info  # which pipeline build ran: megakernel or wavefront, feature variant
    pipeline = info.get("pipeline", "?")
    kernel_name = "render_kernel" if pipeline == "megakernel" else "wf_extend_kernel + wf_shade_kernel + wf_generate_kernel"
    roofline = roofline64 = None
    traffic = None  # dram__bytes_read+write of render_kernel per launch, from the committed ncu capture of this command
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["render_kernel"]
        if tr["scene"] == args.workload and tr["spp"] == spp and tr["n_gpus"] == world:
            traffic = tr["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    if kern_s and bytes_seg is not None:
        achieved = bytes_seg * segs_per_launch / kern_s / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "peak_source": hbm_src, "kernel": kernel_name,
                    "algorithmic_bytes_per_segment": bytes_seg, "segments_per_launch": segs_per_launch,
                    "kernel_ms": kern_s * 1e3,
                    "note": ("megakernel: no ray queues (Q=0); the scene tables are cache-resident, so measured DRAM "
                             "traffic is far BELOW the algorithmic bytes; the binding limit is FP64 issue (roofline_fp64)")
                    if pipeline == "megakernel" else
                            ("wavefront: the algorithmic bytes exclude the path-pool traffic (one 128-byte slot record "
                             "read+written per segment and stage); the binding limit is FP64 issue (roofline_fp64)")}
        fp64_peak = rt.measure_fp64_peak(local_rank)
        ach_tf = flops_seg * segs_per_launch / kern_s / 1e12
        roofline64 = {"bound": "fp64_issue", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                      "frac": ach_tf / fp64_peak, "peak_source": "rt_measure_fp64_peak (DFMA micro-kernel, this GPU, this run)",
                      "algorithmic_flops_per_segment": flops_seg}

    line = {
        "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mrays_per_s": mrays, "segments_per_path": total_rays / max(total_paths, 1.0),
        "config": {"workload": WORKLOADS[args.workload], "scene": args.workload, "width": W, "height": H, "spp": spp,
                   "max_depth": depth, "integrator": "HEAD" if hs.integrator == 0 else "LEGACY",
                   "parallelism": "sample blocks x%d + 1 ncclReduce" % world if world > 1 else "1 GPU",
                   "l2": "256 MiB device memset between timed steps (L2 flush)", "seed": 1,
                   "pipeline": pipeline, "variant": info.get("variant"), "pipeline_info": info},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": sum(e2e_ms) / len(e2e_ms),
                "what": "RtSceneDesc in host memory -> rt_scene_create (compile, BVH build, upload) -> render -> "
                        "fp32 sums copied to pinned host memory"},
        "e2e_ppm": e2e_ppm,
        "gpu_launches": int(launches),  # this rank's render + reduce kernels inside the timed steps (RtStats.kernel_launches)
        "roofline": roofline, "roofline_fp64": roofline64, "cpu_baseline": cpu,
        "algorithmic_tests_per_segment": per_seg, "checksum": checksum,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
