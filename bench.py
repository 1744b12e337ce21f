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
  e2e        Mpaths/s through the C ABI with HOST buffers: RtSceneDesc in host memory ->
             rt_scene_create (compile + upload) -> kernels (-> reduce) -> fp32 sums back in pinned host
             memory; median of >= 5 timed iterations
  roofline   the BINDING bound of this path: FP64 issue.  achieved = algorithmic f64 flops per segment
             (the reference's own tests, counted by the oracle on the same workload) x segments per
             launch / the render kernel's time; peak = a DFMA micro-kernel run in this process
  roofline_hbm  the same kernel against measured HBM copy bandwidth (SURVEY §8(d)'s byte model) with the
             MEASURED dram bytes of an ncu capture of this build (profiles/ncu_traffic.json): this path
             is not HBM-bound, the entry is there to show it
  workloads  the other four configs (RTiOW and Cornell smoke at full spp, the Next Week final scene and the
             triangle-mesh scene - the ones the scaling targets name - at reduced spp): device and e2e
             numbers measured the same way
  cpu_baseline   the f64 oracle (a restatement of the reference: no Rust toolchain here) on the
             box's host cores, on a bounded sample of the same workload (N = 1 only)
`--impl reference` times that CPU oracle as the reference arm (rank 0 only) in a process that never
loads the CUDA library.
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
    # name: description
    "cornell": "Cornell box (BASELINE configs[1]): 600x600, 1000 spp, depth 100, HEAD integrator",
    "cornell_smoke": "Cornell smoke (configs[2]): 600x600, 1000 spp",
    "cornell_smoke_legacy": "Cornell smoke (configs[2]) under the legacy integrator (the smoke scatters: main.rs:82-84): 600x600, 1000 spp",
    "random": "RTiOW random spheres (configs[0]): 500x500, 800 spp, legacy integrator",
    "final": "Next Week final scene (configs[3]): 800x800, 10000 spp",
    "mesh": "Triangle-mesh scene (configs[4]): 3840x2160, 1024 spp (Venus stand-in + teapot)",
}
# The other four configs of BASELINE.json, measured next to the headline: configs[0] and [2] at their full spp, the
# two the scaling targets name ([3], [4]) at a bounded spp (throughput is linear in spp; image size, depth and scene
# are the config's own).  The mesh scene's `e2e` contains one scene compile (~0.25 s for its 394 k triangles) per step:
# at 128 of the config's 1024 spp that is a fifth of the step, at 32 spp (r2-b .. r2-p) it was half of it.
EXTRA_WORKLOADS = {"final": 2000, "cornell_smoke": 1000, "cornell_smoke_legacy": 1000, "random": 800, "mesh": 128}
# a workload that is a scene under another integrator than its own: name -> (scene, integrator)
VARIANT_OF = {"cornell_smoke_legacy": ("cornell_smoke", 1)}  # "also report the legacy integrator"

# f64 bytes a test has to read (the reference's own parameters), SURVEY §8(d) restated for f64:
# AABB 6 doubles; sphere c+r; moving sphere c0,c1,t0,t1,r; rect a0,a1,b0,b1,k; triangle 3 vertices;
# translate 3 doubles / rotate sin,cos (one "xform" is one wrapper).
BYTES = {"box_tests": 48, "sphere_tests": 32, "msphere_tests": 72, "rect_tests": 40, "tri_tests": 72, "xform": 24,
         "medium_tests": 8}
# f64 flops per test, read off the reference's expressions (DESIGN.md "Rooflines")
FLOPS = {"box_tests": 18, "sphere_tests": 45, "msphere_tests": 60, "rect_tests": 14, "tri_tests": 60, "xform": 18,
         "medium_tests": 25}
FLOPS_SHADE = 160  # ONB + cosine/light sample + pdfs + throughput update, per segment


def host_threads():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask the scheduler instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


# ---------------------------------------------------------------------------------------------------------
# The CPU arm.  It needs scene descriptions (librtb200_scenes.so, no CUDA dependency) and the oracle; it must
# not load the product's CUDA library.  `import raytracinginrust_b200` would (its __init__ loads librtb200.so
# and fails loudly without it), so the CPU arm registers a bare package object that only gives the submodules
# _abi / _scenes a home, unless the real package is already imported (the cpu_baseline leg of our own arm).
# ---------------------------------------------------------------------------------------------------------
def scenes_module():
    if "raytracinginrust_b200" not in sys.modules:
        import types
        pkg = types.ModuleType("raytracinginrust_b200")
        pkg.__path__ = [os.path.join(ROOT, "raytracinginrust_b200")]
        sys.modules["raytracinginrust_b200"] = pkg
    import raytracinginrust_b200._scenes as scenes
    return scenes


_ORACLE = None


def oracle_module():
    """oracle_py on its -O3 -march=native build, compiled on this machine (BASELINE.md's flags + -ffp-contract=off)."""
    global _ORACLE
    if _ORACLE is None:
        scenes_module()
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as orc
        try:
            orc.use_native_build()
            orc.BUILD_KIND = "-O3 -march=native -ffp-contract=off, built on this host"
        except (OSError, subprocess.CalledProcessError) as exc:  # no compiler on the box: the generic -O2 build
            sys.stderr.write("oracle native build failed (%s); timing the generic -O2 build\n" % (exc,))
            orc.BUILD_KIND = "-O2 -ffp-contract=off (generic prebuilt)"
        _ORACLE = orc
    return _ORACLE


def cpu_oracle_run(scene_name, width, height, spp_sample, max_depth, integrator, threads=None):
    """The CPU arm: the oracle's sample loop (a restatement of src/main.rs:772-834) on `threads` host threads."""
    scenes = scenes_module()
    orc = oracle_module()
    hs = scenes.HostScene(scene_name)
    osc = orc.OracleScene(hs.scene_desc)
    opts = scenes.render_opts(seed=1, integrator=integrator, sample_begin=0, sample_count=spp_sample)
    threads = threads or host_threads()
    t0 = time.perf_counter()
    _, rays, cnt = osc.render(hs.camera, width, height, spp_sample, max_depth, opts, threads=threads, counters=True)
    dt = time.perf_counter() - t0
    osc.close()
    return {"seconds": dt, "paths": width * height * spp_sample, "rays": rays, "counters": cnt, "threads": threads,
            "build": orc.BUILD_KIND}


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port; the Rust crate cannot be built here)."""
    if rank != 0:
        return
    scenes = scenes_module()
    hs = scenes.HostScene(args.workload)
    W, H, depth, integ, spp_full = hs.width, hs.height, hs.max_depth, hs.integrator, hs.spp
    del hs
    spp_sample = args.ref_spp
    for _ in range(min(args.warmup, 1)):  # one short warm-up pass (page-in, OpenMP pool); the CPU has no clocks to ramp
        cpu_oracle_run(args.workload, W, H, max(1, spp_sample // 8), depth, integ)
    tot_t, tot_paths, tot_rays, cores, build = 0.0, 0, 0, 1, ""
    budget_s, done = 240.0, 0
    for _ in range(args.steps):
        r = cpu_oracle_run(args.workload, W, H, spp_sample, depth, integ)
        tot_t += r["seconds"]
        tot_paths += r["paths"]
        tot_rays += r["rays"]
        cores, build = r["threads"], r["build"]
        done += 1
        if tot_t + r["seconds"] > budget_s:  # keep the whole arm within a few minutes whatever K the driver passes
            break
    value = tot_paths / tot_t / 1e6
    sample = ("%dx%d, %d of %d spp per step, %d of %d steps timed, all %d host threads (OpenMP over scanlines), oracle built %s"
              % (W, H, spp_sample, spp_full, done, args.steps, cores, build))
    assert not any("librtb200.so" in l for l in open("/proc/self/maps")), "the reference arm loaded the CUDA library"
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / done * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mrays_per_s": tot_rays / tot_t / 1e6,
        "config": {"workload": WORKLOADS[args.workload], "scene": args.workload, "width": W, "height": H,
                   "spp": spp_full, "max_depth": depth, "integrator": "HEAD" if integ == 0 else "LEGACY", "seed": 1,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU f64 restatement of the reference (oracle/oracle.cpp); cargo/rustc are absent so `cargo run --release` "
                "cannot be timed; this process holds librtb200_scenes.so and the oracle only (no CUDA library)",
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------------------------------------
class Job:
    """One workload on this rank's GPU: the device-timed steps and the end-to-end iterations."""

    def __init__(self, rt, torch, dist, name, spp, rank, local_rank, world):
        from raytracinginrust_b200.multi_gpu import sample_partition
        self.rt, self.torch, self.dist, self.name = rt, torch, dist, name
        self.rank, self.local_rank, self.world = rank, local_rank, world
        scene_name, integrator = VARIANT_OF.get(name, (name, None))
        self.hs = rt.HostScene(scene_name)
        self.integrator = self.hs.integrator if integrator is None else integrator
        self.W, self.H, self.depth = self.hs.width, self.hs.height, self.hs.max_depth
        self.spp = spp or self.hs.spp
        self.begin, self.count = sample_partition(self.spp, rank, world)
        self.opts = rt.render_opts(seed=1, integrator=self.integrator, sample_begin=self.begin, sample_count=self.count)
        self.out = torch.zeros((self.H, self.W, 3), dtype=torch.float32, device="cuda")
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def step(self, scene):
        if self.count > 0:
            scene.render_device(self.hs.camera, self.W, self.H, self.spp, self.depth, self.opts, self.out.data_ptr(),
                                self.stream.cuda_stream)
        else:
            self.out.zero_()
        if self.dist is not None:
            self.dist.reduce(self.out, dst=0, op=self.dist.ReduceOp.SUM)

    def device_timed(self, steps, warmup, flush):
        """W warm-up steps, then K steps each bracketed by barrier + synchronize, CUDA events on the launching stream,
        an L2 flush (256 MiB memset) before each; per step the max over ranks."""
        torch, dist = self.torch, self.dist
        scene = self.rt.DeviceScene(self.hs.scene_desc, device=self.local_rank)
        for _ in range(warmup):
            flush.fill_(1)
            self.step(scene)
            if self.count > 0:
                scene.render_wait()
        self.barrier()
        t_begin = time.time()
        step_ms, kern_ms, paths, rays, launches = [], [], 0, 0, 0
        for _ in range(steps):
            flush.fill_(0)
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            self.step(scene)
            e1.record(self.stream)
            torch.cuda.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            if self.count > 0:
                st = scene.render_wait()  # the library's own events around its kernels, same stream
                kern_ms.append(st.render_ms)
                paths += st.paths
                rays += st.rays
                launches += st.kernel_launches
        self.barrier()
        t_end = time.time()
        t = torch.tensor(step_ms, dtype=torch.float64, device="cuda")
        cnt = torch.tensor([paths, rays], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # per step: the slowest rank
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_ms = float(t.sum().item())
        res = {"total_ms": total_ms, "paths": float(cnt[0].item()), "rays": float(cnt[1].item()), "steps": steps,
               "kern_ms": kern_ms, "rank_rays": rays, "launches": launches, "t_begin": t_begin, "t_end": t_end,
               "info": scene.render_info, "checksum": float(self.out.double().sum().item()) if self.rank == 0 else 0.0}
        res["value"] = res["paths"] / (total_ms * 1e-3) / 1e6
        res["mrays"] = res["rays"] / (total_ms * 1e-3) / 1e6
        scene.close()
        return res

    def e2e(self, iters):
        """RtSceneDesc in host memory -> rt_scene_create (compile, BVH build, upload) -> render (-> reduce) -> fp32 sums
        in pinned host memory, wall clock, max over ranks per iteration; one untimed iteration, then the median."""
        torch, dist, rt = self.torch, self.dist, self.rt
        from raytracinginrust_b200.multi_gpu import create_scene_distributed
        host_out = torch.empty((self.H, self.W, 3), dtype=torch.float32).pin_memory() if self.rank == 0 else None
        ms, h2d, d2h = [], 0, 0
        for it in range(iters + 1):
            self.barrier()
            t0 = time.perf_counter()
            if dist is not None:  # one compile on rank 0, the blob over ncclBroadcast, one upload per rank
                sc, tables_hash = create_scene_distributed(self.hs.scene_desc, self.rank, self.local_rank, device="cuda")
            else:
                sc, tables_hash = rt.DeviceScene(self.hs.scene_desc, device=self.local_rank), None
            self.step(sc)
            if self.rank == 0:
                host_out.copy_(self.out, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            h2d, d2h = sc.device_bytes, self.W * self.H * 3 * 4
            if self.count > 0:
                sc.render_wait()
            sc.close()
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if it > 0:
                ms.append(float(tt.item()))
        ms.sort()
        med = ms[len(ms) // 2]
        return {"value": (self.W * self.H * self.spp) / (med * 1e-3) / 1e6, "unit": "Mpaths/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": med,
                "iterations": len(ms), "ms_min": ms[0], "ms_max": ms[-1], "tables_hash": tables_hash,
                "what": ("RtSceneDesc in host memory -> rt_scene_create (compile, BVH build, upload) -> render -> "
                         "fp32 sums copied to pinned host memory" if dist is None else
                         "RtSceneDesc in host memory -> rt_compile on rank 0 -> ncclBroadcast of the table blob -> "
                         "rt_scene_create_compiled on every rank -> render -> ncclReduce -> fp32 sums in pinned host memory")
                        + "; median of %d iterations after one untimed" % len(ms)}


def fp64_roofline(flops_seg, rays_per_launch, kern_s, fp64_peak, kernel_name):
    ach_tf = flops_seg * rays_per_launch / kern_s / 1e12
    return {"bound": "fp64_issue", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
            "traffic": None, "kernel": kernel_name, "kernel_ms": kern_s * 1e3,
            "peak_source": "rt_measure_fp64_peak (DFMA micro-kernel, this GPU, this run; MEASURED_PEAKS.json has no FP64 entry)",
            "algorithmic_flops_per_segment": flops_seg, "segments_per_launch": rays_per_launch}


def kernel_names(pipeline):
    return "render_kernel" if pipeline == "megakernel" else "wf_extend_kernel + wf_shade_kernel + wf_generate_kernel"


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries ONE line, the JSON.  Everything else a library writes to file descriptor 1 while the bench runs -
    NCCL's version banner under torchrun at 8 GPUs, whatever NCCL_DEBUG_FILE says - goes to stderr: fd 1 is pointed at
    fd 2 for the run and the JSON line is written to the descriptor stdout had."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cornell", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp (diagnostics only)")
    ap.add_argument("--cpu-spp", type=int, default=400, help="spp of the bounded cpu_baseline sample")
    ap.add_argument("--ref-spp", type=int, default=200, help="spp per step of the --impl reference arm")
    ap.add_argument("--e2e-iters", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true")
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

    import torch
    import raytracinginrust_b200 as rt

    if rt.device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner and warnings: not on stdout, which holds the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    job = Job(rt, torch, dist, args.workload, args.spp, rank, local_rank, world)
    W, H, depth, spp = job.W, job.H, job.depth, job.spp

    sampler = ClockSampler(local_rank) if rank == 0 else None  # started early: nvidia-smi takes ~1 s to come up
    dev = job.device_timed(args.steps, args.warmup, flush)
    if sampler and dev["t_end"] - dev["t_begin"] < 1.0:
        time.sleep(0.3)
    clocks = sampler.stop(dev["t_begin"], dev["t_end"]) if sampler else None
    e2e = job.e2e(max(args.e2e_iters, 5))

    # ---- the two configs the scaling targets name, at a bounded spp: same measurement, fixed step counts ----
    extras = {}
    if not args.no_extra_workloads and args.workload == "cornell" and not args.spp:
        for name, xspp in EXTRA_WORKLOADS.items():
            xj = Job(rt, torch, dist, name, xspp, rank, local_rank, world)
            xd = xj.device_timed(3, 3, flush)
            xe = xj.e2e(3)
            info = dict(kv.split("=", 1) for kv in xd["info"].split() if "=" in kv) if isinstance(xd["info"], str) else xd["info"]
            extras[name] = {"config": {"workload": WORKLOADS[name], "scene": VARIANT_OF.get(name, (name, None))[0],
                                       "integrator": "HEAD" if xj.integrator == 0 else "LEGACY", "width": xj.W, "height": xj.H,
                                       "spp": xj.spp, "spp_of_config": xj.hs.spp, "max_depth": xj.depth,
                                       "pipeline": info.get("pipeline"), "variant": info.get("variant")},
                            "value": xd["value"], "unit": "Mpaths/s", "mrays_per_s": xd["mrays"],
                            "ms_per_step": xd["total_ms"] / xd["steps"], "steps": xd["steps"], "warmup": 3,
                            "e2e": xe, "gpu_launches": int(xd["launches"]), "checksum": xd["checksum"],
                            "_kern_ms": xd["kern_ms"], "_rank_rays": xd["rank_rays"], "_job": (xj.W, xj.H, xj.depth, xj.integrator)}
            del xj

    # ---- the whole `cargo run --release > image.ppm` job at N=1: scene graph -> flatten -> compile/upload ->
    # render -> format_color + P3 text on the GPU -> the file's bytes in host memory (rtb200_render's path) ----
    e2e_ppm = None
    if world == 1:
        what = ("host scene graph -> flatten -> rt_scene_group_create -> rt_render_multi -> rt_encode_ppm "
                "(format_color + P3 text on the GPU) -> the PPM file in host memory")
        # Measured in a child process that holds nothing but the library (tools/ppm_phase_probe.py: one warm-up run,
        # then five): inside this process the same calls take 2-3x longer for a 130 ms render - every run creates
        # and destroys its device scene, and next to torch's context the driver's allocation calls vary from 10 to a
        # few 100 ms (profiles/r1_h_whole_job_phases_cornell.txt: 150 ms wall for 126 ms on the device).
        try:
            child = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ppm_phase_probe.py"), args.workload,
                                    str(W), str(H), str(spp), str(depth), "5"], capture_output=True, text=True, timeout=600)
            res = json.loads(child.stdout.strip().splitlines()[-1])
            runs = sorted(float(x) for x in res["ms"])
            med = runs[len(runs) // 2]
            e2e_ppm = {"value": (W * H * spp) / (med * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": med,
                       "d2h_bytes_per_step": int(res["bytes"]), "first_call_ms": res.get("first_ms"),
                       "what": what + "; child process without torch, median of %d run(s) after one warm-up run" % len(runs)}
        except Exception as exc:  # the leg must not take the bench line down
            sys.stderr.write("e2e_ppm child failed (%s)\n" % (exc,))

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample (N = 1 only) + the algorithmic work per segment (oracle counters) ----
    # Under torchrun the other ranks wait in the closing barrier while rank 0 is here, so at N > 1 the oracle only
    # counts the tests of a small sample (a fraction of a second); the reference arm is the driver's own run.
    cpu = None
    per_seg, bytes_seg, flops_seg = {}, None, None
    hs_integrator = job.hs.integrator
    if not args.no_cpu_baseline:
        cpu_spp = args.cpu_spp if world == 1 else 8
        r = cpu_oracle_run(args.workload, W, H, cpu_spp, depth, hs_integrator)
        per_seg, bytes_seg, flops_seg = algorithmic_work(r["counters"])
        if world == 1:
            cpu = {"value": r["paths"] / r["seconds"] / 1e6, "unit": "Mpaths/s", "cores": r["threads"], "kind": "port",
                   "mrays_per_s": r["rays"] / r["seconds"] / 1e6,
                   "sample": "%dx%d, %d of %d spp, %.1f s, all %d host threads (OpenMP over scanlines), oracle built %s"
                             % (W, H, cpu_spp, spp, r["seconds"], r["threads"], r["build"])}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    kern_ms = dev["kern_ms"]
    kern_s = (sum(kern_ms) / len(kern_ms)) * 1e-3 if kern_ms else None
    segs_per_launch = (dev["rank_rays"] / len(kern_ms)) if kern_ms else 0.0
    info = dict(kv.split("=", 1) for kv in dev["info"].split() if "=" in kv) if isinstance(dev["info"], str) else dev["info"]
    pipeline = info.get("pipeline", "?")
    kernel_name = kernel_names(pipeline)
    roofline = roofline_hbm = None
    fp64_peak = rt.measure_fp64_peak(local_rank)
    # dram__bytes_read+write of the render kernel per launch, from the committed ncu capture of this command; only
    # taken when the capture is of the same workload, spp, GPU count and plane (chunk) count as this run
    traffic, traffic_note = None, "no matching ncu capture in profiles/ncu_traffic.json"
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["render_kernel"]
        same = (tr["scene"] == args.workload and tr["spp"] == spp and tr["n_gpus"] == world
                and str(tr.get("chunks")) == str(info.get("chunks")))
        if same:
            traffic, traffic_note = tr["dram_bytes_per_launch"], tr.get("source", "")
        else:
            traffic_note = "profiles/ncu_traffic.json is of another configuration (scene/spp/gpus/chunks): not used"
    except (OSError, ValueError, KeyError):
        pass
    if kern_s and flops_seg is not None:
        roofline = fp64_roofline(flops_seg, segs_per_launch, kern_s, fp64_peak, kernel_name)
        achieved = bytes_seg * segs_per_launch / kern_s / 1e9
        roofline_hbm = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "traffic": traffic, "traffic_source": traffic_note,
                        "traffic_frac_of_peak": (traffic / kern_s / 1e9 / hbm_peak) if traffic else None,
                        "peak_source": hbm_src, "kernel": kernel_name,
                        "algorithmic_bytes_per_segment": bytes_seg, "segments_per_launch": segs_per_launch,
                        "note": "NOT the binding bound: `achieved` counts the f64 operands of the reference's tests as if every "
                                "test re-read them from HBM; the tables are cache-resident and the measured DRAM traffic "
                                "(`traffic`, mostly the f64 sample planes) is what actually moves"}
        roofline["traffic"] = traffic

    # FP64-issue fraction of the extra workloads (N = 1: a small oracle sample counts the reference's tests)
    for name, x in extras.items():
        xk, xr = x.pop("_kern_ms"), x.pop("_rank_rays")
        xw, xh, xdepth, xint = x.pop("_job")
        if world == 1 and not args.no_cpu_baseline and xk:
            small = {"final": 2, "mesh": 1, "cornell_smoke": 8, "cornell_smoke_legacy": 4, "random": 8}.get(name, 1)
            r = cpu_oracle_run(VARIANT_OF.get(name, (name, None))[0], xw, xh, small, xdepth, xint)
            _, _, xflops = algorithmic_work(r["counters"])
            x["roofline"] = fp64_roofline(xflops, xr / len(xk), sum(xk) / len(xk) * 1e-3, fp64_peak,
                                          kernel_names(x["config"]["pipeline"]))
            x["cpu_baseline"] = {"value": r["paths"] / r["seconds"] / 1e6, "unit": "Mpaths/s", "cores": r["threads"],
                                 "kind": "port", "sample": "%dx%d, %d spp, %.1f s" % (xw, xh, small, r["seconds"])}

    line = {
        "metric": "Mpaths/s", "value": dev["value"], "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev["total_ms"] / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mrays_per_s": dev["mrays"], "segments_per_path": dev["rays"] / max(dev["paths"], 1.0),
        "config": {"workload": WORKLOADS[args.workload], "scene": args.workload, "width": W, "height": H, "spp": spp,
                   "max_depth": depth, "integrator": "HEAD" if hs_integrator == 0 else "LEGACY",
                   "parallelism": "sample blocks x%d + 1 ncclReduce" % world if world > 1 else "1 GPU",
                   "l2": "256 MiB device memset between timed steps (L2 flush)", "seed": 1,
                   "pipeline": pipeline, "variant": info.get("variant"), "pipeline_info": info},
        "clocks": clocks,
        "e2e": e2e,
        "e2e_ppm": e2e_ppm,
        "gpu_launches": int(dev["launches"]),  # this rank's render + reduce kernels inside the timed steps (RtStats.kernel_launches)
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
        "workloads": extras,
        "algorithmic_tests_per_segment": per_seg, "checksum": dev["checksum"],
    }
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
