"""ctypes wrapper of tests/native/build/libtrace_on_host.so — TEST INFRASTRUCTURE ONLY.

The device source of the hot path (csrc/device/trace.cuh) compiled with g++, behind loops that do what
the kernel shells do per thread (see trace_on_host.cpp).  Only tests/ may import this; it is not a CPU
path of the product and nothing under raytracinginrust_b200/ knows it exists.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from raytracinginrust_b200._abi import HIT_DTYPE, RAY_DTYPE, RtCamera, RtRenderOpts, RtSceneDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "build", "libtrace_on_host.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _lib():
    override = os.environ.get("RTB200_TRACE_ON_HOST_LIB")  # e.g. an -fsanitize=address,undefined build of the same sources
    if not override:
        build()
    lib = C.CDLL(override or LIB_PATH)
    lib.toh_last_error.restype = C.c_char_p
    lib.toh_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.POINTER(C.c_void_p)]
    lib.toh_scene_destroy.argtypes = [C.c_void_p]
    lib.toh_scene_destroy.restype = None
    lib.toh_check_tables.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.toh_shutter_limited.argtypes = [C.c_void_p]
    lib.toh_tables_hash.argtypes = [C.c_void_p]
    lib.toh_tables_hash.restype = C.c_uint64
    lib.toh_trace_first_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.toh_camera_rays.argtypes = [C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.POINTER(RtRenderOpts), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.toh_path_radiance.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(RtRenderOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_void_p]
    lib.toh_path_radiance_split.argtypes = lib.toh_path_radiance.argtypes
    lib.toh_render.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                               C.POINTER(RtRenderOpts), C.c_void_p, C.POINTER(C.c_uint64)]
    return lib


lib = _lib()


class TraceOnHostError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("status %d: %s" % (status, message))
        self.status = status


def _check(st):
    if st != 0:
        raise TraceOnHostError(st, lib.toh_last_error().decode())


TABLE_COUNTS = ("prims", "groups", "world_groups", "nodes", "media", "lights", "chains", "bvh_depth")


class CompiledOnHost:
    """compile_scene(desc) -> the flat tables, with the device functions of trace.cuh run over them on the CPU."""

    def __init__(self, scene_desc):
        self._h = C.c_void_p()
        self._desc = scene_desc
        _check(lib.toh_scene_create(scene_desc.ptr, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib.toh_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def check_tables(self):
        """Raises if a structural invariant of the compiled tables is broken; returns the table sizes."""
        counts = (C.c_uint64 * 8)()
        _check(lib.toh_check_tables(self._h, counts))
        return dict(zip(TABLE_COUNTS, (int(c) for c in counts)))

    @property
    def shutter_limited(self):
        return bool(lib.toh_shutter_limited(self._h))

    def tables_hash(self):
        """FNV-1a over every compiled table (padding bytes are zeroed by the compiler, so it is a pure function of the
        description)."""
        return int(lib.toh_tables_hash(self._h))

    def trace_first_hit(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        _check(lib.toh_trace_first_hit(self._h, rays.ctypes.data_as(C.c_void_p), rays.shape[0],
                                       hits.ctypes.data_as(C.c_void_p)))
        return hits

    def path_radiance(self, camera, width, height, max_depth, opts, px, py, sample):
        px, py, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (px, py, sample))
        n = px.shape[0]
        rgb = np.zeros((n, 3), dtype=np.float64)
        seg = np.zeros(n, dtype=np.uint32)
        _check(lib.toh_path_radiance(self._h, C.byref(camera), width, height, max_depth, C.byref(opts),
                                     px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p),
                                     sample.ctypes.data_as(C.c_void_p), n, rgb.ctypes.data_as(C.c_void_p),
                                     seg.ctypes.data_as(C.c_void_p)))
        return rgb, seg

    def path_radiance_split(self, camera, width, height, max_depth, opts, px, py, sample):
        """path_radiance in the wavefront stages' form: world_search, then resolve + path_shade."""
        px, py, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (px, py, sample))
        n = px.shape[0]
        rgb = np.zeros((n, 3), dtype=np.float64)
        seg = np.zeros(n, dtype=np.uint32)
        _check(lib.toh_path_radiance_split(self._h, C.byref(camera), width, height, max_depth, C.byref(opts),
                                           px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p),
                                           sample.ctypes.data_as(C.c_void_p), n, rgb.ctypes.data_as(C.c_void_p),
                                           seg.ctypes.data_as(C.c_void_p)))
        return rgb, seg

    def render(self, camera, width, height, spp, max_depth, opts):
        """Returns (f64 sums HxWx3 rows top-down, {paths, rays, non_finite})."""
        out = np.zeros((height, width, 3), dtype=np.float64)
        stats = (C.c_uint64 * 3)()
        _check(lib.toh_render(self._h, C.byref(camera), width, height, spp, max_depth, C.byref(opts),
                              out.ctypes.data_as(C.c_void_p), stats))
        return out, {"paths": int(stats[0]), "rays": int(stats[1]), "non_finite": int(stats[2])}


def camera_rays(camera, width, height, opts, px, py, sample):
    px, py, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (px, py, sample))
    n = px.shape[0]
    rays = np.zeros(n, dtype=RAY_DTYPE)
    _check(lib.toh_camera_rays(C.byref(camera), width, height, C.byref(opts), px.ctypes.data_as(C.c_void_p),
                               py.ctypes.data_as(C.c_void_p), sample.ctypes.data_as(C.c_void_p), n,
                               rays.ctypes.data_as(C.c_void_p)))
    return rays
