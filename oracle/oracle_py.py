"""ctypes wrapper of oracle/build/liboracle.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by raytracinginrust_b200/ (see oracle.cpp header).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from raytracinginrust_b200._abi import (HIT_DTYPE, RAY_DTYPE, RtCamera, RtRenderOpts, RtSceneDesc)  # noqa: E402

LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def build_native():
    """The timed build of bench.py's CPU legs: -O3 -march=native (still -ffp-contract=off: rustc never fuses), compiled
    ON the machine that runs it - the default build travels to the GPU box from another CPU, so it stays generic.
    Returns the library's path (cached per CPU flag set)."""
    import hashlib
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except (OSError, StopIteration):
        flags = "unknown"
    out = os.path.join("build", "liboracle_native_%s.so" % hashlib.md5(flags.encode()).hexdigest()[:10])
    subprocess.check_call(["make", "-C", _HERE, "-s", "OUT=" + out,
                           "CXXFLAGS=-O3 -march=native -std=c++17 -fPIC -fopenmp -ffp-contract=off -fno-fast-math"])
    return os.path.join(_HERE, out)


class OracleCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("segments", "box_tests", "sphere_tests", "msphere_tests", "rect_tests",
                                          "tri_tests", "medium_tests", "xform")]


def _lib(path=LIB_PATH):
    if path == LIB_PATH and not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(path)
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.POINTER(C.c_void_p)]
    lib.oracle_scene_destroy.argtypes = [C.c_void_p]
    lib.oracle_scene_destroy.restype = None
    lib.oracle_trace_first_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.oracle_camera_rays.argtypes = [C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.POINTER(RtRenderOpts), C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.oracle_path_radiance.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(RtRenderOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_void_p]
    lib.oracle_render.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(RtRenderOpts), C.c_void_p, C.c_int, C.POINTER(C.c_uint64),
                                  C.POINTER(OracleCounters)]
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_draw.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32),
                                                    C.POINTER(C.c_uint32)]
    lib.oracle_draw.restype = None
    dp = C.POINTER(C.c_double)
    lib.oracle_sphere_uv.argtypes = [dp, dp, dp]
    lib.oracle_sphere_uv.restype = None
    lib.oracle_onb.argtypes = [dp, dp]
    lib.oracle_onb.restype = None
    lib.oracle_reflect.argtypes = [dp, dp, dp]
    lib.oracle_reflect.restype = None
    lib.oracle_refract.argtypes = [dp, dp, C.c_double, dp]
    lib.oracle_refract.restype = None
    lib.oracle_reflectance.argtypes = [C.c_double, C.c_double]
    lib.oracle_reflectance.restype = C.c_double
    lib.oracle_pbr_scalar.argtypes = [C.c_int] + [C.c_double] * 5
    lib.oracle_pbr_scalar.restype = C.c_double
    lib.oracle_random_cosine_direction.argtypes = [C.c_double, C.c_double, dp]
    lib.oracle_random_cosine_direction.restype = None
    lib.oracle_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_double, dp, dp]
    lib.oracle_texture.restype = None
    lib.oracle_format_color.argtypes = [dp, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.oracle_format_color.restype = None
    lib.oracle_light_pdf.argtypes = [C.c_void_p, dp, dp]
    lib.oracle_light_pdf.restype = C.c_double
    lib.oracle_bvh_stats.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


lib = _lib()


def use_native_build():
    """Switch this module to the -O3 -march=native build (bench.py's timed CPU legs).  Call before creating scenes."""
    global lib
    lib = _lib(build_native())
    return lib


class OracleError(RuntimeError):
    pass


def _check(st):
    if st != 0:
        raise OracleError("status %d: %s" % (st, lib.oracle_last_error().decode()))


def _d3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


class OracleScene:
    """The reference's object graph rebuilt from an RtSceneDesc (incl. its median-split BVH)."""

    def __init__(self, scene_desc):
        self._h = C.c_void_p()
        self._desc = scene_desc
        _check(lib.oracle_scene_create(scene_desc.ptr, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib.oracle_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def trace_first_hit(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        _check(lib.oracle_trace_first_hit(self._h, rays.ctypes.data_as(C.c_void_p), rays.shape[0],
                                          hits.ctypes.data_as(C.c_void_p)))
        return hits

    def path_radiance(self, camera, width, height, max_depth, opts, px, py, sample):
        px, py, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (px, py, sample))
        n = px.shape[0]
        rgb = np.zeros((n, 3), dtype=np.float64)
        seg = np.zeros(n, dtype=np.uint32)
        _check(lib.oracle_path_radiance(self._h, C.byref(camera), width, height, max_depth, C.byref(opts),
                                        px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p),
                                        sample.ctypes.data_as(C.c_void_p), n, rgb.ctypes.data_as(C.c_void_p),
                                        seg.ctypes.data_as(C.c_void_p)))
        return rgb, seg

    def render(self, camera, width, height, spp, max_depth, opts, threads=0, counters=False):
        """Returns (f64 sums HxWx3 rows top-down, rays[, counters])."""
        out = np.zeros((height, width, 3), dtype=np.float64)
        rays = C.c_uint64()
        cnt = OracleCounters()
        _check(lib.oracle_render(self._h, C.byref(camera), width, height, spp, max_depth, C.byref(opts),
                                 out.ctypes.data_as(C.c_void_p), threads, C.byref(rays),
                                 C.byref(cnt) if counters else None))
        if counters:
            return out, int(rays.value), {n: int(getattr(cnt, n)) for n, _ in OracleCounters._fields_}
        return out, int(rays.value)

    def texture(self, tex_id, u, v, p):
        out = (C.c_double * 3)()
        lib.oracle_texture(self._h, tex_id, u, v, _d3(p), out)
        return np.array(out[:])

    def light_pdf(self, o, v):
        return float(lib.oracle_light_pdf(self._h, _d3(o), _d3(v)))

    def bvh_stats(self, node):
        d, n = C.c_int(), C.c_int()
        _check(lib.oracle_bvh_stats(self._h, node, C.byref(d), C.byref(n)))
        return int(d.value), int(n.value)


def camera_rays(camera, width, height, opts, px, py, sample):
    px, py, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (px, py, sample))
    n = px.shape[0]
    rays = np.zeros(n, dtype=RAY_DTYPE)
    _check(lib.oracle_camera_rays(C.byref(camera), width, height, C.byref(opts), px.ctypes.data_as(C.c_void_p),
                                  py.ctypes.data_as(C.c_void_p), sample.ctypes.data_as(C.c_void_p), n,
                                  rays.ctypes.data_as(C.c_void_p)))
    return rays


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def draw(seed, pixel, sample, bounce, slot, sub):
    a, b = C.c_double(), C.c_double()
    ba, bb = C.c_uint32(), C.c_uint32()
    lib.oracle_draw(seed, pixel, sample, bounce, slot, sub, C.byref(a), C.byref(b), C.byref(ba), C.byref(bb))
    return a.value, b.value, ba.value, bb.value


def sphere_uv(p):
    u, v = C.c_double(), C.c_double()
    lib.oracle_sphere_uv(_d3(p), C.byref(u), C.byref(v))
    return u.value, v.value


def onb(n):
    out = (C.c_double * 9)()
    lib.oracle_onb(_d3(n), out)
    return np.array(out[:]).reshape(3, 3)


def reflect(v, n):
    out = (C.c_double * 3)()
    lib.oracle_reflect(_d3(v), _d3(n), out)
    return np.array(out[:])


def refract(v, n, eta):
    out = (C.c_double * 3)()
    lib.oracle_refract(_d3(v), _d3(n), eta, out)
    return np.array(out[:])


def reflectance(cosine, ir):
    return float(lib.oracle_reflectance(cosine, ir))


def pbr_scalar(which, a, b=0.0, c=0.0, d=0.0, e=0.0):
    """mat.rs:10-44: 0 schlick_fresnel, 1 GTR_1, 2 GTR_2_aniso, 3 smithG_GGX, 4 smithG_GGX_aniso."""
    return float(lib.oracle_pbr_scalar(int(which), a, b, c, d, e))


def random_cosine_direction(r1, r2):
    out = (C.c_double * 3)()
    lib.oracle_random_cosine_direction(r1, r2, out)
    return np.array(out[:])


def format_color(rgb_sum, spp):
    out = (C.c_uint64 * 3)()
    lib.oracle_format_color(_d3(rgb_sum), spp, out)
    return [int(x) for x in out]


def format_image(rgb_sum, spp):
    """format_color over an HxWx3 array of sums -> uint8 (numpy restatement of vec.rs:125-131)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        x = np.sqrt(np.asarray(rgb_sum, dtype=np.float64) / float(spp))
        x = np.where(x < 0.0, 0.0, x)
        x = np.where(x > 0.999, 0.999, x)
        y = 256.0 * x
        y = np.where(np.isnan(y), 0.0, y)
    return y.astype(np.uint64).astype(np.uint8)


def num_threads():
    return int(lib.oracle_num_threads())
