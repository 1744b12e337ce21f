"""_scenes.py — the scene-graph half of the host layer (lib/librtb200_scenes.so): scene catalogue, Camera::new,
flatten() -> RtSceneDesc, format_color and the PPM writer.

This module does not load the CUDA library, so a process that only needs scene descriptions (the CPU oracle arm of
bench.py) can import it without `librtb200.so` in its address space.  The package's __init__ builds on it.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import RtCamera, RtRenderOpts, RtSceneDesc, INTEGRATOR_HEAD

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("RTB200_LIB_DIR") or os.path.join(_HERE, "lib")  # override: tuning sweeps over prebuilt variants
REPO_ROOT = os.path.dirname(_HERE)
ASSETS_DIR = os.path.join(REPO_ROOT, "assets")


class RtError(RuntimeError):
    def __init__(self, status, message):
        name = _abi.STATUS_NAMES[status] if 0 <= status < len(_abi.STATUS_NAMES) else str(status)
        super().__init__("%s: %s" % (name, message))
        self.status = status


def _load(name):
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise ImportError(
            "%s is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C raytracinginrust_b200/csrc`. There is no fallback path." % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


_lib = _load("librtb200_scenes.so")
_u32p = C.POINTER(C.c_uint32)
_lib.rth_last_error.restype = C.c_char_p
_lib.rth_scene_build.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
_lib.rth_scene_free.argtypes = [C.c_void_p]
_lib.rth_scene_free.restype = None
_lib.rth_scene_desc.argtypes = [C.c_void_p]
_lib.rth_scene_desc.restype = C.POINTER(RtSceneDesc)
_lib.rth_scene_camera.argtypes = [C.c_void_p]
_lib.rth_scene_camera.restype = C.POINTER(RtCamera)
_lib.rth_scene_config.argtypes = [C.c_void_p, _u32p]
_lib.rth_scene_config.restype = None
_lib.rth_camera_new.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                 C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(RtCamera)]
_lib.rth_camera_new.restype = None
_lib.rth_format_image.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
_lib.rth_format_image.restype = None
_lib.rth_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64]
_lib.rth_obj_triangle_count.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]


def _check_host(status):
    if status != _abi.RT_OK:
        raise RtError(status, _lib.rth_last_error().decode())


def render_opts(seed=1, integrator=INTEGRATOR_HEAD, sample_begin=0, sample_count=0, flags=0):
    o = RtRenderOpts()
    o.seed, o.integrator, o.sample_begin, o.sample_count, o.flags = seed, integrator, sample_begin, sample_count, flags
    return o


def _vec3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def camera_new(lookfrom, lookat, vup, vfov, aspect_ratio, aperture, focus_dist, time0=0.0, time1=1.0):
    """Camera::new (src/camera.rs:19-49)."""
    cam = RtCamera()
    _lib.rth_camera_new(_vec3(lookfrom), _vec3(lookat), _vec3(vup), vfov, aspect_ratio, aperture, focus_dist, time0,
                         time1, C.byref(cam))
    return cam


class SceneDesc:
    """An RtSceneDesc plus the buffers it points into."""

    def __init__(self, desc, keepalive=()):
        self.desc = desc
        self._keep = keepalive

    @property
    def ptr(self):
        return C.pointer(self.desc) if not isinstance(self.desc, C.POINTER(RtSceneDesc)) else self.desc

    @property
    def struct(self):
        return self.desc.contents if isinstance(self.desc, C.POINTER(RtSceneDesc)) else self.desc


class HostScene:
    """One of the reference's scenes (src/main.rs:153-513) built by the C++ host layer."""

    NAMES = ("random", "cornell", "cornell_smoke", "final", "mesh", "light_room", "two_spheres")

    def __init__(self, name, construction_seed=1, assets_dir=ASSETS_DIR, mesh_detail=0):
        self.name = name
        self._h = C.c_void_p()
        _check_host(_lib.rth_scene_build(name.encode(), construction_seed, assets_dir.encode(), mesh_detail,
                                          C.byref(self._h)))
        cfg = (C.c_uint32 * 5)()
        _lib.rth_scene_config(self._h, cfg)
        self.integrator, self.width, self.height, self.spp, self.max_depth = [int(x) for x in cfg]
        self.camera = _lib.rth_scene_camera(self._h).contents
        self.scene_desc = SceneDesc(_lib.rth_scene_desc(self._h), (self,))

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.rth_scene_free(self._h)
            self._h = None


# ---------------------------------------------------------------------------------------
# Output: format_color (src/vec.rs:125-131) and the P3 writer (src/main.rs:767-769,832)
# ---------------------------------------------------------------------------------------
def format_image(rgb_sum, samples_per_pixel):
    rgb_sum = np.ascontiguousarray(rgb_sum, dtype=np.float32)
    out = np.empty(rgb_sum.shape, dtype=np.uint8)
    _lib.rth_format_image(rgb_sum.ctypes.data_as(C.c_void_p), rgb_sum.size // 3, samples_per_pixel,
                           out.ctypes.data_as(C.c_void_p))
    return out


def write_ppm(path, rgb_sum, samples_per_pixel):
    rgb_sum = np.ascontiguousarray(rgb_sum, dtype=np.float32)
    h, w = rgb_sum.shape[:2]
    _check_host(_lib.rth_write_ppm(path.encode(), rgb_sum.ctypes.data_as(C.c_void_p), w, h, samples_per_pixel))


def obj_triangle_count(path):
    nv, nt = C.c_uint64(), C.c_uint64()
    _check_host(_lib.rth_obj_triangle_count(path.encode(), C.byref(nv), C.byref(nt)))
    return int(nv.value), int(nt.value)
