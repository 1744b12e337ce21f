"""raytracinginrust_b200 — B200-native path-tracing hot path for the RayTracingInRust engine.

This package is the Python harness over two native libraries built by
`__graft_entry__.build()` (or `make -C raytracinginrust_b200/csrc`):

  lib/librtb200.so       the product: CUDA kernels for sm_100a behind the C ABI of
                         include/rtb200.h (what a Rust host links against)
  lib/librtb200_scenes.so  the scene-graph half of the C++ host layer standing in for the
                         Rust host: the reference's scene constructors, Camera::new, flatten(),
                         format_color and the PPM writer (loaded by _scenes.py; no CUDA dependency)
  lib/librtb200_host.so  render() / render_ppm() over the C ABI

There is no CPU fallback anywhere in this package: if the native libraries are not
built, importing fails; if there is no CUDA device, every compute call raises RtError.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (RtCamera, RtHit, RtImage, RtMaterial, RtNode, RtPerlin, RtRay, RtRenderOpts, RtSceneDesc, RtStats,
                   RtTexture, HIT_DTYPE, RAY_DTYPE, INTEGRATOR_HEAD, INTEGRATOR_LEGACY)

from . import _scenes
from ._scenes import (ASSETS_DIR, LIB_DIR, REPO_ROOT, RtError, SceneDesc, _check_host, _load, _vec3, camera_new,
                    format_image, obj_triangle_count, render_opts, write_ppm)

_dev = _load("librtb200.so")
_host = _load("librtb200_host.so")  # depends on the two above

_u32p = C.POINTER(C.c_uint32)
_dev.rt_device_count.restype = C.c_int
_dev.rt_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.POINTER(C.c_void_p)]
_dev.rt_scene_create_ex.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]
_dev.rt_scene_destroy.argtypes = [C.c_void_p]
_dev.rt_scene_destroy.restype = None
_dev.rt_release_cached_memory.argtypes = []
_dev.rt_release_cached_memory.restype = None
_dev.rt_scene_device_bytes.argtypes = [C.c_void_p]
_dev.rt_scene_device_bytes.restype = C.c_uint64
_dev.rt_render_info.argtypes = [C.c_void_p]
_dev.rt_render_info.restype = C.c_char_p
_dev.rt_render.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                           C.POINTER(RtRenderOpts), C.c_void_p, C.POINTER(RtStats)]
_dev.rt_render_device.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(RtRenderOpts), C.c_void_p, C.c_void_p]
_dev.rt_render_wait.argtypes = [C.c_void_p, C.POINTER(RtStats)]
_dev.rt_trace_first_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
_dev.rt_path_radiance.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(RtRenderOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                  C.c_void_p]
_dev.rt_camera_rays.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.POINTER(RtRenderOpts),
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
_dev.rt_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
_dev.rt_last_error.restype = C.c_char_p
_dev.rt_version.restype = C.c_char_p
_dev.rt_scene_group_create.argtypes = [C.POINTER(RtSceneDesc), C.POINTER(C.c_int), C.c_uint32, C.POINTER(C.c_void_p)]
_dev.rt_scene_group_destroy.argtypes = [C.c_void_p]
_dev.rt_scene_group_destroy.restype = None
_dev.rt_scene_group_size.argtypes = [C.c_void_p]
_dev.rt_scene_group_size.restype = C.c_uint32
_dev.rt_scene_group_scene.argtypes = [C.c_void_p, C.c_uint32]
_dev.rt_scene_group_scene.restype = C.c_void_p
_dev.rt_render_multi.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.POINTER(RtRenderOpts), C.c_void_p, C.POINTER(RtStats)]
_dev.rt_encode_rgb8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
_dev.rt_encode_ppm.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64,
                               C.POINTER(C.c_uint64)]

_dev.rt_compile.argtypes = [C.POINTER(RtSceneDesc), C.POINTER(C.c_void_p)]
_dev.rt_compiled_data.argtypes = [C.c_void_p]
_dev.rt_compiled_data.restype = C.c_void_p
_dev.rt_compiled_size.argtypes = [C.c_void_p]
_dev.rt_compiled_size.restype = C.c_uint64
_dev.rt_compiled_hash.argtypes = [C.c_void_p, C.c_uint64]
_dev.rt_compiled_hash.restype = C.c_uint64
_dev.rt_compiled_destroy.argtypes = [C.c_void_p]
_dev.rt_compiled_destroy.restype = None
_dev.rt_scene_create_compiled.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]

_host.rth_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(RtRenderOpts),
                             C.c_int, C.c_void_p, C.POINTER(RtStats)]
_host.rth_render_ppm.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(RtRenderOpts),
                                 C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(RtStats)]


def _check(status):
    if status != _abi.RT_OK:
        raise RtError(status, _dev.rt_last_error().decode())


def device_count():
    return int(_dev.rt_device_count())


def release_cached_memory():
    """rt_release_cached_memory: return the device blocks parked by destroyed scenes to the driver."""
    _dev.rt_release_cached_memory()


def measure_fp64_peak(device=0):
    """Measured FP64 FMA throughput in TFLOP/s (the roofline denominator for this path)."""
    v = C.c_double()
    _check(_dev.rt_measure_fp64_peak(device, C.byref(v)))
    return float(v.value)


def version():
    return _dev.rt_version().decode()


# ---------------------------------------------------------------------------------------
# Scene descriptions
# ---------------------------------------------------------------------------------------
class SceneBuilder:
    """Python-side flatten target: builds an RtSceneDesc node by node.  Method names and
    argument order follow the reference's constructors (Sphere::new, AARect::new, ...)."""

    def __init__(self):
        self.nodes, self.child_index, self.materials, self.textures = [], [], [], []
        self.perlin, self.images, self.texels = [], [], bytearray()

    # textures (src/texture.rs)
    def _tex(self, kind, a=_abi.RT_NONE, b=_abi.RT_NONE, color=(0, 0, 0), scale=0.0):
        t = RtTexture()
        t.kind, t.a, t.b, t.scale = kind, a, b, scale
        t.color[:] = [float(c) for c in color]
        self.textures.append(t)
        return len(self.textures) - 1

    def constant_texture(self, color):
        return self._tex(_abi.TEX_CONSTANT, color=color)

    def check_texture(self, odd, even):
        return self._tex(_abi.TEX_CHECKER, a=odd, b=even)

    def noise_texture(self, scale, ranvec, perm_x, perm_y, perm_z):
        p = RtPerlin()
        p.ranvec[:] = [float(x) for x in np.asarray(ranvec, dtype=np.float64).reshape(-1)]
        p.perm_x[:] = [int(x) for x in perm_x]
        p.perm_y[:] = [int(x) for x in perm_y]
        p.perm_z[:] = [int(x) for x in perm_z]
        self.perlin.append(p)
        return self._tex(_abi.TEX_NOISE, a=len(self.perlin) - 1, scale=scale)

    def image_texture(self, data, width, height):
        im = RtImage()
        im.width, im.height, im.offset = width, height, len(self.texels)
        raw = bytes(data)
        assert len(raw) == width * height * 3
        self.texels += raw
        self.images.append(im)
        return self._tex(_abi.TEX_IMAGE, a=len(self.images) - 1)

    # materials (src/mat.rs)
    def _mat(self, kind, texture=_abi.RT_NONE, albedo=(0, 0, 0), fuzz=0.0, ir=0.0):
        m = RtMaterial()
        m.kind, m.texture, m.fuzz, m.ir = kind, texture, fuzz, ir
        m.albedo[:] = [float(c) for c in albedo]
        self.materials.append(m)
        return len(self.materials) - 1

    def lambertian(self, texture):
        return self._mat(_abi.MAT_LAMBERTIAN, texture=texture)

    def metal(self, albedo, fuzz):
        return self._mat(_abi.MAT_METAL, albedo=albedo, fuzz=fuzz)

    def dielectric(self, ir):
        return self._mat(_abi.MAT_DIELECTRIC, ir=ir)

    def diffuse_light(self, texture):
        return self._mat(_abi.MAT_DIFFUSE_LIGHT, texture=texture)

    def isotropic(self, texture):
        return self._mat(_abi.MAT_ISOTROPIC, texture=texture)

    def pbr(self, base_color, metallic=0.0, subsurface=0.0, specular=0.0, roughness=1.0, specular_tint=0.0,
            anisotropic=0.0, sheen=0.0, sheen_tint=0.0, clearcoat=0.0, clearcoat_gloss=0.0):
        """PBR::new(base_color texture, ...) in the argument order of src/mat.rs:101-115."""
        k = self._mat(_abi.MAT_PBR, texture=base_color)
        self.materials[k].pbr[:] = [float(x) for x in (metallic, subsurface, specular, roughness, specular_tint,
                                                       anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss)]
        return k

    # hittables
    def _node(self, kind, material=_abi.RT_NONE, child=_abi.RT_NONE, count=0, axis=0, v=()):
        n = RtNode()
        n.kind, n.material, n.child, n.count, n.axis = kind, material, child, count, axis
        for i, x in enumerate(v):
            n.v[i] = float(x)
        self.nodes.append(n)
        return len(self.nodes) - 1

    def sphere(self, center, radius, material):
        return self._node(_abi.NODE_SPHERE, material, v=list(center) + [radius])

    def moving_sphere(self, center0, center1, time0, time1, radius, material):
        return self._node(_abi.NODE_MOVING_SPHERE, material, v=list(center0) + list(center1) + [time0, time1, radius])

    def rect(self, plane, a0, a1, b0, b1, k, material):
        return self._node(_abi.NODE_RECT, material, axis=plane, v=[a0, a1, b0, b1, k])

    def triangle(self, v0, v1, v2, material):
        return self._node(_abi.NODE_TRIANGLE, material, v=list(v0) + list(v1) + list(v2))

    def cube(self, pmin, pmax, material):
        return self._node(_abi.NODE_CUBE, material, v=list(pmin) + list(pmax))

    def _children(self, kind, kids, v=()):
        first = len(self.child_index)
        self.child_index += [int(k) for k in kids]
        return self._node(kind, child=first, count=len(kids), v=v)

    def list(self, kids):
        return self._children(_abi.NODE_LIST, kids)

    def bvh(self, kids, time0=0.0, time1=1.0):
        return self._children(_abi.NODE_BVH, kids, v=[time0, time1])

    def translate(self, child, offset):
        return self._node(_abi.NODE_TRANSLATE, child=child, v=list(offset))

    def rotate(self, axis, child, angle):
        return self._node(_abi.NODE_ROTATE, child=child, axis=axis, v=[angle])

    def flip(self, child):
        return self._node(_abi.NODE_FLIP, child=child)

    def medium(self, boundary, density, texture):
        return self._node(_abi.NODE_MEDIUM, material=self.isotropic(texture), child=boundary, v=[density])

    def finish(self, world, lights, background=(0.0, 0.0, 0.0)):
        def arr(ctype, items):
            a = (ctype * max(len(items), 1))(*items)
            return a

        nodes = arr(RtNode, self.nodes)
        kids = (C.c_uint32 * max(len(self.child_index), 1))(*self.child_index)
        mats = arr(RtMaterial, self.materials)
        texs = arr(RtTexture, self.textures)
        perl = arr(RtPerlin, self.perlin)
        imgs = arr(RtImage, self.images)
        texels = (C.c_uint8 * max(len(self.texels), 1)).from_buffer_copy(bytes(self.texels) or b"\0")
        d = RtSceneDesc()
        d.abi_version = _abi.ABI_VERSION
        d.world, d.lights = world, lights
        d.background[:] = [float(c) for c in background]
        d.nodes, d.n_nodes = nodes, len(self.nodes)
        d.child_index, d.n_child_index = kids, len(self.child_index)
        d.materials, d.n_materials = mats, len(self.materials)
        d.textures, d.n_textures = texs, len(self.textures)
        d.perlin, d.n_perlin = perl, len(self.perlin)
        d.images, d.n_images = imgs, len(self.images)
        d.texels, d.n_texel_bytes = texels, len(self.texels)
        return SceneDesc(d, (nodes, kids, mats, texs, perl, imgs, texels))


class HostScene(_scenes.HostScene):
    """One of the reference's scenes (src/main.rs:153-513) built by the C++ host layer, with render()."""

    def render(self, width, height, spp, max_depth, opts=None, device=0):
        """render(world, camera, width, height, spp, max_depth) -> pixels, end to end with host
        buffers: flatten + compile + upload + kernels + read-back (src/main.rs:772-834)."""
        opts = opts or render_opts(integrator=self.integrator)
        out = np.empty((height, width, 3), dtype=np.float32)
        stats = RtStats()
        _check_host(_host.rth_render(self._h, width, height, spp, max_depth, C.byref(opts), device,
                                     out.ctypes.data_as(C.c_void_p), C.byref(stats)))
        return out, stats


    def render_ppm(self, width, height, spp, max_depth, opts=None, n_gpus=1):
        """render() over n_gpus GPUs (0 = all) with format_color and the P3 text done on the GPU -> bytes:
        what `rtb200_render --scene ... > image.ppm` writes."""
        opts = opts or render_opts(integrator=self.integrator)
        cap = 32 + 12 * width * height
        buf = np.empty(cap, dtype=np.uint8)
        n, stats = C.c_uint64(), RtStats()
        _check_host(_host.rth_render_ppm(self._h, width, height, spp, max_depth, C.byref(opts), n_gpus,
                                         buf.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(stats)))
        return buf[:int(n.value)].tobytes(), stats


def compile_scene(scene_desc):
    """rt_compile: the host half of rt_scene_create (graph walk, reference order, SAH BVH build) -> the relocatable
    blob as a numpy uint8 array.  Needs no GPU.  One rank compiles, the bytes travel, every rank calls
    DeviceScene.from_compiled (multi_gpu.render_distributed does that over torch.distributed)."""
    h = C.c_void_p()
    _check(_dev.rt_compile(scene_desc.ptr, C.byref(h)))
    n = int(_dev.rt_compiled_size(h))
    # a view of the library's buffer, not a copy (64 MB for config 5): the handle lives as long as the array does
    buf = (C.c_ubyte * n).from_address(_dev.rt_compiled_data(h))
    buf._owner = _CompiledHandle(h)
    return np.frombuffer(buf, dtype=np.uint8)


class _CompiledHandle:
    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if self._h is not None and _dev is not None:
            _dev.rt_compiled_destroy(self._h)
            self._h = None


def compiled_hash(blob):
    """FNV-1a over the tables of a compiled scene (0: not a blob of this library)."""
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    return int(_dev.rt_compiled_hash(blob.ctypes.data_as(C.c_void_p), blob.size))


class DeviceScene:
    """A scene compiled and resident on one GPU (rt_scene_create)."""

    def __init__(self, scene_desc, device=0, _borrowed=None, _compiled=None, gpu_bvh=False):
        """gpu_bvh=True: rt_scene_create_ex(RT_CREATE_GPU_BVH) - trees of 4096+ primitives are built on the GPU
        (a linear BVH in about a millisecond instead of the host's SAH build; same images, slower traversal)."""
        self._h = C.c_void_p()
        self._desc = scene_desc
        self._owned = _borrowed is None
        if gpu_bvh:
            _check(_dev.rt_scene_create_ex(scene_desc.ptr, device, _abi.CREATE_GPU_BVH, C.byref(self._h)))
        elif _compiled is not None:
            blob = np.ascontiguousarray(_compiled, dtype=np.uint8)
            _check(_dev.rt_scene_create_compiled(blob.ctypes.data_as(C.c_void_p), blob.size, device, C.byref(self._h)))
        elif _borrowed is None:
            _check(_dev.rt_scene_create(scene_desc.ptr, device, C.byref(self._h)))
        else:  # a member of a SceneGroup: the group owns the handle
            self._h = C.c_void_p(_borrowed)
        self.device = device

    @classmethod
    def from_compiled(cls, blob, device=0):
        """rt_scene_create_compiled: upload the tables of a blob made by compile_scene (here or on another rank)."""
        return cls(None, device=device, _compiled=blob)

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                _dev.rt_scene_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def device_bytes(self):
        return int(_dev.rt_scene_device_bytes(self._h))

    @property
    def render_info(self):
        """Which pipeline build the last render ran: {'pipeline': ..., 'variant': ..., ...}."""
        text = (_dev.rt_render_info(self._h) or b"").decode()
        return dict(kv.split("=", 1) for kv in text.split() if "=" in kv)

    def render(self, camera, width, height, spp, max_depth, opts=None, want_sums=True):
        """rt_render.  want_sums=False leaves the image on the device (for encode_rgb8 / encode_ppm)."""
        opts = opts or render_opts()
        out = np.empty((height, width, 3), dtype=np.float32) if want_sums else None
        stats = RtStats()
        _check(_dev.rt_render(self._h, C.byref(camera), width, height, spp, max_depth, C.byref(opts),
                              out.ctypes.data_as(C.c_void_p) if want_sums else None, C.byref(stats)))
        return out, stats

    def encode_rgb8(self, width, height, samples_per_pixel, sums_ptr=0):
        """Vec3::format_color (src/vec.rs:125-131) on the GPU -> (H, W, 3) uint8.  sums_ptr: a W*H*3 fp32
        device buffer on this scene's GPU, or 0 for the image the last render left resident."""
        out = np.empty((height, width, 3), dtype=np.uint8)
        _check(_dev.rt_encode_rgb8(self._h, C.c_void_p(sums_ptr) if sums_ptr else None, width, height, samples_per_pixel,
                                   out.ctypes.data_as(C.c_void_p)))
        return out

    def encode_ppm(self, width, height, samples_per_pixel, sums_ptr=0, capacity=None):
        """The whole P3 file (src/main.rs:767-769,832), formatted on the GPU -> bytes."""
        cap = int(capacity if capacity is not None else 32 + 12 * width * height)
        buf = np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_uint64()
        _check(_dev.rt_encode_ppm(self._h, C.c_void_p(sums_ptr) if sums_ptr else None, width, height, samples_per_pixel,
                                  buf.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return buf[:int(n.value)].tobytes()

    def render_device(self, camera, width, height, spp, max_depth, opts, out_ptr, stream_ptr=0):
        """Enqueue a render into device memory `out_ptr` (W*H*3 fp32) on CUDA stream `stream_ptr`."""
        _check(_dev.rt_render_device(self._h, C.byref(camera), width, height, spp, max_depth, C.byref(opts),
                                     C.c_void_p(out_ptr), C.c_void_p(stream_ptr)))

    def render_wait(self):
        stats = RtStats()
        _check(_dev.rt_render_wait(self._h, C.byref(stats)))
        return stats

    def trace_first_hit(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        _check(_dev.rt_trace_first_hit(self._h, rays.ctypes.data_as(C.c_void_p), rays.shape[0],
                                       hits.ctypes.data_as(C.c_void_p)))
        return hits

    def path_radiance(self, camera, width, height, max_depth, opts, px, py, sample):
        px, py, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (px, py, sample))
        n = px.shape[0]
        rgb = np.zeros((n, 3), dtype=np.float64)
        seg = np.zeros(n, dtype=np.uint32)
        _check(_dev.rt_path_radiance(self._h, C.byref(camera), width, height, max_depth, C.byref(opts),
                                     px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p),
                                     sample.ctypes.data_as(C.c_void_p), n, rgb.ctypes.data_as(C.c_void_p),
                                     seg.ctypes.data_as(C.c_void_p)))
        return rgb, seg

    def camera_rays(self, camera, width, height, opts, px, py, sample):
        px, py, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (px, py, sample))
        n = px.shape[0]
        rays = np.zeros(n, dtype=RAY_DTYPE)
        _check(_dev.rt_camera_rays(self._h, C.byref(camera), width, height, C.byref(opts),
                                   px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p),
                                   sample.ctypes.data_as(C.c_void_p), n, rays.ctypes.data_as(C.c_void_p)))
        return rays


class SceneGroup:
    """A scene replicated on several GPUs of one box, driven by ONE host thread (rt_scene_group_create /
    rt_render_multi): contiguous sample blocks per GPU, one combine kernel over NVLink peer memory."""

    def __init__(self, scene_desc, devices=None):
        self._h = C.c_void_p()
        self._desc = scene_desc
        if devices is None:
            arr, n = None, 0
        else:
            devices = [int(d) for d in devices]
            arr, n = (C.c_int * len(devices))(*devices), len(devices)
        _check(_dev.rt_scene_group_create(scene_desc.ptr, arr, n, C.byref(self._h)))
        self.size = int(_dev.rt_scene_group_size(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _dev.rt_scene_group_destroy(self._h)
            self._h = None

    __del__ = close

    def scene(self, i=0):
        """The i-th per-device scene (0 = the root, which holds the combined image)."""
        h = _dev.rt_scene_group_scene(self._h, i)
        if not h:
            raise IndexError(i)
        return DeviceScene(self._desc, _borrowed=h)

    def render(self, camera, width, height, spp, max_depth, opts=None, want_sums=True):
        opts = opts or render_opts()
        out = np.empty((height, width, 3), dtype=np.float32) if want_sums else None
        stats = RtStats()
        _check(_dev.rt_render_multi(self._h, C.byref(camera), width, height, spp, max_depth, C.byref(opts),
                                    out.ctypes.data_as(C.c_void_p) if want_sums else None, C.byref(stats)))
        return out, stats
