"""The C-ABI library loads without a GPU and exports every symbol include/rtb200.h declares;
argument and scene validation report through status codes, never by unwinding.  No compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", text)))


def test_header_functions_are_exported(rt):
    names = declared_functions()
    assert sorted(names) == sorted(rt._abi.EXPORTS)
    lib = C.CDLL(os.path.join(rt.LIB_DIR, "librtb200.so"))
    for n in names:
        assert hasattr(lib, n), n


def test_struct_sizes_match_header(rt):
    """Compile-free layout check: sizes implied by the header's field lists."""
    A = rt._abi
    assert C.sizeof(A.RtNode) == 24 + 80 and C.sizeof(A.RtMaterial) == 8 + 40 + 80 and C.sizeof(A.RtTexture) == 16 + 32
    assert C.sizeof(A.RtPerlin) == 768 * 8 + 3 * 1024 and C.sizeof(A.RtImage) == 16
    assert C.sizeof(A.RtCamera) == 21 * 8 and C.sizeof(A.RtRenderOpts) == 24 and C.sizeof(A.RtStats) == 64
    assert C.sizeof(A.RtRay) == 56 and C.sizeof(A.RtHit) == 16 + 9 * 8
    assert np.dtype(A.RAY_DTYPE).itemsize == 56 and np.dtype(A.HIT_DTYPE).itemsize == 88
    assert C.sizeof(A.RtSceneDesc) == 16 + 24 + 7 * 16


def test_version_and_device_count(rt):
    assert "sm_100a" in rt.version() and "f64" in rt.version()
    assert rt.device_count() >= 0


def _simple(rt):
    b = rt.SceneBuilder()
    m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
    s = b.sphere((0, 0, 0), 1.0, m)
    light = b.rect(rt._abi.PLANE_XZ, -1, 1, -1, 1, 5, b.diffuse_light(b.constant_texture((4, 4, 4))))
    return b, s, light


def _status(rt, sd):
    h = C.c_void_p()
    st = rt._dev.rt_scene_create(sd.ptr, 0, C.byref(h))
    msg = rt._dev.rt_last_error().decode()
    if st == 0:
        rt._dev.rt_scene_destroy(h)
    return st, msg


def test_scene_validation_error_codes(rt):
    A = rt._abi
    # out-of-range material
    b, s, light = _simple(rt)
    bad = b.sphere((0, 0, 0), 1.0, 99)
    st, msg = _status(rt, b.finish(b.list([bad]), b.list([light])))
    assert st == A.RT_ERR_BAD_ARGUMENT and "material" in msg
    # empty BVH: reference panics "no object in the scene" (bvh.rs:55)
    b, s, light = _simple(rt)
    st, msg = _status(rt, b.finish(b.bvh([]), b.list([light])))
    assert st == A.RT_ERR_EMPTY_SCENE and "no object" in msg
    # lights must be a list
    b, s, light = _simple(rt)
    st, msg = _status(rt, b.finish(b.list([s]), light))
    assert st == A.RT_ERR_BAD_ARGUMENT and "lights" in msg
    # medium inside a medium boundary is not supported
    b, s, light = _simple(rt)
    inner = b.medium(s, 0.1, b.constant_texture((1, 1, 1)))
    outer = b.medium(b.list([inner]), 0.1, b.constant_texture((1, 1, 1)))
    st, msg = _status(rt, b.finish(b.list([outer]), b.list([light])))
    assert st == A.RT_ERR_UNSUPPORTED
    # NaN parameter
    b, s, light = _simple(rt)
    nan = b.sphere((float("nan"), 0, 0), 1.0, 0)
    st, msg = _status(rt, b.finish(b.list([nan]), b.list([light])))
    assert st == A.RT_ERR_BAD_ARGUMENT
    # cycle
    b, s, light = _simple(rt)
    t = b.translate(0, (0, 0, 0))
    b.nodes[t].child = t
    st, msg = _status(rt, b.finish(b.list([t]), b.list([light])))
    assert st == A.RT_ERR_BAD_ARGUMENT and "cycle" in msg
    # abi version
    b, s, light = _simple(rt)
    sd = b.finish(b.list([s]), b.list([light]))
    sd.desc.abi_version = 77
    st, msg = _status(rt, sd)
    assert st == A.RT_ERR_BAD_ARGUMENT and "abi" in msg
    # null arguments
    assert rt._dev.rt_scene_create(None, 0, None) == A.RT_ERR_BAD_ARGUMENT


def test_valid_scene_needs_a_gpu_or_succeeds(rt):
    """A valid scene compiles; without a CUDA device creation fails loudly with RT_ERR_CUDA (no CPU path)."""
    b, s, light = _simple(rt)
    st, msg = _status(rt, b.finish(b.list([s, light]), b.list([light])))
    if rt.device_count() == 0:
        assert st == rt._abi.RT_ERR_CUDA and "no CUDA device" in msg
    else:
        assert st == 0


def test_product_does_not_reference_the_oracle():
    """The product path may not import, include, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "raytracinginrust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                text = open(path).read()
                assert not re.search(r"^\s*(import|from)\s+[^\n]*oracle", text, flags=re.M), path
                assert "liboracle" not in text and "oracle_scene_create" not in text, path
            elif f.endswith((".cpp", ".cu", ".cuh", ".h", ".hpp")):
                text = open(path, errors="ignore").read()
                assert not re.search(r"#\s*include[^\n]*oracle", text), path
                assert "oracle_" not in text, path
            elif f == "Makefile":
                assert "oracle" not in open(path).read(), path


def test_product_does_not_reference_test_infrastructure(rt):
    """tests/native builds the device source for the host so that the CPU tier can check its logic; it is not a
    CPU path of the product: no product source, Makefile or library refers to it, and bench.py never loads it."""
    pkg = os.path.join(ROOT, "raytracinginrust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp", ".inl")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "trace_on_host" not in text and "toh_" not in text, os.path.join(dirpath, f)
                assert not re.search(r"#\s*include[^\n]*tests/", text), os.path.join(dirpath, f)
    for f in ("bench.py", "include/rtb200.h"):
        text = open(os.path.join(ROOT, f)).read()
        assert "trace_on_host" not in text and "toh_" not in text and "tests/native" not in text, f
    for lib in (rt._dev, rt._host):
        for sym in ("toh_scene_create", "toh_render", "toh_trace_first_hit", "toh_path_radiance"):
            assert not hasattr(lib, sym), sym


def test_group_and_encode_argument_checks(rt):
    """The multi-GPU and output entry points validate without a GPU and never compute on the CPU."""
    b, s, light = _simple(rt)
    sd = b.finish(b.list([s, light]), b.list([light]))
    h = C.c_void_p()
    assert rt._dev.rt_scene_group_create(None, None, 0, C.byref(h)) == rt._abi.RT_ERR_BAD_ARGUMENT
    st = rt._dev.rt_scene_group_create(sd.ptr, None, 2, C.byref(h))
    if rt.device_count() == 0:
        assert st == rt._abi.RT_ERR_CUDA and "no CPU path" in rt._dev.rt_last_error().decode()
        assert not h.value
    elif st == 0:
        rt._dev.rt_scene_group_destroy(h)
    assert rt._dev.rt_scene_group_size(None) == 0 and not rt._dev.rt_scene_group_scene(None, 0)
    assert rt._dev.rt_render_multi(None, None, 4, 4, 1, 1, None, None, None) == rt._abi.RT_ERR_BAD_ARGUMENT
    n = C.c_uint64()
    assert rt._dev.rt_encode_rgb8(None, None, 4, 4, 1, None) == rt._abi.RT_ERR_BAD_ARGUMENT
    assert rt._dev.rt_encode_ppm(None, None, 4, 4, 1, None, 0, C.byref(n)) == rt._abi.RT_ERR_BAD_ARGUMENT
    # a malformed graph is reported by the group entry exactly like rt_scene_create
    b2 = rt.SceneBuilder()
    sd2 = b2.finish(0, 0)
    assert rt._dev.rt_scene_group_create(sd2.ptr, None, 1, C.byref(h)) == _status(rt, sd2)[0]


def test_light_list_members_the_convention_cannot_sample_are_refused(rt, orc):
    """A HittableList nested in the light list would be sampled through a second `choose` (hit.rs:94-96); the draw
    convention has one light index per scatter, so the library - and the oracle - refuse it instead of sampling it
    differently from the reference.  A rect light with a plane the enum does not have is a bad argument."""
    b, s, light = _simple(rt)
    nested = b.list([light])
    sd = b.finish(b.list([s, light]), b.list([b.flip(nested)]))
    st, msg = _status(rt, sd)
    assert st == rt._abi.RT_ERR_UNSUPPORTED and "nested in the light list" in msg
    with pytest.raises(orc.OracleError, match="nested in the light list"):
        orc.OracleScene(sd)
    b2, s2, _ = _simple(rt)
    bad = b2.rect(3, -1, 1, -1, 1, 5, b2.diffuse_light(b2.constant_texture((4, 4, 4))))
    good = b2.rect(rt._abi.PLANE_XZ, -1, 1, -1, 1, 5, b2.diffuse_light(b2.constant_texture((4, 4, 4))))
    st, msg = _status(rt, b2.finish(b2.list([s2, good]), b2.list([bad])))
    assert st == rt._abi.RT_ERR_BAD_ARGUMENT and "plane" in msg


def test_many_individually_transformed_instances_compile_in_linear_time(rt):
    """Every distinct Translate / Rotate chain is a chain and a group of its own; their lookup is hashed, so 4000
    instances compile in well under a second (the linear scans it replaces took seconds: quadratic in the count)."""
    import time
    b = rt.SceneBuilder()
    m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
    kids = []
    for k in range(4000):
        kids.append(b.translate(b.rotate(rt._abi.AXIS_Y, b.cube((0, 0, 0), (1, 1, 1), m), 0.01 * k), (2.0 * (k % 64), 0.0, 2.0 * (k // 64))))
    light = b.rect(rt._abi.PLANE_XZ, -1, 1, -1, 1, 50, b.diffuse_light(b.constant_texture((4, 4, 4))))
    sd = b.finish(b.list(kids + [light]), b.list([light]))
    t0 = time.perf_counter()
    blob = rt.compile_scene(sd)
    dt = time.perf_counter() - t0
    assert rt.compiled_hash(blob) != 0
    assert dt < 2.0, dt


def test_scene_create_ex_argument_checks(rt):
    b, s, light = _simple(rt)
    sd = b.finish(b.list([s, light]), b.list([light]))
    h = C.c_void_p()
    assert rt._dev.rt_scene_create_ex(sd.ptr, 0, 0x80, C.byref(h)) == rt._abi.RT_ERR_BAD_ARGUMENT
    assert "create flag" in rt._dev.rt_last_error().decode()
    st = rt._dev.rt_scene_create_ex(sd.ptr, 0, rt._abi.CREATE_GPU_BVH, C.byref(h))
    if rt.device_count() == 0:
        assert st == rt._abi.RT_ERR_CUDA  # the GPU build has no CPU stand-in either
    else:
        assert st == 0
        rt._dev.rt_scene_destroy(h)
