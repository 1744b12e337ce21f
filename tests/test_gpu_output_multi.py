"""GPU tests of what surrounds the render kernels: the single-process multi-GPU entry
(rt_scene_group_create / rt_render_multi) and format_color + the P3 text on the device
(rt_encode_rgb8 / rt_encode_ppm), checked against the host writer and the oracle."""
import io
import os
import subprocess

import numpy as np
import pytest

from util import host_scene

pytestmark = pytest.mark.gpu


def host_ppm_bytes(rt, img, spp, tmp_path):
    path = str(tmp_path / "host.ppm")
    rt.write_ppm(path, img, spp)
    return open(path, "rb").read()


def adversarial_sums(h, w, spp, seed):
    """fp32 sums that exercise every branch of format_color (vec.rs:125-131) and every line length."""
    rng = np.random.default_rng(seed)
    img = (rng.uniform(0, 1.3, size=(h, w, 3)) ** 2 * spp).astype(np.float32)
    flat = img.reshape(-1)
    flat[rng.integers(0, flat.size, flat.size // 7)] = 0.0  # one-digit channels
    k = rng.integers(0, flat.size, 64)
    flat[k[:8]] = np.nan
    flat[k[8:16]] = np.inf
    flat[k[16:24]] = -np.inf
    flat[k[24:32]] = -1.0
    flat[k[32:40]] = -0.0
    flat[k[40:48]] = np.float32(3.0e38)
    flat[k[48:56]] = np.float32(1e-45)  # denormal
    # values next to the 256 quantisation boundaries: sum = spp * (c/256)^2, one ulp either side
    c = rng.integers(1, 256, 4096).astype(np.float64)
    edge = (spp * (c / 256.0) ** 2).astype(np.float32)
    edge = np.nextafter(edge, np.where(rng.integers(0, 2, edge.size) == 0, np.float32(0), np.float32(np.inf))).astype(np.float32)
    pos = rng.integers(0, flat.size, edge.size)
    flat[pos] = edge
    return img


@pytest.mark.parametrize("shape,spp", [((2, 2), 1), ((37, 53), 800), ((64, 64), 1000), ((601, 333), 10000)])
def test_encode_matches_host_writer(rt, orc, tmp_path, shape, spp):
    """rt_encode_rgb8 / rt_encode_ppm on caller device memory: byte-identical to the host's format_color
    and PPM writer (and so to the oracle's format_color), ragged sizes and every special value included."""
    import torch
    h, w = shape
    img = adversarial_sums(h, w, spp, seed=h * 1000 + w)
    hs = host_scene(rt, "cornell")
    dev = rt.DeviceScene(hs.scene_desc, device=0)
    t = torch.from_numpy(img).cuda()
    rgb8 = dev.encode_rgb8(w, h, spp, sums_ptr=t.data_ptr())
    assert np.array_equal(rgb8, rt.format_image(img, spp))
    assert np.array_equal(rgb8, orc.format_image(img.astype(np.float64), spp))
    ppm = dev.encode_ppm(w, h, spp, sums_ptr=t.data_ptr())
    assert ppm == host_ppm_bytes(rt, img, spp, tmp_path)
    assert ppm.startswith(b"P3\n%d %d\n255\n" % (w, h)) and ppm.count(b"\n") == 3 + w * h
    # a buffer that is too small is refused with the size it would have taken
    with pytest.raises(rt.RtError) as e:
        dev.encode_ppm(w, h, spp, sums_ptr=t.data_ptr(), capacity=len(ppm) - 1)
    assert e.value.status == rt._abi.RT_ERR_BAD_ARGUMENT
    dev.close()


def test_encode_resident_image_after_render(rt, orc, tmp_path):
    """rt_render with a NULL host buffer leaves the image on the device; the encoders read it there."""
    hs = host_scene(rt, "cornell")
    dev = rt.DeviceScene(hs.scene_desc, device=0)
    W, H, spp = 83, 47, 16
    opts = rt.render_opts(seed=3, integrator=hs.integrator)
    with pytest.raises(rt.RtError):
        dev.encode_rgb8(W, H, spp)  # nothing resident yet
    img, _ = dev.render(hs.camera, W, H, spp, 50, opts)
    none, stats = dev.render(hs.camera, W, H, spp, 50, opts, want_sums=False)
    assert none is None and stats.paths == W * H * spp and stats.d2h_bytes < 1024
    assert np.array_equal(dev.encode_rgb8(W, H, spp), rt.format_image(img, spp))
    assert dev.encode_ppm(W, H, spp) == host_ppm_bytes(rt, img, spp, tmp_path)
    with pytest.raises(rt.RtError):
        dev.encode_ppm(W + 1, H, spp)  # not the size of the resident image
    dev.close()


def manual_blocks(rt, dev, hs, W, H, spp, depth, seed, n):
    """What rt_render_multi must equal bit for bit: the n sample blocks rendered one after the other
    and added in fp32 in block order."""
    from raytracinginrust_b200.multi_gpu import sample_partition
    acc = None
    for r in range(n):
        b, c = sample_partition(spp, r, n)
        if c == 0:
            continue
        part, _ = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=seed, integrator=hs.integrator,
                                                                        sample_begin=b, sample_count=c))
        acc = part.copy() if acc is None else (acc + part).astype(np.float32)
    return acc


@pytest.mark.parametrize("name,n", [("cornell", 3), ("final", 2), ("cornell", 5)])
def test_render_multi_on_one_device_equals_block_sum(rt, orc, name, n):
    """A group that lists device 0 n times runs the whole multi-GPU path (per-scene streams, events, the
    combine kernel) on a one-GPU box."""
    hs = host_scene(rt, name)
    W, H, spp, depth = 96, 64, 4 if n == 5 else 24, 50   # n = 5 with 4 samples: one member gets no block
    group = rt.SceneGroup(hs.scene_desc, devices=[0] * n)
    assert group.size == n
    img, stats = group.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=9, integrator=hs.integrator))
    assert stats.paths == W * H * spp and stats.render_ms > 0
    dev = rt.DeviceScene(hs.scene_desc, device=0)
    assert np.array_equal(img, manual_blocks(rt, dev, hs, W, H, spp, depth, 9, n))
    full, fs = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=9, integrator=hs.integrator))
    assert fs.rays == stats.rays
    assert np.abs(img - full).max() <= 1e-5 * max(1.0, float(np.abs(full).max()))
    # the combined image stays on the root for the encoders
    assert np.array_equal(group.scene(0).encode_rgb8(W, H, spp), rt.format_image(img, spp))
    # a sub-range of the samples is partitioned the same way
    sub, ss = group.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=9, integrator=hs.integrator, sample_begin=1, sample_count=2))
    assert ss.paths == W * H * 2
    dev.close()
    group.close()


def test_render_multi_all_devices(rt, orc):
    """Every GPU of the box (peer mappings over NVLink where available; the staged path forced as well)."""
    n = rt.device_count()
    if n < 2:
        pytest.skip("one GPU: covered by test_render_multi_on_one_device_equals_block_sum")
    hs = host_scene(rt, "cornell_smoke")
    W, H, spp, depth = 128, 96, 8 * n + 3, 50
    dev = rt.DeviceScene(hs.scene_desc, device=0)
    want = manual_blocks(rt, dev, hs, W, H, spp, depth, 2, n)
    for no_peer in ("", "1"):
        if no_peer:
            os.environ["RTB200_NO_PEER"] = "1"
        try:
            group = rt.SceneGroup(hs.scene_desc)  # all devices
            assert group.size == n
            img, stats = group.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=2, integrator=hs.integrator))
            assert stats.paths == W * H * spp
            assert np.array_equal(img, want)
            group.close()
        finally:
            os.environ.pop("RTB200_NO_PEER", None)
    dev.close()


def test_host_render_ppm_and_cli(rt, orc, tmp_path):
    """render() -> P3 through the host layer with the text made on the GPU equals the host writer on the
    sums of the plain entry; the CLI prints the same file."""
    hs = host_scene(rt, "cornell")
    W, H, spp, depth = 40, 30, 8, 20
    opts = rt.render_opts(seed=1, integrator=hs.integrator)
    img, _ = hs.render(W, H, spp, depth, opts)
    ppm, stats = hs.render_ppm(W, H, spp, depth, opts, n_gpus=1)
    assert ppm == host_ppm_bytes(rt, img, spp, tmp_path)
    assert stats.paths == W * H * spp
    exe = os.path.join(rt.LIB_DIR, "rtb200_render")
    args = [exe, "--scene", "cornell", "--width", str(W), "--height", str(H), "--spp", str(spp), "--depth", str(depth),
            "--assets", rt.ASSETS_DIR]
    a = subprocess.run(args, capture_output=True, timeout=300)
    b = subprocess.run(args + ["--host-ppm"], capture_output=True, timeout=300)
    c = subprocess.run(args + ["--gpus", "0"], capture_output=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0 and c.returncode == 0, (a.stderr, b.stderr, c.stderr)
    assert a.stdout == ppm and b.stdout == ppm
    if rt.device_count() == 1:
        assert c.stdout == ppm


def test_one_render_in_flight_per_scene(rt):
    """A second rt_render_device (or rt_render) before rt_render_wait is refused: the scene's planes, counters and
    wavefront pool are still being written by the first (include/rtb200.h)."""
    import torch
    hs = host_scene(rt, "cornell")
    dev = rt.DeviceScene(hs.scene_desc, device=0)
    opts = rt.render_opts(seed=2, integrator=hs.integrator)
    out = torch.zeros((32, 32, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    dev.render_device(hs.camera, 32, 32, 64, 20, opts, out.data_ptr(), stream)
    with pytest.raises(rt.RtError, match="render pending"):
        dev.render_device(hs.camera, 32, 32, 64, 20, opts, out.data_ptr(), stream)
    with pytest.raises(rt.RtError, match="render pending"):
        dev.render(hs.camera, 32, 32, 64, 20, opts)
    st = dev.render_wait()
    assert st.paths == 32 * 32 * 64
    first = out.clone()
    dev.render_device(hs.camera, 32, 32, 64, 20, opts, out.data_ptr(), stream)  # accepted again; the image is overwritten
    dev.render_wait()
    torch.cuda.synchronize()
    assert torch.equal(first, out)


@pytest.mark.parametrize("name", ["cornell", "final", "mesh"])
def test_scene_from_compiled_blob_renders_the_same_image(rt, name):
    """rt_compile + rt_scene_create_compiled (what a rank does with the blob another rank compiled) against
    rt_scene_create: same tables, so the same image bit for bit and the same counters."""
    hs = host_scene(rt, name)
    blob = rt.compile_scene(hs.scene_desc)
    a = rt.DeviceScene(hs.scene_desc, device=0)
    b = rt.DeviceScene.from_compiled(blob, device=0)
    assert a.device_bytes == b.device_bytes
    opts = rt.render_opts(seed=3, integrator=hs.integrator)
    ia, sa = a.render(hs.camera, 64, 48, 8, 30, opts)
    ib, sb = b.render(hs.camera, 64, 48, 8, 30, opts)
    assert np.array_equal(ia, ib, equal_nan=True)
    assert (sa.paths, sa.rays) == (sb.paths, sb.rays)
    assert a.render_info == b.render_info
    a.close()
    b.close()


def test_device_blocks_are_recycled_between_scenes(rt):
    """The sample planes, the wavefront pool and the tables of a destroyed scene are parked and handed to the next scene
    (api.cu ScratchCache): a render in a recycled block - whatever the previous scene left in it - gives the same image,
    for the megakernel and the wavefront pipeline, and rt_release_cached_memory returns the blocks to the driver."""
    import torch

    def image(name, w, h, spp):
        hs = host_scene(rt, name)
        dev = rt.DeviceScene(hs.scene_desc, device=0)
        img, st = dev.render(hs.camera, w, h, spp, 50, rt.render_opts(seed=4, integrator=hs.integrator))
        assert st.paths == w * h * spp
        dev.close()
        return img

    rt.release_cached_memory()
    first = {n: image(n, 256, 192, 24) for n in ("cornell", "final", "mesh")}
    image("cornell_smoke", 320, 320, 16)   # another scene with other sizes scribbles over what it is given
    for n, want in first.items():
        assert np.array_equal(image(n, 256, 192, 24), want), n
    torch.cuda.synchronize()
    before = torch.cuda.mem_get_info()[0]
    rt.release_cached_memory()
    parked = torch.cuda.mem_get_info()[0] - before
    print("parked blocks returned to the driver: %.1f MB" % (parked / 1e6))
    assert parked >= 64 << 20              # the planes and the wavefront pool of the scenes above were parked
    assert np.array_equal(image("cornell", 256, 192, 24), first["cornell"])
    rt.release_cached_memory()
