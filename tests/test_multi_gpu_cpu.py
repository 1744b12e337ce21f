"""The N>1 host logic on CPU: sample partition and the reduce of the fp32 sum buffers over a
world_size-2 gloo group.  The per-rank renderer is the oracle here (tests may use it); on the
GPU box the same partition feeds rt_render_device + ncclReduce (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_partition_covers_range():
    from raytracinginrust_b200.multi_gpu import sample_partition
    for spp in (1, 7, 8, 1000, 1024, 10000):
        for world in (1, 2, 3, 4, 8):
            parts = [sample_partition(spp, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == spp
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1
            counts = [c for _, c in parts]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sample_partition(8, 2, 2)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle_py as orc
    import raytracinginrust_b200 as rt
    from raytracinginrust_b200.multi_gpu import reduce_sums, sample_partition
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    hs = rt.HostScene("cornell")
    osc = orc.OracleScene(hs.scene_desc)
    W, H, spp, depth = 24, 20, 9, 30
    begin, count = sample_partition(spp, rank, world)
    img, _ = osc.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=3, sample_begin=begin, sample_count=count), threads=2)
    t = torch.from_numpy(img.astype(np.float32))
    reduce_sums(t, dst=0)
    if rank == 0:
        full, _ = osc.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=3), threads=2)
        np.save(out_path, np.stack([t.numpy(), full.astype(np.float32)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_equals_single_render(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got, full = np.load(out)
    # the union of the two sample blocks is the full sample set; only fp32 summation order differs
    assert np.allclose(got, full, rtol=2e-6, atol=1e-7)
    assert full.sum() > 0


def _bcast_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import raytracinginrust_b200 as rt
    from raytracinginrust_b200.multi_gpu import broadcast_compiled
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    hs = rt.HostScene("final")  # textures, Perlin tables, media, three BVHs: every table is non-empty
    blob = broadcast_compiled(hs.scene_desc, rank, src=0)
    np.save(os.path.join(out_dir, "blob%d.npy" % rank), blob)
    dist.barrier()
    dist.destroy_process_group()


def test_compile_once_broadcast_blob_over_gloo(tmp_path):
    """rank 0 compiles, the blob travels, every rank holds the same tables (same FNV hash, same bytes as a compile of
    its own would give) - the N > 1 path of rt_compile / rt_scene_create_compiled without a GPU."""
    import torch.multiprocessing as mp
    import raytracinginrust_b200 as rt
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_bcast_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = np.load(str(tmp_path / "blob0.npy")), np.load(str(tmp_path / "blob1.npy"))
    assert b0.size > 1_000_000 and np.array_equal(b0, b1)
    assert rt.compiled_hash(b0) == rt.compiled_hash(b1) != 0
    own = rt.compile_scene(rt.HostScene("final").scene_desc)
    assert np.array_equal(own, b0)  # compiling is a pure function of the description


def test_compiled_blob_is_validated():
    """rt_scene_create_compiled refuses what is not a blob of this library: truncated, corrupted, foreign bytes.  (Without a
    GPU a valid blob gets as far as 'no CUDA device'; with one it creates the scene - tests/test_gpu_parity.py.)"""
    import ctypes as C
    import raytracinginrust_b200 as rt
    hs = rt.HostScene("cornell")
    blob = rt.compile_scene(hs.scene_desc)
    assert rt.compiled_hash(blob) != 0 and rt.compiled_hash(blob[:16]) == 0

    def status(b):
        h = C.c_void_p()
        b = np.ascontiguousarray(b, dtype=np.uint8)
        st = rt._dev.rt_scene_create_compiled(b.ctypes.data_as(C.c_void_p), b.size, 0, C.byref(h))
        msg = rt._dev.rt_last_error().decode()
        if st == 0:
            rt._dev.rt_scene_destroy(h)
        return st, msg
    st, msg = status(blob)
    assert st == (0 if rt.device_count() else rt._abi.RT_ERR_CUDA), msg
    for bad, what in ((blob[:-16], "size"), (blob[:40], "truncated"), (np.zeros(4096, np.uint8), "not a blob")):
        st, msg = status(bad)
        assert st == rt._abi.RT_ERR_BAD_ARGUMENT and what in msg, (what, msg)
    flipped = blob.copy()
    flipped[-5] ^= 0x40
    st, msg = status(flipped)
    assert st == rt._abi.RT_ERR_BAD_ARGUMENT and "checksum" in msg
    # a description the compiler rejects is reported by rt_compile like by rt_scene_create
    b = rt.SceneBuilder()
    with pytest.raises(rt.RtError):
        rt.compile_scene(b.finish(0, 0))
