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
