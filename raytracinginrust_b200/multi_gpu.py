"""Multi-GPU partition of the sample loop.

Every (pixel, sample) path is independent and the scene is read-only, so the path shards
with no data-path exchange until the end: the scene is replicated on each GPU, rank g of G
renders a contiguous block of sample indices for ALL pixels with the same Philox key
(seed, pixel, sample) — the union of samples is identical for any G — and the fp32 sum
buffers are combined by ONE reduce over NCCL/NVLink to rank 0 (a plain add, because the
buffers hold sums, not means).  One process per GPU; torch.distributed is the plumbing.
"""
import numpy as np


def sample_partition(spp, rank, world_size):
    """Contiguous, balanced split of [0, spp) over ranks: returns (sample_begin, sample_count).
    The first spp % world_size ranks get one extra sample; counts may be 0 when spp < world_size."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(int(spp), int(world_size))
    count = base + (1 if rank < extra else 0)
    begin = rank * base + min(rank, extra)
    return begin, count


def reduce_sums(local_sum, dst=0, group=None):
    """Sum the per-rank W*H*3 fp32 buffers onto rank `dst` (ncclReduce on CUDA tensors over
    NVLink/NVSwitch; gloo on CPU tensors in the tests).  In place; returns the tensor."""
    import torch.distributed as dist
    dist.reduce(local_sum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return local_sum


def render_distributed(device_scene, camera, width, height, spp, max_depth, opts_factory, rank, world_size,
                       out=None, group=None):
    """Render this rank's sample block into a CUDA tensor and reduce to rank 0.

    device_scene: raytracinginrust_b200.DeviceScene resident on this rank's GPU.
    opts_factory(sample_begin, sample_count) -> RtRenderOpts.
    Returns (tensor, stats): on rank 0 the tensor holds the full image's sums."""
    import torch
    begin, count = sample_partition(spp, rank, world_size)
    if out is None:
        out = torch.zeros((height, width, 3), dtype=torch.float32, device="cuda")
    stats = None
    if count > 0:
        stream = torch.cuda.current_stream()
        device_scene.render_device(camera, width, height, spp, max_depth, opts_factory(begin, count), out.data_ptr(),
                                   stream.cuda_stream)
        stats = device_scene.render_wait()
    else:
        out.zero_()
    if world_size > 1:
        reduce_sums(out, dst=0, group=group)
    return out, stats


def broadcast_compiled(scene_desc, rank, src=0, group=None, device=None):
    """Compile on rank `src` only (rt_compile: the graph walk and the SAH BVH build, 0.5 s of host time for the
    394k triangles of config 5) and broadcast the relocatable blob: every rank returns the same uint8 array.
    With one process per GPU each rank would otherwise repeat the compile on the same host cores.
    device: a torch device for the transfer ("cuda" -> ncclBroadcast over NVLink; None -> CPU tensors, gloo)."""
    import torch
    import torch.distributed as dist
    from . import compile_scene
    blob = compile_scene(scene_desc) if rank == src else None
    n = torch.tensor([blob.size if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src, group=group)
    if rank == src:
        t = torch.from_numpy(blob)
        t = t.to(device) if device is not None else t
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src, group=group)
    return blob if rank == src else t.cpu().numpy()


def create_scene_distributed(scene_desc, rank, local_device, src=0, group=None, device=None):
    """rt_scene_create for one-process-per-GPU hosts: one compile, one broadcast, one upload per rank.
    Returns (DeviceScene, hash of the tables) - the hash is the same on every rank by construction."""
    from . import DeviceScene, compiled_hash
    blob = broadcast_compiled(scene_desc, rank, src=src, group=group, device=device)
    return DeviceScene.from_compiled(blob, device=local_device), compiled_hash(blob)
