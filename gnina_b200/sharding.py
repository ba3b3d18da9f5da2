"""Pose sharding across GPUs: poses are independent, so each rank scores a contiguous range with
replicated weights + receptor and the only collective is ONE all_gather of the per-pose results at the end
(NCCL on GPUs, gloo in the CPU tests).  The reference has no multi-GPU path at all."""
import numpy as np


def shard_range(n_items, rank, world):
    """contiguous, balanced: the first (n % world) ranks get one extra item"""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def slice_poses(lig_xyz, lig_types, pose_offsets, lo, hi):
    off = np.asarray(pose_offsets)
    a, b = int(off[lo]), int(off[hi])
    return lig_xyz[a:b], lig_types[a:b], (off[lo:hi + 1] - off[lo]).astype(np.int32)


def score_sharded(score_fn, lig_xyz, lig_types, pose_offsets, centers=None, group=None, device=None):
    """score_fn(xyz, types, offsets, centers) -> tuple of K float arrays (one value per pose).
    Returns the K arrays for ALL poses on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return score_fn(lig_xyz, lig_types, pose_offsets, centers)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(pose_offsets) - 1
    lo, hi = shard_range(n, rank, world)
    xs, ts, offs = slice_poses(lig_xyz, lig_types, pose_offsets, lo, hi)
    outs = score_fn(xs, ts, offs, None if centers is None else centers[lo:hi])
    k = len(outs)
    per = (n + world - 1) // world                     # equal-sized shards for the gather, tail padded
    local = torch.zeros(k, per, dtype=torch.float32)
    for i, o in enumerate(outs):
        local[i, : hi - lo] = torch.as_tensor(np.asarray(o, np.float32))
    if device is not None:
        local = local.to(device)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local, group=group)
    full = [np.empty(n, np.float32) for _ in range(k)]
    for r in range(world):
        rlo, rhi = shard_range(n, r, world)
        g = gathered[r].cpu().numpy()
        for i in range(k):
            full[i][rlo:rhi] = g[i, : rhi - rlo]
    return tuple(full)
