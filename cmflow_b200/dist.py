"""Multi-GPU plumbing: one process per GPU, frame pairs sharded by contiguous index, no data-path collective.

Frame pairs are independent in eval mode (BatchNorm uses running statistics; the only recurrent state, CMFlow-T's
256-float GRU vector, is per clip and stays rank-local), so the forward needs NO communication.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only to combine final results: max elapsed time,
total pairs, and -- when a caller wants the whole batch's outputs on every rank -- an all-gather of the small
per-pair results.  The reference's only multi-GPU mechanism is single-process nn.DataParallel (models/model.py:40-42).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return world, rank, local


def shard_bounds(n_items, world, rank):
    """Contiguous [lo, hi) of `n_items` for `rank`; the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max(value, device="cpu"):
    """max over ranks of a python float (e.g. the device-timed elapsed ms)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def reduce_sum(value, device="cpu"):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


def all_gather_shards(local, n_items):
    """Reassemble per-rank shards (leading dim = that rank's shard_bounds size) into the full (n_items, ...) tensor on
    every rank.  Shards may differ by one row; they are padded to the largest for the collective."""
    if not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_items, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], 0)


def sharded_forward(net, pc1, pc2, ft1, ft2, gather=True):
    """Run `net` (cmflow_b200.cmflow.CMFlow) on this rank's contiguous shard of a global batch held identically on every
    rank; returns the rank-local outputs, or with gather=True the full-batch (pre_trans, sf_agg, stat_cls, mask)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    B = pc1.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    sf, cls, T, mask = net(pc1[lo:hi].contiguous(), pc2[lo:hi].contiguous(), ft1[lo:hi].contiguous(), ft2[lo:hi].contiguous(), None, "test")
    if not gather:
        return sf, cls, T, mask
    return (all_gather_shards(sf, B), all_gather_shards(cls, B), all_gather_shards(T, B),
            all_gather_shards(mask.to(torch.uint8), B).bool())
