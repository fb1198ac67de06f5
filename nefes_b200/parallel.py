"""Multi-GPU plumbing for the two places the path shards.

Training: rays of a step are split across ranks, weight replicas are identical, ONE all-reduce
(SUM) over the flat gradient buffers per step, then Adam with grad_scale = 1/world.
Refinement / full renders: queries (or contiguous ray ranges) are split across ranks with no
collective until a final gather.  One process per GPU, torch.distributed (NCCL on GPUs; the same
code runs on gloo for the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous [lo, hi) slice of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_strided(n: int, rank: int, world_size: int):
    """Query ids r, r+W, r+2W, ... (refinement: DFM_APR_refine.py:204 loop, one query at a time)."""
    return list(range(rank, n, world_size))


def allreduce_grads(params, bucket: torch.Tensor | None = None):
    """Sum gradients over ranks.  All gradients are packed into ONE buffer so a step costs a single
    collective (2.8 MB for coarse+fine: latency-bound, so one launch beats many).  Returns the
    number of elements reduced.  The caller divides by world size (FlatAdam grad_scale)."""
    rank, ws = world()
    grads = [p.grad for p in params if p.grad is not None]
    if ws == 1 or not grads:
        return 0
    if len(grads) == 1:
        dist.all_reduce(grads[0], op=dist.ReduceOp.SUM)
        return grads[0].numel()
    n = sum(g.numel() for g in grads)
    if bucket is None or bucket.numel() < n:
        bucket = torch.empty(n, dtype=grads[0].dtype, device=grads[0].device)
    off = 0
    for g in grads:
        bucket[off:off + g.numel()].copy_(g.reshape(-1))
        off += g.numel()
    dist.all_reduce(bucket[:n], op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        g.copy_(bucket[off:off + g.numel()].view_as(g))
        off += g.numel()
    return n


def gather_rows(local: torch.Tensor, ids, n_total: int):
    """Final gather of per-query results (e.g. refined poses [n_local, 12]) onto every rank,
    placed at their global query ids."""
    rank, ws = world()
    out = torch.zeros((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if ws == 1:
        out[torch.as_tensor(ids, device=local.device)] = local
        return out
    counts = [len(shard_strided(n_total, r, ws)) for r in range(ws)]
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    for r in range(ws):
        idx = torch.as_tensor(shard_strided(n_total, r, ws), device=local.device, dtype=torch.long)
        out[idx] = bufs[r][:counts[r]]
    return out
