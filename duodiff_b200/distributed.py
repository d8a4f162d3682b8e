"""Data-parallel sharding of the sample batch: one process per GPU, no collective inside the loop,
one all-gather of the finished samples at the end (plus the early-exit logs).  Works over NCCL (GPU) and gloo (CPU
tests of the host logic)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_bounds(global_batch: int, rank: int, world: int):
    """Rows [lo, hi) of the global batch owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local: torch.Tensor, global_rows: int) -> torch.Tensor:
    """Concatenate per-rank row blocks (possibly uneven) in rank order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(global_rows, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_rows, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def all_reduce_weighted_mean(local_mean: torch.Tensor, local_rows: int, global_rows: int) -> torch.Tensor:
    """Batch-mean logs (eesampler.py:71) of the shards -> the mean over the global batch."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_mean
    acc = local_mean * float(local_rows)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc / float(global_rows)


def get_samples_sharded(get_samples_fn, global_batch: int, *, noise=None, y=None, seed: int = 0, **kw):
    """Run `sampler.get_samples`-like `get_samples_fn(batch_size=..., seed=..., noise=..., y=...)` on this rank's
    shard of a global batch and gather the finished samples.  With injected noise [1000, G, C, H, W] and an injected
    x_T the result is row-for-row the single-process result (the kernels are batch-invariant)."""
    rank, world, _ = (dist.get_rank(), dist.get_world_size(), 0) if dist.is_initialized() else (0, 1, 0)
    lo, hi = shard_bounds(global_batch, rank, world)
    n_local = noise[:, lo:hi].contiguous() if noise is not None else None
    y_local = y[lo:hi].contiguous() if y is not None else None
    out = get_samples_fn(batch_size=hi - lo, seed=seed + rank, noise=n_local, y=y_local, **kw)
    samples = out[0] if isinstance(out, tuple) else out
    t = torch.as_tensor(samples)
    if dist.is_initialized() and dist.get_backend() == "nccl":
        t = t.cuda()
    return all_gather_rows(t, global_batch)
