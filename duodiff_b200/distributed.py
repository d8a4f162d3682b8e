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


def global_x_T(seed: int, global_batch: int, shape) -> torch.Tensor:
    """The initial noise of the whole global batch, drawn like the reference's single process does (sampler.py:99-100:
    seed_everything(seed), then one CPU randn of the full batch).  Every rank draws the same tensor and keeps its rows."""
    from ._io import seed_everything
    seed_everything(seed)
    return torch.randn(global_batch, *shape)


def get_samples_sharded(get_samples_fn, global_batch: int, *, shape=None, noise=None, y=None, seed: int = 0,
                        x_T=None, **kw):
    """Run `sampler.get_samples`-like `get_samples_fn(batch_size=..., seed=..., x_T=..., noise_row_offset=..., noise=...,
    y=...)` on this rank's shard of a global batch and gather the finished samples on every rank.

    The result is the single-process result of `get_samples(batch_size=global_batch, seed=seed)` row for row, whatever
    the world size: every rank draws the global x_T (`shape` = [C,H,W]; or is handed `x_T` [G,C,H,W]) and keeps rows
    [lo, hi); the per-step noise is either injected (`noise` [1000,G,C,H,W]) or the in-kernel Philox stream keyed by
    (seed, t, global element index) through `noise_row_offset=lo`; the kernels are batch-invariant.  So a caller
    looping over seeds gets the same, disjoint sample sets as on one GPU (no `seed + rank` collisions).  A rank whose
    shard is empty (world > global_batch) skips the computation and contributes no rows."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(global_batch, rank, world)
    if x_T is None:
        if shape is None:
            raise ValueError("get_samples_sharded needs shape=[C,H,W] (to draw the global x_T) or x_T")
        x_T = global_x_T(seed, global_batch, tuple(shape))
    x_T = torch.as_tensor(x_T)
    use_cuda = dist.is_initialized() and dist.get_backend() == "nccl"
    if hi > lo:
        n_local = noise[:, lo:hi].contiguous() if noise is not None else None
        y_local = y[lo:hi].contiguous() if y is not None else None
        out = get_samples_fn(batch_size=hi - lo, seed=seed, x_T=x_T[lo:hi].contiguous(), noise_row_offset=lo,
                             noise=n_local, y=y_local, **kw)
        samples = out[0] if isinstance(out, tuple) else out
        t = torch.as_tensor(samples)
    else:
        c, h, w = x_T.shape[1:]
        t = torch.zeros(0, h, w, c)
    if use_cuda:
        t = t.cuda()
    return all_gather_rows(t, global_batch)


def gather_ee_logs(error_prediction_local: torch.Tensor, indices_local: torch.Tensor, global_rows: int):
    """eesampler logs of the shards -> the logs of the global batch (eesampler.py:54-55,71-72):
    ``indices_by_timestep`` [1000, B_local] is gathered along the batch axis in rank order, and
    ``error_prediction_by_timestep`` [1000, depth] (a batch mean) is combined as a row-count-weighted mean."""
    local_rows = indices_local.shape[1]
    idx = all_gather_rows(indices_local.t().contiguous(), global_rows).t().contiguous()
    err = all_reduce_weighted_mean(error_prediction_local, local_rows, global_rows)
    return err, idx
