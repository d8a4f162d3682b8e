"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, rank-ordered gather of uneven shards,
weighted reduction of the early-exit logs, and shard-equivalence of a sharded run with injected noise."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from duodiff_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_batch_exactly():
    for g in (1, 7, 128, 257, 1000):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(g, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == g
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_get_samples(batch_size, seed, noise, y, x_T_global, lo_hi):
    """A batch-invariant stand-in for the sampler: x_0 depends only on each row's own x_T and noise."""
    lo, hi = lo_hi[seed]  # seed = base + rank -> this rank's rows
    x = x_T_global[lo:hi].clone()
    for t in range(999, 989, -1):
        x = 0.9 * x + 0.1 * noise[t] + (0 if y is None else y.view(-1, 1, 1, 1).float() * 1e-3)
    return x.permute(0, 2, 3, 1).contiguous().numpy(), []


def _worker(rank, world, port, G, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x_T = torch.randn(G, 3, 4, 4, generator=g)
        noise = torch.randn(1000, G, 3, 4, 4, generator=g)
        y = torch.arange(G)
        lo_hi = {r: D.shard_bounds(G, r, world) for r in range(world)}
        got = D.get_samples_sharded(_fake_get_samples, G, noise=noise, y=y, seed=0, x_T_global=x_T, lo_hi=lo_hi)
        ref, _ = _fake_get_samples(G, 0, noise, y, x_T, {0: (0, G)})
        assert torch.equal(got, torch.from_numpy(ref)), "sharded run differs from the single-process run"
        # early-exit logs
        lo, hi = lo_hi[rank]
        scores = torch.arange(G, dtype=torch.float32).view(1, G).repeat(5, 1)  # [depth, G]
        m = D.all_reduce_weighted_mean(scores[:, lo:hi].mean(1), hi - lo, G)
        assert torch.allclose(m, scores.mean(1))
        idx = D.all_gather_rows(torch.arange(lo, hi).view(-1, 1), G)
        assert torch.equal(idx.flatten(), torch.arange(G))
        if rank == 0:
            open(tmp, "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_equivalence(tmp_path):
    flag = tmp_path / "ok"
    mp.spawn(_worker, args=(2, _free_port(), 7, str(flag)), nprocs=2, join=True)  # 7 rows: uneven shards (4 + 3)
    assert flag.read_text() == "ok"
