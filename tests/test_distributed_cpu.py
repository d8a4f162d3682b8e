"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, rank-ordered gather of uneven shards,
weighted reduction of the early-exit logs, shard-equivalence of a sharded run (injected noise AND the seed-keyed
stream), disjoint samples for consecutive seeds, and empty shards."""
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


def _keyed_noise(seed, t, row, shape):
    """Stand-in for the in-kernel Philox stream: a function of (seed, t, GLOBAL row) only."""
    g = torch.Generator().manual_seed((seed * 1000 + t) * 100003 + row)
    return torch.randn(*shape, generator=g)


def _fake_get_samples(batch_size, seed, x_T, noise_row_offset, noise, y):
    """A batch-invariant stand-in for sampler.get_samples with the same keyword contract: x_0 of a row depends only on
    that row's x_T, its injected noise (or the stream keyed by its global row index) and its label."""
    assert x_T.shape[0] == batch_size
    x = x_T.clone()
    for t in range(999, 989, -1):
        z = noise[t] if noise is not None else torch.stack(
            [_keyed_noise(seed, t, noise_row_offset + b, x.shape[1:]) for b in range(batch_size)])
        x = 0.9 * x + 0.1 * z + (0 if y is None else y.view(-1, 1, 1, 1).float() * 1e-3)
    return x.permute(0, 2, 3, 1).contiguous().numpy(), []


def _worker(rank, world, port, G, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = (3, 4, 4)
        g = torch.Generator().manual_seed(0)
        noise = torch.randn(1000, G, *shape, generator=g)
        y = torch.arange(G)
        # (1) injected noise: sharded == single process, row for row
        got = D.get_samples_sharded(_fake_get_samples, G, shape=shape, noise=noise, y=y, seed=5)
        ref, _ = _fake_get_samples(G, 5, D.global_x_T(5, G, shape), 0, noise, y)
        assert torch.equal(got, torch.from_numpy(ref)), "sharded run differs from the single-process run"
        # (2) seed-keyed stream (no injected noise): still the single-process result, for every seed; consecutive
        # seeds give different samples on every row (the old `seed + rank` derivation repeated them across calls)
        outs = []
        for seed in (0, 1, 2):
            got = D.get_samples_sharded(_fake_get_samples, G, shape=shape, seed=seed)
            ref, _ = _fake_get_samples(G, seed, D.global_x_T(seed, G, shape), 0, None, None)
            assert torch.equal(got, torch.from_numpy(ref)), seed
            outs.append(got)
        for a in range(3):
            for b in range(a + 1, 3):
                d = (outs[a][:, None] - outs[b][None]).flatten(2).abs().amax(-1)  # [G, G] row-to-row distance
                assert d.min().item() > 1e-3, f"seeds {a} and {b} share a sample"
        # (3) early-exit logs
        lo, hi = D.shard_bounds(G, rank, world)
        scores = torch.arange(G, dtype=torch.float32).view(1, G).repeat(5, 1)  # [depth, G]
        m = D.all_reduce_weighted_mean(scores[:, lo:hi].mean(1), hi - lo, G)
        assert torch.allclose(m, scores.mean(1))
        idx = D.all_gather_rows(torch.arange(lo, hi).view(-1, 1), G)
        assert torch.equal(idx.flatten(), torch.arange(G))
        idx_log = torch.arange(1000 * G, dtype=torch.float32).view(1000, G)
        err_log = idx_log.mean(1, keepdim=True).repeat(1, 5)
        e, i = D.gather_ee_logs(idx_log[:, lo:hi].mean(1, keepdim=True).repeat(1, 5), idx_log[:, lo:hi].contiguous(), G)
        assert torch.equal(i, idx_log) and torch.allclose(e, err_log)
        # (4) more ranks than samples: the empty shard contributes no rows
        got1 = D.get_samples_sharded(_fake_get_samples, 1, shape=shape, seed=3)
        ref1, _ = _fake_get_samples(1, 3, D.global_x_T(3, 1, shape), 0, None, None)
        assert torch.equal(got1, torch.from_numpy(ref1))
        if rank == 0:
            open(tmp, "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_equivalence(tmp_path):
    flag = tmp_path / "ok"
    mp.spawn(_worker, args=(2, _free_port(), 7, str(flag)), nprocs=2, join=True)  # 7 rows: uneven shards (4 + 3)
    assert flag.read_text() == "ok"
