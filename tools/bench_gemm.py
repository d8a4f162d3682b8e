"""GEMM micro-benchmark on the U-ViT shapes (CUDA events, 20 launches each, inputs > L2 rotated).
    python tools/bench_gemm.py [--variants 2,1] [--debug 0,1,2,4]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from duodiff_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="2,1")
ap.add_argument("--debug", default="0")
ap.add_argument("--M", type=int, default=257 * 128)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--only", default="")
ap.add_argument("--pdl", type=int, default=1)
ap.add_argument("--bn128", type=int, default=0)
ap.add_argument("--ts", type=int, default=1)
ap.add_argument("--n", type=int, default=20)
ap.add_argument("--lncfg", type=int, default=0)
a = ap.parse_args()
L = _lib.load()
_lib.check(L.ddb_set_option(b"pdl", a.pdl))
_lib.check(L.ddb_set_option(b"gemm_bn128", a.bn128))
_lib.check(L.ddb_set_option(b"gemm_ts", a.ts))
_lib.check(L.ddb_set_option(b"gemm_ln_cfg", a.lncfg))
dev = torch.device("cuda:0")
M, D = a.M, a.D
shapes = [("qkv", 3 * D, D, 0, 1), ("proj", D, D, 0, 3), ("fc1", 4 * D, D, 0, 2), ("fc2", D, 4 * D, 0, 3),
          ("skip", D, D, D, 0)]
NBUF = 3
for name, N, K0, K1, epi in shapes:
    if a.only and name not in a.only.split(","):
        continue
    K = K0 + K1
    a0 = [torch.randn(M, K0, device=dev).bfloat16() for _ in range(NBUF)]
    a1 = [torch.randn(M, K1, device=dev).bfloat16() for _ in range(NBUF)] if K1 else [None] * NBUF
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    colsum = w.float().sum(1).contiguous()
    NP = K // 64  # partial statistics per row, as written by the producing GEMM's epilogue
    stats = torch.stack([torch.zeros(M, NP, device=dev), torch.full((M, NP), 64.0, device=dev)], 2).contiguous()
    res = [torch.randn(M, N, device=dev).bfloat16() for _ in range(NBUF)] if epi == 3 else [None] * NBUF
    out = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NBUF)]
    sout = torch.empty(M, N // 64, 2, device=dev)
    for variant in [int(v) for v in a.variants.split(",")]:
        for dbg in [int(v) for v in a.debug.split(",")]:
            if dbg and variant != 2:
                continue
            _lib.check(L.ddb_set_option(b"gemm_debug", dbg))

            def run(i):
                so = sout if (variant == 2 and epi in (0, 3)) else None
                _lib.check(L.ddb_op_gemm(_lib.ptr(a0[i % NBUF]), _lib.ptr(a1[i % NBUF]), _lib.ptr(w), _lib.ptr(bias),
                                         _lib.ptr(colsum), _lib.ptr(stats), NP, K, _lib.ptr(res[i % NBUF]),
                                         _lib.ptr(out[i % NBUF]), _lib.ptr(so), M, N, K0, K1, epi, variant,
                                         _lib.current_stream_ptr()))
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = a.n
            e0.record()
            for i in range(n):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            print(f"{name:5s} N={N:5d} K={K:5d} epi={epi} variant={variant} debug={dbg}: {us:7.1f} us "
                  f"{2 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
    _lib.check(L.ddb_set_option(b"gemm_debug", 0))
