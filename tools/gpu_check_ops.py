"""Operator-level GPU diagnostics: every CUDA kernel against a torch fp32 reference on identical bf16 inputs.
Run on the GPU box:  python tools/gpu_check_ops.py [group ...]   (groups: gemm attn misc)"""
import sys
import json
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from duodiff_b200 import _lib  # noqa: E402

L = _lib.load()
dev = torch.device("cuda:0")
RESULTS = []


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
    ok = bool(torch.isfinite(got).all().item()) and rel < tol
    r = dict(name=name, max_abs=err.max().item(), ref_max=denom, rel_l2=rel, ok=ok)
    RESULTS.append(r)
    print(json.dumps(r), flush=True)
    if not ok and got.dim() == 2:
        bad = (err > 0.05 * denom)
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"   bad rows: n={rows.numel()} first={rows[:12].tolist()} last={rows[-4:].tolist()}")
        print(f"   bad cols: n={cols.numel()} first={cols[:12].tolist()} last={cols[-4:].tolist()}")
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
    return ok


def gemm_case(M, N, K0, K1, epi, seed=0, variant=2):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a0 = (torch.randn(M, K0, generator=g) * 1.0 + 0.3).to(dev).bfloat16()
    a1 = torch.randn(M, K1, generator=g).to(dev).bfloat16() if K1 else None
    K = K0 + K1
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, generator=g).to(dev).bfloat16() if epi == 3 else None
    A = a0.float() if a1 is None else torch.cat([a0.float(), a1.float()], 1)
    colsum = stats = None
    if epi in (1, 2):
        colsum = w.float().sum(1).contiguous()
        mean = A.mean(1)
        m2 = ((A - mean[:, None]) ** 2).sum(1)
        stats = torch.stack([mean, m2], 1).contiguous()
    want_stats = variant == 2 and epi in (0, 3)
    sout = torch.zeros(M, N // 64, 2, device=dev) if want_stats else None
    _lib.check(L.ddb_op_gemm(_lib.ptr(a0), _lib.ptr(a1), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(colsum),
                             _lib.ptr(stats), 1, K, _lib.ptr(res), _lib.ptr(out), _lib.ptr(sout), M, N, K0, K1, epi,
                             variant, _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    acc = A @ w.float().t()
    if epi == 0:
        ref = acc + bias
    elif epi in (1, 2):
        rstd = torch.rsqrt(m2 / K + 1e-5)
        ref = (acc - mean[:, None] * colsum[None]) * rstd[:, None] + bias
        if epi == 2:
            ref = torch.nn.functional.gelu(ref)
    else:
        ref = acc + bias + res.float()
    ok = report(f"gemm v{variant} M={M} N={N} K0={K0} K1={K1} epi={epi}", out, ref, 1e-2)
    if want_stats:
        chunks = ref.view(M, N // 64, 64)
        mean = chunks.mean(-1)
        m2 = ((chunks - mean[..., None]) ** 2).sum(-1)
        ok &= report("   chunk mean", sout[..., 0], mean, 1e-4)
        ok &= report("   chunk M2", sout[..., 1], m2, 1e-3)
    return ok


def group_gemm():
    ok = True
    ok &= gemm_case(128, 256, 64, 0, 0, variant=1)
    ok &= gemm_case(1000, 512, 512, 0, 3, variant=1)
    ok &= gemm_case(128, 256, 64, 0, 0)
    ok &= gemm_case(256, 256, 64, 0, 0)
    ok &= gemm_case(128, 256, 512, 0, 0)
    ok &= gemm_case(300, 512, 512, 0, 0)
    ok &= gemm_case(1000, 512, 512, 512, 0)
    for epi in (1, 2, 3):
        ok &= gemm_case(1000, 512, 512, 0, epi)
    ok &= gemm_case(257 * 16, 1536, 512, 0, 1)
    ok &= gemm_case(257 * 16, 2048, 512, 0, 2)
    ok &= gemm_case(257 * 16, 512, 2048, 0, 3)
    ok &= gemm_case(257 * 128, 512, 512, 0, 3)
    ok &= gemm_case(257 * 128, 1536, 512, 0, 1)
    return ok


def group_attn():
    ok = True
    for (B, Lq, H, variant) in [(1, 257, 1, 2), (2, 257, 8, 2), (3, 258, 12, 2), (128, 257, 8, 2), (2, 257, 8, 1),
                                (2, 100, 2, 1)]:
        g = torch.Generator(device="cpu").manual_seed(B * 7 + H)
        D = H * 64
        qkv = (torch.randn(B * Lq, 3 * D, generator=g) * 1.5).to(dev).bfloat16()
        out = torch.zeros(B * Lq, D, device=dev, dtype=torch.bfloat16)
        _lib.check(L.ddb_op_attention(_lib.ptr(qkv), _lib.ptr(out), B, Lq, H, variant, _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        x = qkv.float().view(B, Lq, 3, H, 64).permute(2, 0, 3, 1, 4)
        q, k, v = x[0], x[1], x[2]
        att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
        ref = att.permute(0, 2, 1, 3).reshape(B * Lq, D)
        ok &= report(f"attention B={B} L={Lq} H={H} variant={variant}", out, ref, 1e-2)
        if variant == 2:
            xt = out.float().view(B, Lq, D)
            rt = ref.view(B, Lq, D)
            ex = Lq - 256
            print("   extras rows rel", ((xt[:, :ex] - rt[:, :ex]).norm() / rt[:, :ex].norm()).item(),
                  " patch rows rel", ((xt[:, ex:] - rt[:, ex:]).norm() / rt[:, ex:].norm()).item(), flush=True)
    return ok


def group_misc():
    ok = True
    for D in (512, 768, 1024):
        x = (torch.randn(777, D, device=dev) * 2 + 1).bfloat16()
        stats = torch.zeros(777, 2, device=dev)
        _lib.check(L.ddb_op_ln_stats(_lib.ptr(x), 777, D, _lib.ptr(stats), _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        xf = x.float()
        mean = xf.mean(1)
        m2 = ((xf - mean[:, None]) ** 2).sum(1)
        ok &= report(f"ln_stats D={D}", stats, torch.stack([mean, m2], 1), 1e-5)
    return ok


if __name__ == "__main__":
    groups = sys.argv[1:] or ["gemm", "attn", "misc"]
    print(L.ddb_version().decode(), torch.cuda.get_device_name(0))
    allok = True
    for gname in groups:
        try:
            allok &= {"gemm": group_gemm, "attn": group_attn, "misc": group_misc}[gname]()
        except Exception as e:  # noqa: BLE001
            print(f"GROUP {gname} FAILED: {type(e).__name__}: {e}", flush=True)
            allok = False
    print("ALL_OK" if allok else "SOME_FAILED")
    sys.exit(0 if allok else 1)
