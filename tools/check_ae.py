"""GPU check of the autoencoder decoder, layer by layer against oracle/ae_oracle.py (test infrastructure).

    python tools/check_ae.py [small|full|bench] ...
"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from duodiff_b200.autoencoder import FrozenAutoencoderKL  # noqa: E402
from oracle import ae_oracle as A  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def dd_of(spec):
    return dict(double_z=True, z_channels=spec.z_channels, resolution=spec.resolution, in_channels=3,
                out_ch=spec.out_ch, ch=spec.ch, ch_mult=spec.ch_mult, num_res_blocks=spec.num_res_blocks,
                attn_resolutions=[], dropout=0.0)


def layerwise(spec, B, seed=0):
    sd = A.random_state_dict(spec, seed)
    ae = FrozenAutoencoderKL(dd_of(spec), spec.embed_dim, state_dict=sd, scale_factor=spec.scale_factor, max_batch=B)
    g = torch.Generator().manual_seed(seed + 1)
    z = torch.randn(B, spec.z_channels, spec.z_res, spec.z_res, generator=g) * spec.scale_factor
    tap = {}
    torch.set_num_threads(16)
    ref = A.decode(sd, spec, z, tap)
    zc = z.cuda()
    worst = 0.0
    for i, (name, c, h, w, f32) in enumerate(ae.ops()):
        if c == 0 or name not in tap:
            continue
        out, dump = ae.decode_debug(zc, i)
        t = tap[name]
        d = dump.cpu()[:, :t.shape[1]]
        e = rel(d, t)
        worst = max(worst, e)
        print(f"  op {i:3d} {name:28s} [{c:4d},{h:3d},{w:3d}] rel-L2 {e:.3e}  max|ref| {float(t.abs().max()):.2f}")
    out = ae.decode(zc).cpu()
    print(f"final image rel-L2 {rel(out, ref):.3e} max-abs {float((out - ref).abs().max()):.3e} "
          f"|ref|max {float(ref.abs().max()):.2f}  worst layer {worst:.3e}")
    return rel(out, ref)


def bench(B, chunk, iters=3):
    spec = A.AESpec()
    sd = A.random_state_dict(spec, 0)
    ae = FrozenAutoencoderKL(dd_of(spec), 4, state_dict=sd, max_batch=chunk)
    z = torch.randn(B, 4, 32, 32, device="cuda") * 0.18215
    for _ in range(2):
        ae.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        ae.decode(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    prof = ae.profile_decode(z)
    tot_f = sum(v["flops"] for v in prof.values())
    print(f"decode B={B} chunk={chunk}: {ms:.2f} ms  ({ms / B:.3f} ms/img, {tot_f / ms / 1e9:.1f} TFLOP/s algorithmic, "
          f"{tot_f / B / 1e9:.1f} GFLOP/img)")
    for k, v in prof.items():
        tf = v["flops"] / v["ms"] / 1e9 if v["ms"] > 0 else 0.0
        print(f"    {k:14s} {v['ms']:8.3f} ms  {tf:7.1f} TFLOP/s")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "small"
    if mode == "small":
        layerwise(A.AESpec(ch=64, ch_mult=[1, 2], num_res_blocks=1, resolution=32), B=2)
        layerwise(A.AESpec(ch=64, ch_mult=[1, 2, 2], num_res_blocks=1, resolution=128), B=3, seed=5)
    elif mode == "full":
        layerwise(A.AESpec(), B=2)
    elif mode == "ab":  # A/B of the MT = 2 work units for the N = 128 layers
        from duodiff_b200 import _lib
        for v in (0, 1, 0, 1):
            _lib.check(_lib.load().ddb_set_option(b"conv_mt2", v))
            print("conv_mt2 =", v)
            bench(64, 32)
    elif mode == "bench":
        bench(int(sys.argv[2]) if len(sys.argv) > 2 else 32, int(sys.argv[3]) if len(sys.argv) > 3 else 16)
