"""Attention micro-benchmark on the U-ViT shapes (CUDA events, inputs rotated over > L2 worth of buffers).
    python tools/bench_attn.py [--B 128] [--L 257] [--H 8] [--variants 2,1] [--n 50]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from duodiff_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--L", type=int, default=257)
ap.add_argument("--H", type=int, default=8)
ap.add_argument("--variants", default="2")
ap.add_argument("--n", type=int, default=50)
ap.add_argument("--check", action="store_true")
ap.add_argument("--debug", default="0")
ap.add_argument("--lib", default=None, help="alternative libduodiff_b200.so (A/B builds)")
ap.add_argument("--opt", action="append", default=[], help="ddb_set_option name=value (repeatable), e.g. attn_token=0")
a = ap.parse_args()
if a.lib:  # an older build (A/B): bind only what this tool calls
    import ctypes as C
    Lb = C.CDLL(a.lib)
    Lb.ddb_op_attention.restype = C.c_int
    Lb.ddb_op_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    Lb.ddb_set_option.restype = C.c_int
    Lb.ddb_set_option.argtypes = [C.c_char_p, C.c_int32]
    Lb.ddb_last_error.restype = C.c_char_p
    _lib._lib = Lb  # _lib.check() reads the error text from this library
else:
    Lb = _lib.load()
for o in a.opt:
    k, v = o.split("=")
    _lib.check(Lb.ddb_set_option(k.encode(), int(v)))
dev = torch.device("cuda:0")
B, L, H = a.B, a.L, a.H
D = H * 64
NB = 3
qkv = [(torch.randn(B * L, 3 * D, device=dev) * 1.5).bfloat16() for _ in range(NB)]
out = [torch.zeros(B * L, D, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
flops = 4.0 * B * H * L * L * 64
for variant, dbg in [(int(v), int(d)) for v in a.variants.split(",") for d in a.debug.split(",")]:

    def run(i):
        _lib.check(Lb.ddb_op_attention(_lib.ptr(qkv[i % NB]), _lib.ptr(out[i % NB]), B, L, H, variant,
                                       _lib.current_stream_ptr()))
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    if a.check:
        x = qkv[0].float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
        ref = (torch.softmax(x[0] @ x[1].transpose(-1, -2) * 0.125, -1) @ x[2]).permute(0, 2, 1, 3).reshape(B * L, D)
        err = ((out[0].float() - ref).norm() / ref.norm()).item()
        print(f"variant {variant}: rel-L2 vs fp32 math {err:.2e}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a.n
    print(f"attention B={B} L={L} H={H} variant={variant} debug={dbg}: {us:7.1f} us  {flops / us / 1e6:7.1f} TFLOP/s "
          f"({(4 * B * L * D * 2) / us / 1e3:6.0f} GB/s algorithmic qkv+out)", flush=True)
