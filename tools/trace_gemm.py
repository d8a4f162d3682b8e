"""Per-tile timeline of the CTA-pair GEMM (cluster 0): MMA issuer and epilogue warp 0 clock64 stamps."""
import argparse, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from duodiff_b200 import _lib
ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="fc1"); ap.add_argument("--debug", type=int, default=0)
a = ap.parse_args()
L = _lib.load(); dev = torch.device("cuda:0")
M, D = 257 * 128, 512
shapes = {"qkv": (3 * D, D, 0, 1), "proj": (D, D, 0, 3), "fc1": (4 * D, D, 0, 2), "fc2": (D, 4 * D, 0, 3), "skip": (D, D, D, 0)}
N, K0, K1, epi = shapes[a.shape]; K = K0 + K1
a0 = torch.randn(M, K0, device=dev).bfloat16(); a1 = torch.randn(M, K1, device=dev).bfloat16() if K1 else None
w = (torch.randn(N, K, device=dev) * 0.05).bfloat16(); bias = torch.randn(N, device=dev); colsum = w.float().sum(1).contiguous()
NP = K // 64
stats = torch.stack([torch.zeros(M, NP, device=dev), torch.full((M, NP), 64.0, device=dev)], 2).contiguous()
res = torch.randn(M, N, device=dev).bfloat16() if epi == 3 else None
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); sout = torch.empty(M, N // 64, 2, device=dev)
tr = torch.zeros(256 + 74 * 4, dtype=torch.int64, device=dev)
def run():
    so = sout if epi in (0, 3) else None
    _lib.check(L.ddb_op_gemm(_lib.ptr(a0), _lib.ptr(a1), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(colsum), _lib.ptr(stats), NP, K,
                             _lib.ptr(res), _lib.ptr(out), _lib.ptr(so), M, N, K0, K1, epi, 2, _lib.current_stream_ptr()))
_lib.check(L.ddb_set_option(b"gemm_debug", a.debug))
for _ in range(3): run()
torch.cuda.synchronize()
_lib.check(L.ddb_debug_set_ptr(b"gemm_trace", tr.data_ptr()))
run()
torch.cuda.synchronize()
_lib.check(L.ddb_debug_set_ptr(b"gemm_trace", None))
tc = tr.cpu(); t = tc[:256].view(16, 16); t0 = int(t[0, 0])
print(f"{a.shape} debug={a.debug}: cycles relative to the MMA thread's first stamp")
print("tile | MMA: start  tempty_ok  first_full  issued(all) | EPI(wg0,warp0): start aux_ok tfull_ok | c0: ld_done math_done bar_done store_issued | c1: ld_done math_done bar_done store_issued")
for i in range(16):
    v = [int(x) - t0 if int(x) else -1 for x in t[i]]
    if v[0] < 0 and i > 0: break
    print(f"{i:4d} | {v[0]:8d} {v[1]:8d} {v[2]:8d} {v[3]:8d} | {v[4]:8d} {v[5]:8d} {v[6]:8d} | {v[7]:8d} {v[8]:8d} {v[9]:8d} {v[10]:8d} | {v[11]:8d} {v[12]:8d} {v[13]:8d} {v[14]:8d}")

