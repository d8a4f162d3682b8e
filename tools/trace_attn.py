"""Phase timeline of the persistent attention kernel (CTA 0): clock64 stamps per item and query tile."""
import sys
VAR = int(sys.argv[1]) if len(sys.argv) > 1 else 2
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from duodiff_b200 import _lib
import os
Lb = _lib.load(os.environ.get("DDB_LIB"))  # DDB_LIB=build/lib_x.so: trace an A/B build
dev = torch.device("cuda:0")
B, L, H = 128, 257, 8
D = H * 64
qkv = [(torch.randn(B * L, 3 * D, device=dev) * 1.5).bfloat16() for _ in range(3)]
out = torch.zeros(B * L, D, device=dev, dtype=torch.bfloat16)
tr = torch.zeros(8, 2, 16, dtype=torch.int64, device=dev)
for i in range(3):
    _lib.check(Lb.ddb_op_attention(_lib.ptr(qkv[i]), _lib.ptr(out), B, L, H, VAR, _lib.current_stream_ptr()))
torch.cuda.synchronize()
_lib.check(Lb.ddb_debug_set_ptr(b"attn_trace", tr.data_ptr()))
_lib.check(Lb.ddb_op_attention(_lib.ptr(qkv[0]), _lib.ptr(out), B, L, H, VAR, _lib.current_stream_ptr()))
torch.cuda.synchronize()
_lib.check(Lb.ddb_debug_set_ptr(b"attn_trace", None))
t = tr.cpu()
t0 = int(t[0, 0, 0])
names = ["start", "s_ok", "rel", "pass1", "p_done", "se_next", "o_ok", "end"]
print("tile  item " + " ".join(f"{n:>8s}" for n in names) + "   | deltas: swait rel pass1 pass2 se_nxt owait epi")
for tile in range(2):
    for it in range(7):
        v = [int(x) - t0 for x in t[it, tile, :8]]
        d = [v[i + 1] - v[i] for i in range(7)]
        m = [int(x) - t0 for x in t[it, tile, 8:14]]
        print(f"{tile:4d} {it:5d} tokwait {int(t[it, tile, 14]) - t0 - v[3]:5d} " + " ".join(f"{x:8d}" for x in v) + "   | " + " ".join(f"{x:6d}" for x in d)
              + f"   || mma: s_issued {m[0]} phalf_seen {m[1]} pv1_issued {m[2]} pfull_seen {m[3]} pv2_issued {m[4]} o_done {m[5]}"
              + f"  (WG p_done->mma seen {m[3] - v[4]}, issue {m[4] - m[3]}, exec {m[5] - m[4]}, mma o_done->WG o_ok {v[6] - m[5]})"
              + f"  token wait {int(t[it, tile, 14]) - t0 - v[3]}")
