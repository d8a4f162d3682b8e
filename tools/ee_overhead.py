"""Price list of the early-exit bookkeeping in a NEVER-EXIT step (threshold 0, compaction on): whole 1000-step passes
through the captured step graph with parts of the bookkeeping left out (ddb_set_option "ee_debug", bench-only), next to
the plain backbone.   python tools/ee_overhead.py [--config celeba] [--batch 128]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import duodiff_b200 as ddb  # noqa: E402
from bench import CONFIGS, PAIRS, synthetic_probes_  # noqa: E402
from duodiff_b200 import _lib  # noqa: E402
from duodiff_b200.ddpm import Sampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="celeba")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=400)
a = ap.parse_args()
dev = torch.device("cuda:0")
_, full, default_b = PAIRS[a.config]
B = a.batch or default_b
pf = CONFIGS[full]
torch.manual_seed(1234)
net = ddb.EarlyExitUViT(ddb.UViT(**pf, max_batch=B), "mlp_probe_per_layer")
synthetic_probes_(net, pf["depth"], 1.0, 0.5)
net = net.eval().to(dev)
eng = net.engine(B)
lib = _lib.load()
y = torch.randint(0, pf["num_classes"], (B,), device=dev) if pf["num_classes"] > 0 else None
x0 = torch.randn(B, pf["in_chans"], pf["img_size"], pf["img_size"], device=dev)
never = Sampler(eng, None, float("inf"), B, ee_threshold=0.0, ee_mode=1)
plain = Sampler(eng, None, float("inf"), B)


def timed(smp) -> float:
    t_last = 1000 - a.steps
    smp.run(x0.clone(), y=y, seed=0, t_first=999, t_last=950, use_graph=True)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(2):
        x = x0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        smp.run(x, y=y, seed=1, t_first=999, t_last=t_last, use_graph=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    return best


base = timed(plain)
print(f"plain backbone                                   {base:.4f} ms/step")
for dbg, what in [(0, "never-exit, everything on"), (1, "- row move launches"), (3, "- row move, decision launches"),
                  (7, "- move, decision, probe partials in fc2"), (15, "- all of the above, first-layer probe pass"),
                  (0, "never-exit, everything on (again)")]:
    _lib.check(lib.ddb_set_option(b"ee_debug", dbg))
    ms = timed(never)
    print(f"ee_debug={dbg:2d} {what:45s} {ms:.4f} ms/step  (+{(ms - base) * 1e3:6.1f} us, +{(ms / base - 1) * 100:4.1f} %)")
_lib.check(lib.ddb_set_option(b"ee_debug", 0))
base = timed(plain)
print(f"plain backbone (again)                           {base:.4f} ms/step")
