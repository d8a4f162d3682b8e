"""Minimal driver for profiling: build the two CelebA backbones and run forwards at the headline batch.
    python tools/run_forward.py [--config celeba] [--batch 128] [--iters 2] [--which full|shallow|both]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import duodiff_b200 as ddb  # noqa: E402
from duodiff_b200.configs import CONFIGS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="celeba")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--which", default="full")
ap.add_argument("--opt", default="", help="runtime options, e.g. alt_dir=0,pdl=1")
a = ap.parse_args()
if a.opt:
    from duodiff_b200 import _lib
    for kv in a.opt.split(","):
        k, v = kv.split("=")
        _lib.check(_lib.load().ddb_set_option(k.encode(), int(v)))
dev = torch.device("cuda:0")
torch.manual_seed(0)
names = {"full": [a.config], "shallow": [a.config + "_3"], "both": [a.config + "_3", a.config]}[a.which]
for name in names:
    p = CONFIGS[name]
    net = ddb.UViT(**p, max_batch=a.batch).eval().to(dev)
    x = torch.randn(a.batch, p["in_chans"], p["img_size"], p["img_size"], device=dev)
    t = torch.full((a.batch,), 500.0, device=dev)
    y = torch.randint(0, p["num_classes"], (a.batch,), device=dev) if p["num_classes"] > 0 else None
    for _ in range(a.iters):
        out = net(x, t, y)
    torch.cuda.synchronize()
    print(name, "ok", float(out.abs().mean()))
