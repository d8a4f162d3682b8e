"""Minimal driver for profiling the sampler step: the CelebA pair at the headline batch, a few eager DuoDiff steps across
the hand-off (every kernel of the step, fused tail included, as an individual launch).
    python tools/run_steps.py [--config celeba] [--batch 128] [--steps 4]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import duodiff_b200 as ddb  # noqa: E402
from duodiff_b200.configs import CONFIGS  # noqa: E402
from duodiff_b200.ddpm import Sampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="celeba")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
ps, pf = CONFIGS[a.config + "_3"], CONFIGS[a.config]
early = ddb.UViT(**ps, max_batch=a.batch).eval().to(dev)
late = ddb.UViT(**pf, max_batch=a.batch).eval().to(dev)
y = torch.randint(0, pf["num_classes"], (a.batch,), device=dev) if pf["num_classes"] > 0 else None
x = torch.randn(a.batch, pf["in_chans"], pf["img_size"], pf["img_size"], device=dev)
smp = Sampler(early.engine(a.batch), late.engine(a.batch), 300, a.batch)
half = a.steps // 2
smp.run(x, y=y, seed=1, t_first=699 + half, t_last=700 - (a.steps - half), use_graph=False)
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
