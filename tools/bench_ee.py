"""Early-exit sampling throughput (BASELINE config 3: DeeDiff CelebA, mlp_probe_per_layer, threshold 0.08):
simulate mode (reference semantics: every layer/probe/head evaluated) vs compaction mode, on SYNTHETIC probes --
random-init probes sit at 0.47..0.55 and never exit, so the probe biases are set to
logit-space offsets that decrease with depth (mean exit layer ~0.6 depth, per-sample spread from the random weights).
    python tools/bench_ee.py [--batch 128] [--steps 200] [--threshold 0.08]"""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import duodiff_b200 as ddb  # noqa: E402
from duodiff_b200.configs import CONFIGS  # noqa: E402
from duodiff_b200.ddpm import Sampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=200, help="DDPM steps timed per mode (t = 999 ... 1000-steps)")
ap.add_argument("--threshold", type=float, default=0.08)
ap.add_argument("--slope", type=float, default=0.35)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, depth = a.batch, CONFIGS["celeba"]["depth"]
torch.manual_seed(1234)
net = ddb.EarlyExitUViT(ddb.UViT(**CONFIGS["celeba"], max_batch=B), "mlp_probe_per_layer")
with torch.no_grad():
    for i in range(depth):
        net.matrix[f"{i}"].classifier[0].weight.mul_(4.0)
        net.matrix[f"{i}"].classifier[0].bias.fill_(-a.slope * i)
net = net.eval().to(dev)
eng = net.engine(B)
gen = torch.Generator(device=dev).manual_seed(7)
x0 = torch.randn(B, 3, 64, 64, device=dev, generator=gen)
res = {}
for thr, mode, label in ((a.threshold, 0, "simulate"), (a.threshold, 1, "compact"), (0.0, 1, "compact, threshold 0 (never exits)")):
    smp = Sampler(eng, None, float("inf"), B, ee_threshold=thr, ee_mode=mode)
    exit_log = torch.zeros(1000, B, device=dev, dtype=torch.int32)
    score_log = torch.zeros(1000, depth, device=dev)
    x = x0.clone()
    smp.run(x, seed=1, t_first=999, t_last=990, exit_log=exit_log, score_log=score_log)  # warm-up + graph capture
    torch.cuda.synchronize()
    x = x0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    smp.run(x, seed=1, t_first=999, t_last=1000 - a.steps, exit_log=exit_log, score_log=score_log)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    idx = exit_log[1000 - a.steps:].float()
    res[label] = ms
    print(f"{label:38s}: {ms:7.3f} ms/step -> {B / (ms * 1e-3 * 1000):6.1f} images/s at 1000 steps; "
          f"mean exit layer {idx.mean().item():5.2f} of {depth} (min {int(idx.min())}, max {int(idx.max())})", flush=True)
print(f"compaction speed-up over simulate: {res['simulate'] / res['compact']:.2f}x")
