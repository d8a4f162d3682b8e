"""Mean exit layer of the synthetic DeeDiff probes over the WHOLE 1000-step trajectory as a function of the bias slope
(bench.py --ee --slope): random-init probes never cross the threshold, so bias_i = -slope * i positions the exits.
    python tools/ee_calibrate.py [--config celeba] [--batch 128] [--slopes 0.5,1,2,4] [--wscale 4]"""
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
ap.add_argument("--threshold", type=float, default=0.08)
ap.add_argument("--slopes", default="0.5,1,2,4")
ap.add_argument("--wscale", type=float, default=4.0)
a = ap.parse_args()
dev = torch.device("cuda:0")
p = CONFIGS[a.config]
B, depth = a.batch, p["depth"]
for slope in [float(v) for v in a.slopes.split(",")]:
    torch.manual_seed(1234)
    net = ddb.EarlyExitUViT(ddb.UViT(**p, max_batch=B), "mlp_probe_per_layer")
    with torch.no_grad():
        for i in range(depth):
            net.matrix[f"{i}"].classifier[0].weight.mul_(a.wscale)
            net.matrix[f"{i}"].classifier[0].bias.fill_(-slope * i)
    net = net.eval().to(dev)
    smp = Sampler(net.engine(B), None, float("inf"), B, ee_threshold=a.threshold, ee_mode=1)
    x = torch.randn(B, p["in_chans"], p["img_size"], p["img_size"], device=dev,
                    generator=torch.Generator(device=dev).manual_seed(7))
    y = torch.randint(0, p["num_classes"], (B,), device=dev) if p["num_classes"] > 0 else None
    exit_log = torch.zeros(1000, B, device=dev, dtype=torch.int32)
    score_log = torch.zeros(1000, depth, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    smp.run(x, y=y, seed=1, exit_log=exit_log, score_log=score_log)
    e1.record()
    torch.cuda.synchronize()
    idx = exit_log.float()
    per_t = idx.mean(1)
    print(f"slope {slope:4.2f} wscale {a.wscale}: mean exit {idx.mean().item():5.2f}/{depth}  "
          f"(t=999: {per_t[999].item():.1f}, 750: {per_t[750].item():.1f}, 500: {per_t[500].item():.1f}, "
          f"250: {per_t[250].item():.1f}, 0: {per_t[0].item():.1f})  |x| max {x.abs().max().item():.1f}  "
          f"{e0.elapsed_time(e1):.0f} ms per pass (incl. capture)  hist {torch.bincount(exit_log.flatten().long().cpu(), minlength=depth + 1).tolist()}",
          flush=True)
    del smp, net
