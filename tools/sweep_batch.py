"""Batch sweep of full DuoDiff sampling (resident inputs, CUDA-graph replay) for one config pair:
    python tools/sweep_batch.py imagenet256 64,128,256,512,1024
One warm-up segment (graph capture) and one timed 1000-step pass per batch size; prints images/s and the fraction of
the measured sustained bf16 peak (MEASURED_PEAKS.json) the whole path reaches."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (FLOP model, config pairs)
import duodiff_b200 as ddb  # noqa: E402
from duodiff_b200.configs import CONFIGS  # noqa: E402
from duodiff_b200.ddpm import Sampler  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "imagenet256"
batches = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "64,128,256").split(",")]
shallow, full, _ = bench.PAIRS[name]
ps, pf = CONFIGS[shallow], CONFIGS[full]
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops_sustained"] if (ROOT / "MEASURED_PEAKS.json").exists() else 1364.9
fpi = 300 * bench.forward_flops(ps)["total"] + 700 * bench.forward_flops(pf)["total"]
dev = torch.device("cuda:0")
print(f"# {name}: {shallow} + {full}, t_switch = 300, {fpi / 1e12:.2f} TFLOP per image; sustained bf16 peak {peak} TFLOP/s")
for B in batches:
    torch.manual_seed(1234)
    early = ddb.UViT(**ps, max_batch=B).eval().to(dev)
    late = ddb.UViT(**pf, max_batch=B).eval().to(dev)
    y = torch.randint(0, pf["num_classes"], (B,), device=dev) if pf["num_classes"] > 0 else None
    smp = Sampler(early.engine(B), late.engine(B), 300, B)
    C, H = pf["in_chans"], pf["img_size"]
    x = torch.randn(B, C, H, H, device=dev)
    smp.run(x.clone(), y=y, seed=0, t_first=999, t_last=995)   # shallow graph
    smp.run(x.clone(), y=y, seed=0, t_first=699, t_last=695)   # full graph
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    xx = x.clone()
    e0.record()
    smp.run(xx, y=y, seed=1)
    out = smp.finalize(xx)
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    s = e0.elapsed_time(e1) / 1e3
    ips = B / s
    print(f"B = {B:5d}: {s:8.2f} s per batch  {ips:7.2f} images/s  {ips * fpi / 1e12:7.1f} TFLOP/s  "
          f"{ips * fpi / 1e12 / peak:5.3f} of sustained peak   mem {torch.cuda.max_memory_allocated() / 2**30:5.1f} GiB (torch side)",
          flush=True)
    del smp, early, late
    torch.cuda.empty_cache()
