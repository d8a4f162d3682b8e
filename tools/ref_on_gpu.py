"""Context measurement (NOT the optimisation target): the UNMODIFIED reference modules (baseline/_ref, git-ignored copy
of the reference's models/ utils/ sampler.py ...) on one B200 as shipped -- fp32 with TF32 off (the reference's real
default: sampler.py:25-38,129-133, no autocast) -- and under bf16 autocast, at the BASELINE batch of each config.
Per config: seconds per forward of the shallow and the full backbone, the reference's own DDPM update, and the DuoDiff
images/s they extrapolate to (300 shallow + 700 full steps + 1000 updates).

    python tools/ref_on_gpu.py [celeba imagenet64 ...]   -> gpurun_out/ref_on_gpu.json"""
import json
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"
sys.path.insert(0, str(REF))
mpl, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
mpl.pyplot = pp
sys.modules.setdefault("matplotlib", mpl)
sys.modules.setdefault("matplotlib.pyplot", pp)

import torch  # noqa: E402
import yaml  # noqa: E402

import sampler as ref_sampler  # noqa: E402  (baseline/_ref/sampler.py)
from models.uvit import UViT  # noqa: E402

PAIRS = {"celeba": ("uvit_celeba_3", "uvit_celeba", 128), "cifar10": ("uvit_cifar10_3", "uvit_cifar10", 128),
         "imagenet64": ("uvit_imagenet64_3", "uvit_imagenet64", 256),
         "imagenet256": ("uvit_imagenet256_3", "uvit_imagenet256", 256)}
dev = "cuda:0"


def cfg(n):
    p = yaml.safe_load(open(REF / "configs" / f"{n}.yaml"))["model_params"]
    p.pop("classifier_type", None)  # stray key of uvit_imagenet64.yaml
    return p


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n / 1e3


out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "reference_device": str(ref_sampler.device),
       "configs": {}}
for name in (sys.argv[1:] or ["celeba"]):
    sh, fu, B = PAIRS[name]
    r = {"batch": B}
    for mode in ("fp32_tf32_off", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        secs = {}
        for key, cn in (("shallow", sh), ("full", fu)):
            p = cfg(cn)
            torch.manual_seed(0)
            m = UViT(**p).eval().to(dev)
            x = torch.randn(B, p["in_chans"], p["img_size"], p["img_size"], device=dev)
            t = 500 * torch.ones(B, device=dev)
            y = torch.randint(0, p["num_classes"], (B,), device=dev) if p["num_classes"] > 0 else None

            def run():
                with torch.no_grad():
                    if mode == "bf16_autocast":
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            return m(x, t, y)
                    return m(x, t, y)
            secs[key] = timeit(run)
            del m
            torch.cuda.empty_cache()
        eps = torch.randn_like(x)
        secs["update"] = timeit(lambda: ref_sampler.predict_noise_postprocessing(eps, x, 500), n=20)
        total = 300 * secs["shallow"] + 700 * secs["full"] + 1000 * secs["update"]
        r[mode] = dict(seconds=secs, duodiff_images_per_s=B / total)
        print(name, mode, r[mode], flush=True)
    out["configs"][name] = r
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "ref_on_gpu.json", "w"), indent=1)
print(json.dumps(out))
