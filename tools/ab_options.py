"""A/B of runtime options on the full DuoDiff CelebA sampling loop (B = 128, resident inputs, CUDA-graph replay):
    python tools/ab_options.py gemm_ts=1 pdl=0 ...
Each option set gets a fresh Sampler (graphs are captured with the options in force); prints ms per 1000 steps and the
median SM clock / power sampled with nvidia-smi during the timed region."""
import subprocess, sys, threading, time
import torch
sys.path.insert(0, "/root/repo")
import duodiff_b200 as ddb
from duodiff_b200 import _lib
from duodiff_b200.configs import CONFIGS
from duodiff_b200.ddpm import Sampler

import os
dev = torch.device("cuda:0")
CFG = os.environ.get("AB_CONFIG", "celeba")          # AB_CONFIG=imagenet64 AB_BATCH=256 AB_REPS=1 python tools/ab_options.py ...
B = int(os.environ.get("AB_BATCH", "128"))
REPS = int(os.environ.get("AB_REPS", "3"))
torch.manual_seed(1234)
early = ddb.UViT(**CONFIGS[CFG + "_3"], max_batch=B).eval().to(dev)
late = ddb.UViT(**CONFIGS[CFG], max_batch=B).eval().to(dev)
YCLS = CONFIGS[CFG]["num_classes"]
Y = torch.randint(0, YCLS, (B,), device=dev) if YCLS > 0 else None
XSHAPE = (B, CONFIGS[CFG]["in_chans"], CONFIGS[CFG]["img_size"], CONFIGS[CFG]["img_size"])
lib = _lib.load()


class Smi(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False
    def run(self):
        while not self.stop:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True).stdout.strip().split(",")
            try: self.rows.append((float(o[0]), float(o[1])))
            except Exception: pass
            time.sleep(0.05)


def run(opts):
    for k, v in opts.items():
        _lib.check(lib.ddb_set_option(k.encode(), int(v)))
    smp = Sampler(early.engine(B), late.engine(B), 300, B)
    x = torch.randn(*XSHAPE, device=dev)
    smp.run(x.clone(), y=Y, seed=0, t_first=999, t_last=990)
    smp.run(x.clone(), y=Y, seed=0, t_first=699, t_last=690)
    torch.cuda.synchronize()
    smi = Smi(); smi.start()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for r in range(REPS):
        smp.run(x.clone(), y=Y, seed=r + 1)
    e1.record(); torch.cuda.synchronize()
    smi.stop = True; smi.join()
    ms = e0.elapsed_time(e1) / REPS
    clk = sorted(r[0] for r in smi.rows)[len(smi.rows) // 2] if smi.rows else 0
    pw = sorted(r[1] for r in smi.rows)[len(smi.rows) // 2] if smi.rows else 0
    print(f"{opts}: {ms:.1f} ms / 1000 steps = {B / ms * 1e3:.2f} img/s   sm {clk:.0f} MHz  {pw:.0f} W", flush=True)
    for k in opts:
        _lib.check(lib.ddb_set_option(k.encode(), {"pdl": 1, "gemm_variant": 2, "alt_dir": 1, "conv_mt2": 1, "attn_discard": 1, "mlp_split": 0, "l2_hints": 0, "gemm_ln_cfg": 0, "attn_token": 1}.get(k, 0)))


run({})
for arg in sys.argv[1:]:
    run(dict(kv.split("=") for kv in arg.split(",")))
run({})
