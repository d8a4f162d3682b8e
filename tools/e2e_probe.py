import sys, time, torch
sys.path.insert(0, "/root/repo")
import duodiff_b200 as ddb
from duodiff_b200 import sampler as S, _io
from duodiff_b200.ddpm import Sampler
from duodiff_b200.configs import CONFIGS
dev = torch.device("cuda:0")
B = 128
torch.manual_seed(1234)
early = ddb.UViT(**CONFIGS["celeba_3"], max_batch=B).eval().to(dev)
late = ddb.UViT(**CONFIGS["celeba"], max_batch=B).eval().to(dev)
def sync(): torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    _io.seed_everything(rep)
    x = torch.randn(B, 3, 64, 64)
    t1 = time.perf_counter()
    x = x.pin_memory().to(dev, non_blocking=True); sync()
    t2 = time.perf_counter()
    e, l = early.engine(B), late.engine(B); sync()
    t3 = time.perf_counter()
    smp = Sampler(e, l, 300, B); sync()
    t4 = time.perf_counter()
    smp.run(x, seed=rep, use_graph=True)
    t5 = time.perf_counter()
    sync()
    t6 = time.perf_counter()
    out = smp.finalize(x).cpu().numpy()
    t7 = time.perf_counter()
    print(f"rep {rep}: randn {t1-t0:.3f} h2d {t2-t1:.3f} engine {t3-t2:.3f} sampler {t4-t3:.3f} run-call {t5-t4:.3f} run-sync {t6-t5:.3f} fin {t7-t6:.3f} total {t7-t0:.3f}")
t0 = time.perf_counter()
out = S.get_samples(early, B, S.predict_noise_postprocessing, seed=5, num_channels=3, sample_height=64, sample_width=64, late_model=late, t_switch=300, device=dev)[0]
print("get_samples", time.perf_counter() - t0)
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
out = S.get_samples(early, B, S.predict_noise_postprocessing, seed=6, num_channels=3, sample_height=64, sample_width=64, late_model=late, t_switch=300, device=dev)[0]
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
