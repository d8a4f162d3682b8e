#!/usr/bin/env python
"""bench.py — DuoDiff sampling throughput (images/sec, 1000 DDPM steps, t_switch=300) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config celeba] [--batch 128]

A "step" is one pass of the hot path over one batch: a full 1000-step DuoDiff sampling of `batch` images per GPU
(300 shallow-U-ViT steps, 700 full-U-ViT steps, 1000 DDPM updates).  One JSON line on stdout (rank 0):
  value  = whole-job images/sec with x_T already resident in HBM, CUDA-event timed, max over ranks
  e2e    = the same through the public API duodiff_b200.sampler.get_samples(): x_T drawn on the host, pinned
           H2D copy, 1000 steps, (x+1)/2 NHWC, D2H to numpy — all inside the timed region
  roofline / kernels = per-kernel CUDA-event timings of one shallow + one full forward (ddb_profile_forward)
  cpu_baseline = the oracle (or baseline/_ref when present) on the host cores, bounded sample, rank 0, N=1 only
`--impl reference` times the reference's own CPU implementation of the path instead.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from duodiff_b200.configs import CONFIGS  # noqa: E402

PAIRS = {  # BASELINE.json configs -> (shallow, full, default batch per GPU)
    "cifar10": ("cifar10_3", "cifar10", 128),
    "celeba": ("celeba_3", "celeba", 128),
    "imagenet64": ("imagenet64_3", "imagenet64", 256),
    "imagenet256": ("imagenet256_3", "imagenet256", 256),
}
T_SWITCH = 300
METRIC = "images/sec (DuoDiff sampling, 1000 steps)"


def forward_flops(p: dict) -> dict:
    """Algorithmic FLOPs per image per forward, split per kernel category."""
    D, d = p["embed_dim"], p["depth"]
    N = (p["img_size"] // p["patch_size"]) ** 2
    L = N + (2 if p["num_classes"] > 0 else 1)
    pd = p["patch_size"] ** 2 * p["in_chans"]
    hid = int(D * p["mlp_ratio"])
    per = dict(gemm_qkv=d * 2 * L * D * 3 * D, gemm_proj=d * 2 * L * D * D, gemm_fc1=d * 2 * L * D * hid,
               gemm_fc2=d * 2 * L * hid * D, gemm_skip=(d // 2) * 2 * L * 2 * D * D, attention=d * 4 * L * L * D,
               gemm_decode=2 * L * D * pd, embed=2 * N * pd * D,
               conv=18 * p["in_chans"] ** 2 * p["img_size"] ** 2)
    per["total"] = sum(per.values())
    return per


def peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), tflops_burst=d.get("bf16_tflops"),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons, power = [], 0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                power = max(power, float(r[6]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [c for c in sm if c > 0]
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), power_w_max=power or None, samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU baseline
def _ref_modules():
    """The unmodified reference under baseline/_ref (staged in the build container; travels with gpurun)."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "models" / "uvit.py").exists():
        return None
    sys.path.insert(0, str(ref))
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            from models.uvit import UViT as RefUViT  # type: ignore
        return RefUViT
    except Exception:  # noqa: BLE001
        return None
    finally:
        sys.path.remove(str(ref))


def cpu_sample(pair: str, batch: int, reps: int, use_reference: bool) -> dict:
    """Bounded CPU sample of the same workload: `reps` x (one shallow forward + one full forward + 2 DDPM updates)
    at `batch` images on the host cores, extrapolated to 300 shallow + 700 full steps."""
    import contextlib
    import io

    from oracle import uvit_oracle as O
    shallow, full, _ = PAIRS[pair]
    torch.manual_seed(1234)
    RefUViT = _ref_modules() if use_reference else None
    fwd, kind = {}, "port"
    for name in (shallow, full):
        p = CONFIGS[name]
        if RefUViT is not None:
            with contextlib.redirect_stdout(io.StringIO()):
                m = RefUViT(**p).eval()
            fwd[name] = (lambda mm: (lambda x, t, y: mm(x, t, y)))(m)
            kind = "reference"
        else:
            import duodiff_b200 as ddb
            sd = ddb.UViT(**p).state_dict()
            spec = O.UViTSpec.from_params(p)
            fwd[name] = (lambda s, sp: (lambda x, t, y: O.uvit_forward(s, sp, x, t, y)))(sd, spec)
    p = CONFIGS[full]
    sch = O.ddpm_schedule()
    x = torch.randn(batch, p["in_chans"], p["img_size"], p["img_size"])
    y = torch.randint(0, p["num_classes"], (batch,)) if p["num_classes"] > 0 else None
    times = {shallow: [], full: [], "update": []}
    with torch.no_grad():
        for rep in range(reps + 1):  # first rep is warm-up
            for name, t in ((shallow, 800), (full, 400)):
                tt = t * torch.ones(batch)
                t0 = time.perf_counter()
                eps = fwd[name](x, tt, y)
                t1 = time.perf_counter()
                x = O.predict_noise_step(sch, eps, x, t, torch.randn_like(x))
                t2 = time.perf_counter()
                if rep:
                    times[name].append(t1 - t0)
                    times["update"].append(t2 - t1)
    ts, tf, tu = (statistics.mean(times[k]) for k in (shallow, full, "update"))
    total = T_SWITCH * ts + (1000 - T_SWITCH) * tf + 1000 * tu
    return dict(value=batch / total, unit="images/sec", cores=torch.get_num_threads(), kind=kind,
                sample=f"{reps}x(1 {shallow} fwd + 1 {full} fwd + 2 DDPM updates) at batch {batch}, fp32, "
                       f"extrapolated to {T_SWITCH}+{1000 - T_SWITCH} steps "
                       f"(t_shallow={ts:.3f}s t_full={tf:.3f}s t_update={tu:.4f}s)",
                host_cpus=os.cpu_count())


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    batch = 8
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host thread
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    vals, info = [], None
    t_begin = time.perf_counter()
    for i in range(warm + steps):
        info = cpu_sample(args.config, batch, 1, use_reference=True)
        if i >= warm:
            vals.append(info["value"])
    ms = (time.perf_counter() - t_begin) * 1e3 / (warm + steps)
    v = statistics.mean(vals)
    shallow, full, default_b = PAIRS[args.config]
    B = args.batch or default_b
    info["value"] = v
    # same metric / unit / config as the GPU arm; what was actually timed is described in cpu_baseline.sample
    line = dict(impl="reference", metric=METRIC, value=v, unit="images/sec", n_gpus=args.gpus, steps=steps,
                warmup=warm, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic (random-init weights, N(0,1) x_T)",
                config=dict(workload=f"DuoDiff {args.config} ({shallow}+{full}) 1000 DDPM steps, t_switch={T_SWITCH}",
                            batch_per_gpu=B, global_batch=B * args.gpus, t_switch=T_SWITCH,
                            parallelism=f"dp{args.gpus}", l2="n/a (host cores)"),
                cpu_baseline=info, e2e=dict(value=v, unit="images/sec", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    import torch.distributed as dist

    import duodiff_b200 as ddb
    from duodiff_b200 import _lib
    from duodiff_b200 import sampler as S
    from duodiff_b200.ddpm import Sampler

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shallow, full, default_b = PAIRS[args.config]
    B = args.batch or default_b
    ps, pf = CONFIGS[shallow], CONFIGS[full]
    torch.manual_seed(1234)  # identical random-init weights on every rank
    early = ddb.UViT(**ps, max_batch=B).eval().to(dev)
    late = ddb.UViT(**pf, max_batch=B).eval().to(dev)
    C, H = pf["in_chans"], pf["img_size"]
    y = None
    if pf["num_classes"] > 0:
        y = torch.randint(0, min(ps["num_classes"], pf["num_classes"]), (B,), device=dev)
    lib = _lib.load()
    ae = None
    if args.decode:  # latent configs: KL-autoencoder decode after the loop (sampler.py:141-143), random-init weights
        from duodiff_b200 import autoencoder as AE
        assert C == 4 and H == 32, "--decode needs a latent config (imagenet256)"
        with contextlib.redirect_stdout(sys.stderr):  # the shim prints like the reference; stdout carries the JSON line
            ae = AE.FrozenAutoencoderKL(AE.DEFAULT_DDCONFIG, 4, state_dict=AE.random_init_state_dict(seed=4321),
                                        max_batch=min(B, 32))
    smp = Sampler(early.engine(B), late.engine(B), T_SWITCH, B)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    n_total = args.warmup + args.steps
    x_all = [torch.randn(B, C, H, H, device=dev, generator=gen) for _ in range(min(n_total, 4))]
    oshape = (B, 256, 256, 3) if ae is not None else (B, H, H, C)
    gathered = [torch.empty(*oshape, device=dev) for _ in range(world)] if world > 1 else None

    def one_pass(i: int):
        x = x_all[i % len(x_all)].clone()
        smp.run(x, y=y, seed=rank * 7919 + i, use_graph=True)
        out = smp.finalize(ae.decode(x) if ae is not None else x)
        if world > 1:
            dist.all_gather(gathered, out)  # the path's only collective: finished samples
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input throughput (value)
    for i in range(args.warmup):
        one_pass(i)
    barrier()
    l0 = lib.ddb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for i in range(args.steps):
            out = one_pass(args.warmup + i)
        e1.record()
        barrier()
    launches = lib.ddb_launch_count() - l0
    ms = e0.elapsed_time(e1)
    assert torch.isfinite(out).all(), "non-finite samples"

    # ---- end-to-end through the public API (host x_T -> H2D -> 1000 steps -> NHWC -> D2H numpy)
    def e2e_pass(i: int):
        with contextlib.redirect_stdout(sys.stderr):
            return _e2e_pass(i)

    def _e2e_pass(i: int):
        return S.get_samples(early, B, S.predict_noise_postprocessing, seed=rank * 104729 + i, num_channels=C,
                             sample_height=H, sample_width=H, use_ddim=False, ddim_steps=50, ddim_eta=0.0,
                             timesteps_save=[], y=y, autoencoder=ae, late_model=late, t_switch=T_SWITCH,
                             device=dev)[0]

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for i in range(min(args.warmup, 1)):
        e2e_pass(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        host = e2e_pass(100 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert host.shape == oshape

    # ---- reduce over ranks (max time)
    tt = torch.tensor([ms, e2e_s * 1e3, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = tt[0].item(), tt[1].item()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (ms_max / 1e3)
    e2e_val = world * B * e2e_steps / (e2e_ms_max / 1e3)
    pk = peaks()
    fl_s, fl_f = forward_flops(ps), forward_flops(pf)
    flops_per_image = T_SWITCH * fl_s["total"] + (1000 - T_SWITCH) * fl_f["total"]

    # ---- per-kernel roofline (CUDA events around every launch of one forward, eager)
    kernels = {}
    xs = x_all[0]
    tvec = torch.full((B,), 500.0, device=dev)
    for name, net, fl in ((shallow, early, fl_s), (full, late, fl_f)):
        eng = net.engine(B)
        for _ in range(2):
            prof = eng.profile_forward(xs, tvec, y)
        reps = 5
        acc = {k: 0.0 for k in prof}
        for _ in range(reps):
            prof = eng.profile_forward(xs, tvec, y)
            for k, v in prof.items():
                acc[k] += v["ms"] / reps
        tot = sum(acc.values())
        kernels[name] = {
            k: dict(ms=round(acc[k], 4), launches=prof[k]["launches"], share=round(acc[k] / tot, 4),
                    tflops=round(fl[k] * B / (acc[k] * 1e-3) / 1e12, 1) if k in fl and acc[k] > 0 else None)
            for k in acc if prof[k]["launches"]}
        kernels[name]["_forward_ms_sum"] = round(tot, 4)
    kf = kernels[full]
    # the north star's "fraction of bf16 tensor-core peak on the U-ViT GEMMs": GEMM FLOPs / summed GEMM-kernel time
    gk = [k for k in kf if k.startswith("gemm_") and k != "gemm_decode"]
    gemm_tf = sum(fl_f[k] for k in gk) * B / (sum(kf[k]["ms"] for k in gk) * 1e-3) / 1e12
    gemm_only = dict(tflops=round(gemm_tf, 1), frac_of_sustained_peak=round(gemm_tf / pk["tflops"], 4),
                     frac_of_burst_peak=round(gemm_tf / pk["tflops_burst"], 4) if pk.get("tflops_burst") else None,
                     kernels=gk)
    dom = max((k for k in kf if not k.startswith("_")), key=lambda k: kf[k]["ms"])
    dom_ms = kf[dom]["ms"] / kf[dom]["launches"]
    dom_flops = fl_f[dom] * B / kf[dom]["launches"]
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get(dom)
    roofline = dict(bound="tensor", kernel=f"{dom} (gemm2_tcgen05_kernel, CTA pair)" if dom.startswith("gemm") else dom,
                    achieved=round(achieved, 1), peak=pk["tflops"], unit="TFLOP/s", frac=round(achieved / pk["tflops"], 4),
                    traffic=traffic, peak_source=pk["source"] + ", sustained bf16",
                    flops_per_launch=dom_flops, avg_launch_ms=round(dom_ms, 4),
                    gemm_only=gemm_only,
                    whole_path=dict(tflops=round(value * flops_per_image / 1e12, 1),
                                    frac=round(value * flops_per_image / 1e12 / (pk["tflops"] * world), 4),
                                    flops_per_image=flops_per_image))

    cpu = cpu_sample(args.config, 8, 1, use_reference=True) if world == 1 and not args.no_cpu else None
    nbytes = B * C * H * H * 4
    out_bytes = B * 256 * 256 * 3 * 4 if ae is not None else nbytes
    ae_info = None
    if ae is not None:  # per-category device time of one decode of the batch (CUDA events around every launch)
        prof = ae.profile_decode(x_all[0])
        ae_info = dict(ms_per_batch=round(sum(v["ms"] for v in prof.values()), 3),
                       gflop_per_image=round(sum(v["flops"] for v in prof.values()) / B / 1e9, 1),
                       categories={k: round(v["ms"], 3) for k, v in prof.items()})
    line = dict(metric=METRIC, value=round(value, 3), unit="images/sec", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(ms_max / args.steps, 2), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic (random-init weights, N(0,1) x_T, Philox z_t)",
                config=dict(workload=f"DuoDiff {args.config} ({shallow}+{full}) 1000 DDPM steps, t_switch={T_SWITCH}"
                            + (" + KL-autoencoder decode to 3x256x256" if ae is not None else ""),
                            batch_per_gpu=B, global_batch=B * world, t_switch=T_SWITCH, parallelism=f"dp{world}",
                            l2="no flush: one DDPM step touches ~1 GB of activations+weights (> 126 MB L2)"),
                clocks=clocks.summary(),
                e2e=dict(value=round(e2e_val, 3), unit="images/sec", h2d_bytes_per_step=nbytes,
                         d2h_bytes_per_step=out_bytes, steps=e2e_steps, api="duodiff_b200.sampler.get_samples"),
                gpu_launches=int(tt[2].item()), roofline=roofline, kernels=kernels, cpu_baseline=cpu)
    if ae_info is not None:
        line["autoencoder"] = ae_info
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=list(PAIRS), default="celeba")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the config's BASELINE batch)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--decode", action="store_true",
                    help="latent configs: decode the samples with the KL autoencoder inside the timed regions")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun, one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__,
               *sys.argv[1:]]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
