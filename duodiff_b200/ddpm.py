"""DDPM schedule + the Python owner of a ``ddb_sampler`` handle (the whole t-loop runs inside the C library)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import Engine

import numpy as np

RULES = ("predict_noise", "predict_original", "predict_previous")


def ddim_timesteps(ddim_steps: int) -> list:
    """sampler.py:104 — np.linspace(0, 999, ddim_steps).astype(int)[::-1]."""
    ts = [int(v) for v in np.linspace(0, 999, ddim_steps).astype(int)[::-1]]
    if any(s >= t for t, s in zip(ts[:-1], ts[1:])):
        raise ValueError(f"ddim_steps={ddim_steps}: the strided schedule must be strictly decreasing (sampler.py:106)")
    return ts


def ddim_coefficients(ddim_steps: int, ddim_eta: float):
    """[1000,4] fp32 table {c0, c1, sigma2, d} for the DDIM update (sampler.py:110-120), mode 2 of the step kernel:
        x' = (c0*(x - c1*eps) + d*eps) + sigma2*z,   c0 = sqrt(abar_s/abar_t), c1 = sqrt(1-abar_t),
        d = sqrt(1 - abar_s - sigma2), sigma2 = betas_tilde[t]*eta (the reference adds sigma_t^2 * z), z = 0 if s == 0.
    Rows of timesteps the schedule does not visit stay zero."""
    sch = schedule()
    ab, bt = sch["alphas_bar"], sch["betas_tilde"]
    table = torch.zeros(1000, 4, dtype=torch.float32)
    ts = ddim_timesteps(ddim_steps)
    for t, s in zip(ts[:-1], ts[1:]):
        sigma2 = bt[t] * ddim_eta
        table[t, 0] = torch.sqrt(ab[s] / ab[t])
        table[t, 1] = torch.sqrt(1 - ab[t])
        table[t, 2] = sigma2 if s > 0 else 0.0
        table[t, 3] = torch.sqrt(1 - ab[s] - sigma2)
    return table.contiguous(), 2


def schedule() -> dict:
    """The five schedule vectors, built with the reference's own torch calls (sampler.py:40-44) on the host so the
    coefficient table is bit-identical to a CPU run of the reference."""
    betas = torch.linspace(1e-4, 0.02, 1000)
    alphas = 1 - betas
    alphas_bar = torch.cumprod(alphas, dim=0)
    alphas_bar_previous = torch.cat([torch.tensor([1.0]), alphas_bar[:-1]])
    betas_tilde = betas * (1 - alphas_bar_previous) / (1 - alphas_bar)
    return dict(betas=betas, alphas=alphas, alphas_bar=alphas_bar, alphas_bar_previous=alphas_bar_previous,
                betas_tilde=betas_tilde)


def step_coefficients(rule: str = "predict_noise", variance: str = "beta_tilde"):
    """[1000,4] fp32 table {c0, c1, sigma, 0} and the kernel's evaluation mode for a post-processing rule.

    predict_noise    (sampler.py:47-56):  mode 0  x' = c0*(x - c1*eps) + sigma*z, c0 = sqrt(1/a), c1 = (1-a)/sqrt(1-abar)
    predict_original (sampler.py:59-72):  mode 1  x' = (c1*out + c0*x) + sigma*z
    predict_previous (sampler.py:75-79):  mode 1  with c0 = 0, c1 = 1
    variance: 'beta_tilde' (sampler.py:50, eesampler.py:76) or 'beta' (ddpm_core.py:56-79 default)."""
    s = schedule()
    var = s["betas_tilde"] if variance == "beta_tilde" else s["betas"]
    sigma = torch.sqrt(var)
    if rule == "predict_noise":
        c0 = torch.sqrt(1 / s["alphas"])
        c1 = (1 - s["alphas"]) / torch.sqrt(1 - s["alphas_bar"])
        mode = 0
    elif rule == "predict_original":
        c1 = torch.sqrt(s["alphas_bar_previous"]) * s["betas"] / (1 - s["alphas_bar"])
        c0 = torch.sqrt(s["alphas"]) * (1 - s["alphas_bar_previous"]) / (1 - s["alphas_bar"])
        mode = 1
    elif rule == "predict_previous":
        c0, c1, mode = torch.zeros(1000), torch.ones(1000), 1
    else:
        raise ValueError(f"unknown parametrization {rule!r}; expected one of {RULES}")
    table = torch.stack([c0, c1, sigma, torch.zeros(1000)], dim=1).to(torch.float32).contiguous()
    return table, mode


def switch_step(t_switch) -> int:
    """sampler.py:135-136: the swap happens after the step at t == 1000 - t_switch, i.e. `early` runs t >= 1000 -
    t_switch.  The equality can only trigger for integral 1 <= t_switch <= 1000; otherwise the early model runs
    the whole trajectory (the reference default t_switch=inf)."""
    try:
        ts = float(t_switch)
    except (TypeError, ValueError):
        return -1
    if ts != ts or ts in (float("inf"), float("-inf")) or ts != int(ts):
        return -1
    ts = int(ts)
    return 1000 - ts if 1 <= ts <= 1000 else -1


_SAMPLER_CACHE: "dict[tuple, Sampler]" = {}


def cached_sampler(early: Engine, late: "Engine | None", t_switch, batch: int, rule: str = "predict_noise",
                   variance: str = "beta_tilde", ee_threshold: "float | None" = None, ee_mode: int = 0) -> "Sampler":
    """Sampler for this (models, batch, rule): its buffers and captured CUDA graphs are reused across get_samples()
    calls (the graphs read t, the seed and x from sampler-owned device memory, so they do not depend on the call)."""
    key = (early.handle.value, late.handle.value if late is not None else None, switch_step(t_switch), batch, rule,
           variance, ee_threshold, ee_mode)
    smp = _SAMPLER_CACHE.get(key)
    if smp is None:
        while len(_SAMPLER_CACHE) >= 4:  # bounded: each entry owns a few x-sized device buffers and two graphs
            _SAMPLER_CACHE.pop(next(iter(_SAMPLER_CACHE)))
        smp = _SAMPLER_CACHE[key] = Sampler(early, late, t_switch, batch, rule, variance, ee_threshold, ee_mode)
    return smp


class Sampler:
    """get_samples()' inner loop (sampler.py:128-139 / eesampler.py:57-82) as one C call."""

    def __init__(self, early: Engine, late: Engine | None, t_switch, batch: int, rule: str = "predict_noise",
                 variance: str = "beta_tilde", ee_threshold: float | None = None, ee_mode: int = 0):
        self.lib = _lib.load()
        self.early, self.late, self.batch = early, late, batch
        if isinstance(rule, tuple) and rule[0] == "ddim":  # ("ddim", ddim_steps, ddim_eta)
            table, mode = ddim_coefficients(int(rule[1]), float(rule[2]))
        else:
            table, mode = step_coefficients(rule, variance)
        self.coef = table
        self.device = early.device
        if late is not None and late.device != early.device:
            raise _lib.DuoDiffError(f"early model on {early.device}, late model on {late.device}")
        handle = C.c_void_p()
        # early exit is switched by ee_mode (-1 = off), never by the sign of the threshold: a negative threshold is a
        # legal value of eesampler.py's --threshold (every sample then takes layer 0's head)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ddb_sampler_create(
                early.handle, late.handle if late is not None else None,
                switch_step(t_switch) if late is not None else -1, batch, table.data_ptr(), mode,
                0.0 if ee_threshold is None else float(ee_threshold), -1 if ee_threshold is None else int(ee_mode),
                C.byref(handle)))
        self.handle = handle

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                with torch.cuda.device(self.device):
                    self.lib.ddb_sampler_destroy(h)
            except (AttributeError, TypeError):  # interpreter shutdown: torch is already torn down
                pass
            self.handle = None

    def set_noise_offset(self, first_row: int) -> None:
        """This shard holds rows [first_row, first_row + batch) of a global batch: the in-kernel Philox noise is keyed
        by the GLOBAL element index, so the shards together draw exactly the single-process noise."""
        _lib.check(self.lib.ddb_sampler_set_noise_offset(self.handle, int(first_row)))

    def _check_xy(self, x, y, noise):
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[0] == self.batch
        if x.device != self.device:
            raise _lib.DuoDiffError(f"x is on {x.device} but the sampler's models live on {self.device}")
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous()
            assert noise.shape[0] == 1000 and noise[0].numel() == x.numel()
        if y is not None:
            y = self.early.check_labels(y, self.batch)
            if self.late is not None:
                y = self.late.check_labels(y, self.batch)
        return y

    def profile_step(self, x, t: int, late: bool, y=None) -> dict:
        """Per-kernel-category device time of one eager sampling step as the sampler runs it."""
        y = self._check_xy(x, y, None)
        n = len(Engine.PROF_CATEGORIES)
        ms, cnt = (C.c_float * n)(), (C.c_int32 * n)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ddb_sampler_profile_step(self.handle, x.data_ptr(), _lib.ptr(y), int(t), int(bool(late)),
                                                         ms, cnt, _lib.current_stream_ptr()))
        return {k: dict(ms=float(ms[i]), launches=int(cnt[i])) for i, k in enumerate(Engine.PROF_CATEGORIES)}

    def run(self, x, y=None, noise=None, seed: int = 0, t_first: int = 999, t_last: int = 0, eps_trace=None,
            x_trace=None, exit_log=None, score_log=None, use_graph: bool = True):
        """In place on x [B,C,H,W] f32 cuda.  noise: None (device Philox) or [1000, *x.shape] f32 cuda indexed by t."""
        y = self._check_xy(x, y, noise)
        graph = bool(use_graph) and eps_trace is None and x_trace is None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ddb_sampler_run(
                self.handle, x.data_ptr(), _lib.ptr(y), _lib.ptr(noise), int(seed) & (2**64 - 1), t_first, t_last,
                _lib.ptr(eps_trace), _lib.ptr(x_trace), _lib.ptr(exit_log), _lib.ptr(score_log), int(graph),
                _lib.current_stream_ptr()))
        return x

    def run_list(self, x, t_list, late_flags, y=None, noise=None, seed: int = 0, eps_trace=None, x_trace=None,
                 use_graph: bool = True):
        """The same loop over an explicit timestep list (DDIM): model at t_list[k] (late backbone if late_flags[k]),
        then the sampler's update with the coefficients of t_list[k]."""
        y = self._check_xy(x, y, noise)
        n = len(t_list)
        assert n == len(late_flags)
        ts = (C.c_int32 * n)(*[int(t) for t in t_list])
        lf = (C.c_uint8 * n)(*[1 if f else 0 for f in late_flags])
        graph = bool(use_graph) and eps_trace is None and x_trace is None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ddb_sampler_run_list(
                self.handle, x.data_ptr(), _lib.ptr(y), _lib.ptr(noise), int(seed) & (2**64 - 1), ts, lf, n,
                _lib.ptr(eps_trace), _lib.ptr(x_trace), int(graph), _lib.current_stream_ptr()))
        return x

    def finalize(self, x):
        """(x + 1) / 2, 'b c h w -> b h w c' (sampler.py:145-146)."""
        B, Cc, H, W = x.shape
        out = torch.empty(B, H, W, Cc, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(self.lib.ddb_finalize_nhwc(x.data_ptr(), out.data_ptr(), B, Cc, H, W,
                                                  _lib.current_stream_ptr()))
        return out
