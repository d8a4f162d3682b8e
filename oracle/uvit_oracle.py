"""ORACLE — test infrastructure only.  Never imported by the product path (duodiff_b200/).

A functional, fp32, plain-PyTorch restatement of the reference's sampling hot path, operating directly on a
reference ``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
arm may import this package, and only as the checker.

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` (in the build container), runs them on seeded random-init weights and commits inputs, weights and
outputs as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them.

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class UViTSpec:
    """model_params of configs/*.yaml (models/uvit.py:229-247)."""
    img_size: int
    patch_size: int
    in_chans: int
    embed_dim: int
    depth: int
    num_heads: int
    mlp_ratio: float = 4
    qkv_bias: bool = False
    num_classes: int = -1
    normalize_timesteps: bool = True

    @property
    def extras(self) -> int:  # models/uvit.py:274-279
        return 2 if self.num_classes > 0 else 1

    @property
    def num_patches(self) -> int:  # models/uvit.py:262
        return (self.img_size // self.patch_size) ** 2

    @property
    def patch_dim(self) -> int:  # models/uvit.py:327
        return self.patch_size ** 2 * self.in_chans

    @classmethod
    def from_params(cls, params: dict) -> "UViTSpec":
        keys = cls.__dataclass_fields__.keys()
        return cls(**{k: v for k, v in params.items() if k in keys})


# ------------------------------------------------------------------------------------------------ U-ViT pieces
def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """models/uvit.py:95-115 — [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(max_period) i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def patch_embed(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, p: int) -> torch.Tensor:
    """models/uvit.py:221-225 — stride-p conv == per-patch linear over (C, p1, p2); tokens in (h, w) order."""
    B, C, H, W = x.shape
    D = weight.shape[0]
    patches = x.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), C * p * p)
    return patches @ weight.reshape(D, C * p * p).t() + bias


def unpatchify(tok: torch.Tensor, channels: int) -> torch.Tensor:
    """models/uvit.py:125-132 — 'B (h w) (p1 p2 C) -> B C (h p1) (w p2)' (channel innermost in the token)."""
    B, N, pd = tok.shape
    p = int(round((pd // channels) ** 0.5))
    h = w = int(round(N ** 0.5))
    assert h * w == N and p * p * channels == pd
    return tok.reshape(B, h, w, p, p, channels).permute(0, 5, 1, 3, 2, 4).reshape(B, channels, h * p, w * p)


def attention(sd: dict, pfx: str, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    """models/uvit.py:155-168 — qkv split as (K=3, H, hd); softmax(q k^T / sqrt(hd)) v; proj."""
    B, L, D = x.shape
    hd = D // num_heads
    qkv = F.linear(x, sd[pfx + "qkv.weight"], sd.get(pfx + "qkv.bias"))
    qkv = qkv.reshape(B, L, 3, num_heads, hd).permute(2, 0, 3, 1, 4).float()
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd)), dim=-1) @ v
    out = att.permute(0, 2, 1, 3).reshape(B, L, D)
    return F.linear(out, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])


def block(sd: dict, pfx: str, x: torch.Tensor, skip: torch.Tensor | None, num_heads: int) -> torch.Tensor:
    """models/uvit.py:203-208 — [skip_linear(cat(x, skip))]; x += attn(LN1 x); x += fc2(gelu_erf(fc1(LN2 x)))."""
    D = x.shape[-1]
    if pfx + "skip_linear.weight" in sd:
        x = F.linear(torch.cat([x, skip], dim=-1), sd[pfx + "skip_linear.weight"], sd[pfx + "skip_linear.bias"])
    h = F.layer_norm(x, (D,), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], 1e-5)
    x = x + attention(sd, pfx + "attn.", h, num_heads)
    h = F.layer_norm(x, (D,), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], 1e-5)
    h = F.linear(h, sd[pfx + "mlp.fc1.weight"], sd[pfx + "mlp.fc1.bias"])
    h = F.gelu(h)  # nn.GELU() default: exact erf form (models/uvit.py:74,88)
    x = x + F.linear(h, sd[pfx + "mlp.fc2.weight"], sd[pfx + "mlp.fc2.bias"])
    return x


def output_head(sd: dict, pfx: str, x: torch.Tensor, spec: UViTSpec) -> torch.Tensor:
    """models/uvit.py:377-382 and models/early_exit.py:22-28 — LN, Linear(D->pd), drop extras, unpatchify, 3x3 conv."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[pfx + "norm.weight"], sd[pfx + "norm.bias"], 1e-5)
    h = F.linear(h, sd[pfx + "decoder_pred.weight"], sd[pfx + "decoder_pred.bias"])
    h = unpatchify(h[:, spec.extras:, :], spec.in_chans)
    if pfx + "final_layer.weight" not in sd:  # conv=False: final_layer is nn.Identity() (models/uvit.py:329-333)
        return h
    return F.conv2d(h, sd[pfx + "final_layer.weight"], sd[pfx + "final_layer.bias"], padding=1)


def embed_tokens(sd: dict, pfx: str, spec: UViTSpec, x: torch.Tensor, timesteps: torch.Tensor,
                 y: torch.Tensor | None) -> torch.Tensor:
    """models/uvit.py:352-365 — patch tokens, time token in front, optional label token in front of that, + pos."""
    if spec.normalize_timesteps:
        timesteps = timesteps.float() / 1000
    tok = patch_embed(x, sd[pfx + "patch_embed.proj.weight"], sd[pfx + "patch_embed.proj.bias"], spec.patch_size)
    time_token = timestep_embedding(timesteps, spec.embed_dim)
    if pfx + "time_embed.0.weight" in sd:  # mlp_time_embed=True (models/uvit.py:264-272): Linear -> SiLU -> Linear
        time_token = F.linear(F.silu(F.linear(time_token, sd[pfx + "time_embed.0.weight"], sd[pfx + "time_embed.0.bias"])),
                              sd[pfx + "time_embed.2.weight"], sd[pfx + "time_embed.2.bias"])
    tok = torch.cat((time_token.unsqueeze(1), tok), dim=1)
    if y is not None and (pfx + "label_emb.weight") in sd:
        tok = torch.cat((sd[pfx + "label_emb.weight"][y].unsqueeze(1), tok), dim=1)
    return tok + sd[pfx + "pos_embed"]


def _block_prefixes(spec: UViTSpec, pfx: str):
    half = spec.depth // 2
    return ([f"{pfx}in_blocks.{i}." for i in range(half)] + [f"{pfx}mid_block."]
            + [f"{pfx}out_blocks.{i}." for i in range(half)])


def uvit_forward(sd: dict, spec: UViTSpec, x: torch.Tensor, timesteps: torch.Tensor, y: torch.Tensor | None = None,
                 pfx: str = "", capture: list | None = None) -> torch.Tensor:
    """models/uvit.py:351-383.  ``capture`` (optional list) receives the hidden state after every block."""
    h = embed_tokens(sd, pfx, spec, x, timesteps, y)
    if capture is not None:
        capture.append(h)
    half = spec.depth // 2
    skips = []
    for i, bp in enumerate(_block_prefixes(spec, pfx)):
        if i < half:
            h = block(sd, bp, h, None, spec.num_heads)
            skips.append(h)
        elif i == half:
            h = block(sd, bp, h, None, spec.num_heads)
        else:
            h = block(sd, bp, h, skips.pop(), spec.num_heads)
        if capture is not None:
            capture.append(h)
    return output_head(sd, pfx, h, spec)


def mlp_probe(sd: dict, i, x: torch.Tensor) -> torch.Tensor:
    """models/early_exit.py:31-37 — mean over all tokens of sigmoid(Linear(D,1)); i = key of EarlyExitUViT.matrix."""
    w, b = sd[f"matrix.{i}.classifier.0.weight"], sd[f"matrix.{i}.classifier.0.bias"]
    return torch.sigmoid(F.linear(x, w, b)).mean(dim=1).squeeze()


def attention_probe(sd: dict, i: int, x: torch.Tensor) -> torch.Tensor:
    """models/early_exit.py:40-80 (num_heads = 1) — a learned query attends over x[:, 1:], the pooled value goes
    through Linear -> SiLU -> Linear(D, 1); explicit softmax instead of F.scaled_dot_product_attention."""
    p = f"matrix.{i}."
    xs = x[:, 1:, :]                                                      # :73 "ignore time vector"
    kv = F.linear(xs, sd[p + "weight_kv.weight"], sd[p + "weight_kv.bias"])
    D = xs.shape[-1]
    k, v = kv[..., :D], kv[..., D:]                                       # "b l (k h hd) -> k b h l hd", k=2, h=1
    q = sd[p + "q"].reshape(1, 1, D)
    att = torch.softmax((q @ k.transpose(-2, -1)) / math.sqrt(D), dim=-1)  # [B, 1, L-1]
    pooled = att @ v                                                      # [B, 1, D]
    h = F.silu(F.linear(pooled, sd[p + "classification.0.weight"], sd[p + "classification.0.bias"]))
    return F.linear(h, sd[p + "classification.2.weight"], sd[p + "classification.2.bias"]).squeeze()


def probe_key(classifier_type: str, i: int, t: int) -> str:
    """models/early_exit.py:194-204 (get_classifer): which entry of ``matrix`` scores layer i at timestep t."""
    if classifier_type in ("mlp_probe_per_layer", "attention_probe"):
        return f"{i}"
    if classifier_type == "mlp_probe_per_timestep":
        return f"{t}"
    if classifier_type == "mlp_probe_per_layer_per_timestep":
        return f"{i}, {t}"
    raise ValueError(f"Unknown classifier type: {classifier_type}")


def ee_forward(sd: dict, spec: UViTSpec, x: torch.Tensor, timesteps: torch.Tensor, y: torch.Tensor | None = None,
               classifier_type: str = "mlp_probe_per_layer"):
    """models/early_exit.py:268-320 (MLP probe types; t = int(timesteps[0]) picks the probes of the timestep-indexed
    layouts, :269) -> (eps, [cls_i], [out_i])."""
    pfx = "uvit."
    t_int = int(timesteps[0])
    h = embed_tokens(sd, pfx, spec, x, timesteps, y)
    half = spec.depth // 2
    head_pfx = ([f"in_blocks_heads.{i}." for i in range(half)] + ["mid_block_head."]
                + [f"out_blocks_heads.{i}." for i in range(half)])
    skips, cls, outs = [], [], []
    for i, bp in enumerate(_block_prefixes(spec, pfx)):
        outs.append(output_head(sd, head_pfx[i], h, spec))
        if classifier_type == "attention_probe":
            cls.append(attention_probe(sd, i, h))
        else:
            cls.append(mlp_probe(sd, probe_key(classifier_type, i, t_int), h))
        if i < half:
            h = block(sd, bp, h, None, spec.num_heads)
            skips.append(h)
        elif i == half:
            h = block(sd, bp, h, None, spec.num_heads)
        else:
            h = block(sd, bp, h, skips.pop(), spec.num_heads)
    return output_head(sd, pfx, h, spec), cls, outs


def ee_select(eps_full: torch.Tensor, cls: list, outs: list, threshold: float):
    """eesampler.py:62-68 — first layer with probe <= threshold (else full model); gather its output."""
    outputs = torch.stack(outs + [eps_full])
    scores = torch.stack([c.reshape(-1) for c in cls] + [torch.zeros_like(cls[0].reshape(-1))])
    indices = torch.argmax((scores <= threshold).int(), dim=0)
    B = eps_full.shape[0]
    return outputs[indices, torch.arange(B, device=indices.device)], indices, scores


# ------------------------------------------------------------------------------------------------ DDPM
def ddpm_schedule(device="cpu") -> dict:
    """sampler.py:40-44 == eesampler.py:33-37 == ddpm_core.py:64-70 (fp32, same torch calls)."""
    betas = torch.linspace(1e-4, 0.02, 1000).to(device)
    alphas = 1 - betas
    alphas_bar = torch.cumprod(alphas, dim=0)
    alphas_bar_previous = torch.cat([torch.tensor([1.0], device=device), alphas_bar[:-1]])
    betas_tilde = betas * (1 - alphas_bar_previous) / (1 - alphas_bar)
    return dict(betas=betas, alphas=alphas, alphas_bar=alphas_bar, alphas_bar_previous=alphas_bar_previous,
                betas_tilde=betas_tilde)


def predict_noise_step(sch: dict, model_output, x, t: int, z):
    """sampler.py:47-56 / eesampler.py:74-82; z is the injected N(0,I) tensor (ignored at t == 0 where z = 0)."""
    alpha_t, alpha_bar_t = sch["alphas"][t], sch["alphas_bar"][t]
    sigma_t = torch.sqrt(sch["betas_tilde"][t])
    z = z if t > 0 else 0
    return (torch.sqrt(1 / alpha_t) * (x - (1 - alpha_t) / (torch.sqrt(1 - alpha_bar_t)) * model_output)) + sigma_t * z


def predict_original_step(sch: dict, model_output, x, t: int, z):
    """sampler.py:59-72."""
    alpha_t, alpha_bar_t = sch["alphas"][t], sch["alphas_bar"][t]
    abp, beta_t = sch["alphas_bar_previous"][t], sch["betas"][t]
    sigma_t = torch.sqrt(sch["betas_tilde"][t])
    z = z if t > 0 else 0
    return (torch.sqrt(abp) * beta_t * model_output / (1 - alpha_bar_t)
            + torch.sqrt(alpha_t) * (1 - abp) * x / (1 - alpha_bar_t)) + sigma_t * z


def predict_previous_step(sch: dict, model_output, x, t: int, z):
    """sampler.py:75-79."""
    sigma_t = torch.sqrt(sch["betas_tilde"][t])
    z = z if t > 0 else 0
    return model_output + sigma_t * z


STEP_RULES = {"predict_noise": predict_noise_step, "predict_original": predict_original_step,
              "predict_previous": predict_previous_step}


def sample_ddpm(early, late, t_switch, x_T: torch.Tensor, noise, y=None, rule: str = "predict_noise",
                t_first: int = 999, t_last: int = 0, trace: dict | None = None) -> torch.Tensor:
    """sampler.py:128-139 — the DDPM loop with the DuoDiff hand-off (:135-136).

    ``early`` / ``late`` are callables (x, timesteps, y) -> model output.  ``noise`` is either a tensor
    [1000, *x.shape] indexed by t (injected noise) or a callable t -> z (e.g. torch.randn_like drawing from the
    global generator, which reproduces the reference's RNG stream).  Returns x_0 (before the (x+1)/2 epilogue)."""
    sch = ddpm_schedule(x_T.device)
    step = STEP_RULES[rule]
    x, model = x_T, early
    for t in range(t_first, t_last - 1, -1):
        time_tensor = t * torch.ones(x.shape[0], device=x.device)
        with torch.no_grad():
            out = model(x, time_tensor, y)
        z = None
        if t > 0:
            z = noise(t) if callable(noise) else noise[t]
        if trace is not None:
            trace.setdefault("x_in", []).append(x)
            trace.setdefault("eps", []).append(out)
        x = step(sch, out, x, t, z)
        if t == 1000 - t_switch:
            model = late
    return x


def ddim_timesteps(ddim_steps: int):
    """sampler.py:104 — np.linspace(0, 999, ddim_steps).astype(int)[::-1] as a list of python ints."""
    import numpy as np
    return [int(v) for v in np.linspace(0, 999, ddim_steps).astype(int)[::-1]]


def ddim_step(sch: dict, model_output, x, t: int, s: int, eta: float, z):
    """sampler.py:112-120 (with the reference's quirk: the noise term is sigma_t^2 * z, not sigma_t * z)."""
    ab = sch["alphas_bar"]
    sigma_t_squared = sch["betas_tilde"][t] * eta
    mean = torch.sqrt(ab[s] / ab[t]) * (x - torch.sqrt(1 - ab[t]) * model_output)
    mean = mean + torch.sqrt(1 - ab[s] - sigma_t_squared) * model_output
    return mean + sigma_t_squared * z if z is not None else mean


def sample_ddim(early, late, t_switch, x_T: torch.Tensor, noise, ddim_steps: int, eta: float, y=None,
                trace: dict | None = None, n_pairs: int | None = None) -> torch.Tensor:
    """sampler.py:103-126 — the DDIM branch: strided timesteps, z = 0 for the last pair (s == 0), and the hand-off
    `if t < 1000 - t_switch: model = late_model` AFTER the step at t (so the first step below the boundary still runs
    on the early model).  `noise` as in sample_ddpm (indexed by t).  n_pairs truncates the loop (tests)."""
    sch = ddpm_schedule(x_T.device)
    ts = ddim_timesteps(ddim_steps)
    x, model = x_T, early
    pairs = list(zip(ts[:-1], ts[1:]))
    for t, s in pairs[:n_pairs]:
        time_tensor = t * torch.ones(x.shape[0], device=x.device)
        with torch.no_grad():
            out = model(x, time_tensor, y)
        z = None
        if s > 0:
            z = noise(t) if callable(noise) else noise[t]
        if trace is not None:
            trace.setdefault("t", []).append(t)
            trace.setdefault("late", []).append(model is late)
        x = ddim_step(sch, out, x, t, s, eta, z)
        if late is not None and t < 1000 - t_switch:
            model = late
    return x


def to_samples_nhwc(x: torch.Tensor) -> torch.Tensor:
    """sampler.py:145-146 — (x + 1) / 2, 'b c h w -> b h w c' (un-clipped)."""
    return ((x + 1) / 2).permute(0, 2, 3, 1).contiguous()


def ee_sample(ee_model, threshold: float, depth: int, x_T: torch.Tensor, noise, y=None, t_first: int = 999,
              t_last: int = 0, trace: dict | None = None):
    """eesampler.py:57-82 — returns (x_0, error_prediction_by_timestep [1000, depth], indices_by_timestep [1000, B]).
    ``trace`` (tests): per step the input x_t and the per-sample probe outputs [depth+1, B], keyed by t."""
    sch = ddpm_schedule(x_T.device)
    B = x_T.shape[0]
    err_log = torch.zeros(1000, depth)
    idx_log = torch.zeros(1000, B)
    x = x_T
    for t in range(t_first, t_last - 1, -1):
        time_tensor = t * torch.ones(B, device=x.device)
        with torch.no_grad():
            eps_full, cls, outs = ee_model(x, time_tensor, y)
        eps, indices, scores = ee_select(eps_full, cls, outs, threshold)
        err_log[t] = scores.mean(axis=1)[:depth]
        idx_log[t, :] = indices
        if trace is not None:
            trace.setdefault("x_in", {})[t] = x
            trace.setdefault("scores", {})[t] = scores
        z = None
        if t > 0:
            z = noise(t) if callable(noise) else noise[t]
        x = predict_noise_step(sch, eps, x, t, z)
    return x, err_log, idx_log
