"""Shared test helpers: fixture loading and random-init weight factories (no reference import at test time)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"

from duodiff_b200.configs import CONFIGS  # noqa: E402,F401


def load_fixture(name: str):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


def split_fixture(fx: dict, wprefix: str = "w::", pprefix: str = "p::"):
    sd = {k[len(wprefix):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith(wprefix)}
    params = {k[len(pprefix):]: v.item() for k, v in fx.items() if k.startswith(pprefix)}
    return sd, params


def heat_(module: torch.nn.Module, seed: int, scale: float = 4.0) -> None:
    """Make LayerNorm affine / biases / attention matter: the reference init (std 0.02, zero bias, unit LN)
    leaves them untested (SURVEY.md §8c 'hot' variant)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            elif "pos_embed" in name:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
            elif p.dim() == 2 and "label_emb" not in name:
                p.mul_(scale)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
