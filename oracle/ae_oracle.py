"""ORACLE — test infrastructure only.  Never imported by the product path (duodiff_b200/).

Functional fp32 restatement of ``FrozenAutoencoderKL.decode`` (models/utils/autoencoder.py:486-490 -> ``Decoder.forward``
:416-449), operating directly on a reference ``state_dict``.  Pinned against the reference itself by
``tests/golden/ae_decode_tiny.npz`` (made by ``tests/golden/make_golden.py`` from the unmodified reference module).

Every function cites the reference lines it restates.  ``decode`` can also return the named intermediate tensors the
CUDA path exposes through ``ddb_ae_decode_debug`` so that a parity failure can be localised to one layer.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class AESpec:
    """ddconfig of get_autoencoder (models/utils/autoencoder.py:503-516); only what the decoder reads."""
    ch: int = 128
    out_ch: int = 3
    ch_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 4
    resolution: int = 256
    embed_dim: int = 4
    scale_factor: float = 0.18215

    @property
    def z_res(self) -> int:  # autoencoder.py:353
        return self.resolution // 2 ** (len(self.ch_mult) - 1)


def _swish(x):  # autoencoder.py:32-34
    return x * torch.sigmoid(x)


def _gn(sd, key, x):  # Normalize: GroupNorm(32, C, eps=1e-6, affine=True), autoencoder.py:37-40
    return F.group_norm(x, 32, sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _conv(sd, key, x, pad):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=1, padding=pad)


def resnet_block(sd, key, x, tap: Optional[dict] = None, name: str = ""):
    """ResnetBlock.forward with temb=None, dropout p=0 (autoencoder.py:116-137)."""
    h = _swish(_gn(sd, key + ".norm1", x))
    if tap is not None:
        tap[name + ".norm1"] = h
    h = _conv(sd, key + ".conv1", h, 1)
    if tap is not None:
        tap[name + ".conv1"] = h
    h = _swish(_gn(sd, key + ".norm2", h))
    if tap is not None:
        tap[name + ".norm2"] = h
    h = _conv(sd, key + ".conv2", h, 1)
    if key + ".nin_shortcut.weight" in sd:  # in_channels != out_channels, conv_shortcut=False (:130-135)
        x = _conv(sd, key + ".nin_shortcut", x, 0)
    elif key + ".conv_shortcut.weight" in sd:
        x = _conv(sd, key + ".conv_shortcut", x, 1)
    out = x + h
    if tap is not None:
        tap[name + (".conv2+nin" if key + ".nin_shortcut.weight" in sd else ".conv2+res")] = out
    return out


def attn_block(sd, key, x, tap: Optional[dict] = None, name: str = ""):
    """AttnBlock.forward (autoencoder.py:165-189): one head over h*w tokens, scale c^-0.5."""
    h_ = _gn(sd, key + ".norm", x)
    if tap is not None:
        tap[name + ".norm"] = h_
    q, k, v = (_conv(sd, key + s, h_, 0) for s in (".q", ".k", ".v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    if tap is not None:
        tap[name + ".pv"] = h_
    out = x + _conv(sd, key + ".proj_out", h_, 0)
    if tap is not None:
        tap[name + ".proj_out+res"] = out
    return out


def decode(sd: Dict[str, torch.Tensor], spec: AESpec, z: torch.Tensor, tap: Optional[dict] = None) -> torch.Tensor:
    """FrozenAutoencoderKL.decode (autoencoder.py:486-490) + Decoder.forward (:416-449).  ``tap`` (a dict) receives the
    NCHW intermediates under the op names of csrc/autoencoder.cu."""
    sd = {k: v.float() for k, v in sd.items()}
    z = (1.0 / spec.scale_factor) * z.float()
    z = _conv(sd, "post_quant_conv", z, 0)
    if tap is not None:
        tap["post_quant_conv"] = z
    h = _conv(sd, "decoder.conv_in", z, 1)
    if tap is not None:
        tap["conv_in"] = h
    h = resnet_block(sd, "decoder.mid.block_1", h, tap, "mid.block_1")
    h = attn_block(sd, "decoder.mid.attn_1", h, tap, "mid.attn_1")
    h = resnet_block(sd, "decoder.mid.block_2", h, tap, "mid.block_2")
    n_levels = len(spec.ch_mult)
    for lvl in reversed(range(n_levels)):
        for j in range(spec.num_res_blocks + 1):
            h = resnet_block(sd, f"decoder.up.{lvl}.block.{j}", h, tap, f"up.{lvl}.block.{j}")
        if lvl != 0:  # Upsample.forward (:52-56): nearest 2x, then 3x3 conv
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", h, 1)
            if tap is not None:
                tap[f"up.{lvl}.upsample"] = h
    h = _swish(_gn(sd, "decoder.norm_out", h))
    if tap is not None:
        tap["norm_out"] = h
    return _conv(sd, "decoder.conv_out", h, 1)


def random_state_dict(spec: AESpec, seed: int, hot: bool = True) -> Dict[str, torch.Tensor]:
    """Decoder-side state_dict with the reference's key names and shapes (Decoder.__init__, autoencoder.py:320-412).
    PyTorch-default Conv2d init scaled so that activations stay O(1) through the residual stack; ``hot`` perturbs
    GroupNorm affine parameters and biases so they are actually exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(key, cout, cin, k, gain=1.0):
        fan = cin * k * k
        sd[key + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (gain / fan ** 0.5)
        sd[key + ".bias"] = torch.randn(cout, generator=g) * (0.1 if hot else 0.0)

    def norm(key, c):
        sd[key + ".weight"] = 1.0 + (torch.randn(c, generator=g) * 0.2 if hot else torch.zeros(c))
        sd[key + ".bias"] = torch.randn(c, generator=g) * 0.2 if hot else torch.zeros(c)

    def res(key, cin, cout):
        norm(key + ".norm1", cin)
        conv(key + ".conv1", cout, cin, 3, 1.4)
        norm(key + ".norm2", cout)
        conv(key + ".conv2", cout, cout, 3, 0.7)
        if cin != cout:
            conv(key + ".nin_shortcut", cout, cin, 1)

    conv("post_quant_conv", spec.z_channels, spec.embed_dim, 1)
    n = len(spec.ch_mult)
    c = spec.ch * spec.ch_mult[-1]
    conv("decoder.conv_in", c, spec.z_channels, 3)
    res("decoder.mid.block_1", c, c)
    norm("decoder.mid.attn_1.norm", c)
    for s in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{s}", c, c, 1, 1.5 if s in ("q", "k") else 1.0)
    res("decoder.mid.block_2", c, c)
    for lvl in reversed(range(n)):
        co = spec.ch * spec.ch_mult[lvl]
        for j in range(spec.num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{j}", c, co)
            c = co
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", c, c, 3)
    norm("decoder.norm_out", c)
    conv("decoder.conv_out", spec.out_ch, c, 3)
    return sd
