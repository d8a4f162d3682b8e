"""Drop-in for the reference's ``models/uvit.py`` module interface (UViT ctor, state_dict layout, forward).

Same constructor signature as models/uvit.py:229-247, same parameter tree (so reference checkpoints load with
``load_state_dict`` — SURVEY.md Q16), same ``forward(x, timesteps, y=None) -> eps``.  The parameters are only a
container: ``forward`` hands them to the hand-written sm_100a kernels behind the C ABI (include/duodiff_b200.h).
There is no PyTorch compute path; without the CUDA library and a B200 the forward raises.
"""
from __future__ import annotations

import torch
from torch import nn

from .engine import Engine


def _linear(i: int, o: int, bias: bool = True) -> nn.Linear:
    return nn.Linear(i, o, bias=bias)


class _Attn(nn.Module):  # parameter names of models/uvit.py:150-153
    def __init__(self, dim: int, qkv_bias: bool):
        super().__init__()
        self.qkv = _linear(dim, 3 * dim, qkv_bias)
        self.proj = _linear(dim, dim)


class _Mlp(nn.Module):  # models/uvit.py:82-84
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = _linear(dim, hidden)
        self.fc2 = _linear(hidden, dim)


class _Block(nn.Module):  # models/uvit.py:185-198
    def __init__(self, dim: int, hidden: int, qkv_bias: bool, long_skip: bool):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _Attn(dim, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, hidden)
        self.skip_linear = _linear(2 * dim, dim) if long_skip else None


class _PatchEmbed(nn.Module):  # models/uvit.py:214-219
    def __init__(self, patch_size: int, in_chans: int, embed_dim: int):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class UViT(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                 num_classes, normalize_timesteps, qk_scale=None, norm_layer=nn.LayerNorm, mlp_time_embed=False,
                 use_checkpoint=False, conv=True, skip=True, max_batch: int | None = None):
        super().__init__()
        # qk_scale: accepted and ignored, like the reference (its attention always uses SDPA's default 1/sqrt(head_dim),
        # models/uvit.py:158-163); use_checkpoint: a training-time memory trade, no effect on the forward.
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("duodiff_b200.UViT: norm_layer != nn.LayerNorm has no kernel")
        del qk_scale, use_checkpoint
        self.num_features = self.embed_dim = embed_dim
        self.normalize_timesteps = normalize_timesteps
        self.num_classes = num_classes
        self.in_chans = in_chans
        self.depth = depth
        self.img_size, self.patch_size, self.num_heads = img_size, patch_size, num_heads
        self.mlp_hidden = int(embed_dim * mlp_ratio)
        self.num_patches = (img_size // patch_size) ** 2
        self.patch_dim = patch_size ** 2 * in_chans
        self.extras = 2 if num_classes > 0 else 1
        self.max_batch = max_batch

        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        # models/uvit.py:264-272 (the library serves the MLP from a table of the 1000 integer timesteps)
        self.time_embed = (nn.Sequential(_linear(embed_dim, 4 * embed_dim), nn.SiLU(), _linear(4 * embed_dim, embed_dim))
                           if mlp_time_embed else nn.Identity())
        self.label_emb = nn.Embedding(num_classes, embed_dim) if num_classes > 0 else None
        self.pos_embed = nn.Parameter(torch.zeros(1, self.extras + self.num_patches, embed_dim))
        half = depth // 2
        mk = lambda long_skip: _Block(embed_dim, self.mlp_hidden, qkv_bias, long_skip and skip)  # noqa: E731
        self.in_blocks = nn.ModuleList([mk(False) for _ in range(half)])
        self.mid_block = mk(False)
        self.out_blocks = nn.ModuleList([mk(True) for _ in range(half)])
        self.norm = nn.LayerNorm(embed_dim)
        self.decoder_pred = _linear(embed_dim, self.patch_dim)
        self.final_layer = nn.Conv2d(in_chans, in_chans, 3, padding=1) if conv else nn.Identity()
        self._reset_parameters()
        self._engine: Engine | None = None
        self._engine_key = None

    # initialisation of models/uvit.py:335-345: trunc-normal(0.02) Linear weights / pos_embed, zero biases,
    # unit LayerNorm; Conv2d layers keep PyTorch's default init.
    def _reset_parameters(self):
        nn.init.trunc_normal_(self.pos_embed, std=0.02, a=-2.0, b=2.0)
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.trunc_normal_(mod.weight, std=0.02, a=-2.0, b=2.0)
                if mod.bias is not None:
                    nn.init.zeros_(mod.bias)
            elif isinstance(mod, nn.LayerNorm):
                nn.init.ones_(mod.weight)
                nn.init.zeros_(mod.bias)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed"}

    @property
    def device(self):
        return next(self.parameters()).device

    # ------------------------------------------------------------------ engine management
    def engine_kwargs(self, early_exit: bool = False) -> dict:
        return dict(img_size=self.img_size, patch_size=self.patch_size, in_chans=self.in_chans,
                    embed_dim=self.embed_dim, depth=self.depth, num_heads=self.num_heads,
                    mlp_hidden=self.mlp_hidden, num_classes=self.num_classes,
                    normalize_timesteps=self.normalize_timesteps, early_exit=early_exit)

    def _weights_key(self, batch: int):
        return (max(batch, self.max_batch or 0),
                tuple((p.data_ptr(), p._version) for p in self.parameters()))

    def engine(self, batch: int) -> Engine:
        """(Re)build the C-side model when the parameters or the batch capacity changed."""
        key = self._weights_key(batch)
        if (self._engine is None or self._engine_key[1] != key[1] or self._engine.max_batch < batch):
            self._engine = None
            self._engine = Engine(self.state_dict(), max_batch=key[0], **self.engine_kwargs())
            self._engine_key = key
        return self._engine

    def invalidate(self):
        self._engine = None

    def forward(self, x, timesteps, y=None):
        """models/uvit.py:351-383 on the B200 kernels. x [B,C,H,W] f32 cuda, timesteps [B], y [B] int64 | None."""
        if self.label_emb is not None and y is None:
            raise ValueError("class-conditional UViT needs y (the reference crashes here: 257 vs 258 tokens)")
        with torch.no_grad():
            return self.engine(x.shape[0]).forward(x, timesteps, y if self.label_emb is not None else None)
