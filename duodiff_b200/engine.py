"""Owner of one ``ddb_model`` handle: hands the reference ``state_dict`` to the C library and launches forwards.

PyTorch is used for device memory and streams only; all arithmetic happens inside libduodiff_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class Engine:
    """One re-packed U-ViT (or early-exit U-ViT) on one CUDA device: the device of the state_dict's tensors (or, for
    CPU tensors, the current device).  Every call switches to that device, so a model on cuda:1 works while cuda:0 is
    current; inputs on another device are rejected."""

    def __init__(self, state_dict: dict, *, img_size: int, patch_size: int, in_chans: int, embed_dim: int,
                 depth: int, num_heads: int, mlp_hidden: int, num_classes: int, normalize_timesteps: bool,
                 early_exit: int, max_batch: int, ln_eps: float = 1e-5):
        """early_exit: 0 = plain U-ViT, 1 / 2 / 3 = EarlyExitUViT with MLP probes per layer / per timestep / per layer and
        timestep (ddb_uvit_config.early_exit)."""
        if not torch.cuda.is_available():
            raise _lib.DuoDiffError("duodiff_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = _lib.UViTConfig(img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_hidden,
                                   num_classes, int(bool(normalize_timesteps)), int(early_exit), max_batch,
                                   ln_eps)
        self.in_chans, self.img_size, self.depth, self.max_batch = in_chans, img_size, depth, max_batch
        self.early_exit = bool(early_exit)
        self.num_classes = num_classes
        devs = {t.device for t in state_dict.values() if t.is_cuda}
        if len(devs) > 1:
            raise _lib.DuoDiffError(f"state_dict tensors live on several devices: {sorted(map(str, devs))}")
        dev = devs.pop() if devs else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        with torch.cuda.device(dev):
            keep, arr = [], (_lib.Tensor * len(state_dict))()
            for i, (name, t) in enumerate(state_dict.items()):
                t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                keep.append(t)
                arr[i] = _lib.Tensor(name.encode(), t.data_ptr(), t.numel())
            torch.cuda.synchronize(dev)
            handle = C.c_void_p()
            _lib.check(self.lib.ddb_model_create(C.byref(self.cfg), arr, len(state_dict), C.byref(handle)))
            self.handle = handle
            del keep  # the library holds its own re-packed copies

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                with torch.cuda.device(self.device):
                    self.lib.ddb_model_destroy(h)
            except (AttributeError, TypeError):  # interpreter shutdown: torch is already torn down
                pass
            self.handle = None

    # ------------------------------------------------------------------ single forwards
    def _check_inputs(self, x, timesteps, y):
        if not (x.is_cuda and x.dtype == torch.float32):
            raise _lib.DuoDiffError("x must be a CUDA float32 tensor [B,C,H,W]")
        if x.device != self.device:
            raise _lib.DuoDiffError(f"x is on {x.device} but the model lives on {self.device}")
        B = x.shape[0]
        if tuple(x.shape[1:]) != (self.in_chans, self.img_size, self.img_size):
            raise _lib.DuoDiffError(f"x has shape {tuple(x.shape)}, model expects [B,{self.in_chans},"
                                    f"{self.img_size},{self.img_size}]")
        x = x.contiguous()
        t = timesteps.to(device=x.device, dtype=torch.float32).contiguous()
        if t.numel() != B:
            raise _lib.DuoDiffError("timesteps must have one entry per sample")
        if y is not None:
            y = self.check_labels(y, B)
        return x, t, y, B

    def check_labels(self, y, B: int):
        """Labels index the embedding table (models/uvit.py:361-363): like nn.Embedding, out-of-range labels raise
        IndexError instead of reading outside the table."""
        y = y.to(device=self.device, dtype=torch.int64).contiguous()
        if y.numel() != B:
            raise _lib.DuoDiffError("y must have one label per sample")
        if self.num_classes > 0 and B > 0:
            lo, hi = int(y.min()), int(y.max())
            if lo < 0 or hi >= self.num_classes:
                raise IndexError(f"label {lo if lo < 0 else hi} is out of range for {self.num_classes} classes")
        return y

    def forward(self, x, timesteps, y=None):
        """UViT.forward (models/uvit.py:351-383)."""
        x, t, y, B = self._check_inputs(x, timesteps, y)
        eps = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ddb_uvit_forward(self.handle, x.data_ptr(), t.data_ptr(), _lib.ptr(y), B,
                                                 eps.data_ptr(), _lib.current_stream_ptr()))
        return eps

    def ee_forward(self, x, timesteps, y=None, threshold: float = 0.0, mode: int = 0, want_all: bool = True):
        """EarlyExitUViT.forward + eesampler selection. Returns (eps_selected, exit_idx, scores, outputs)."""
        x, t, y, B = self._check_inputs(x, timesteps, y)
        eps = torch.empty_like(x)
        idx = torch.empty(B, device=x.device, dtype=torch.int32)
        scores = torch.empty(self.depth, B, device=x.device) if want_all else None
        # per-layer head outputs only exist in simulate mode (compaction skips every head but the exit layer's)
        outputs = torch.empty(self.depth + 1, *x.shape, device=x.device) if (want_all and mode == 0) else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ddb_ee_forward(self.handle, x.data_ptr(), t.data_ptr(), _lib.ptr(y), B,
                                               float(threshold), mode, eps.data_ptr(), idx.data_ptr(),
                                               _lib.ptr(scores), _lib.ptr(outputs), _lib.current_stream_ptr()))
        return eps, idx, scores, outputs

    PROF_CATEGORIES = ("embed", "ln_stats", "gemm_qkv", "attention", "gemm_proj", "gemm_fc1", "gemm_fc2", "gemm_skip",
                       "gemm_decode", "conv", "ee_other", "ddpm", "tail")

    def profile_forward(self, x, timesteps, y=None, ee: bool = False) -> dict:
        """Per-kernel-category device time of one forward (CUDA events on the launching stream)."""
        x, t, y, B = self._check_inputs(x, timesteps, y)
        eps = torch.empty_like(x)
        n = len(self.PROF_CATEGORIES)
        ms = (C.c_float * n)()
        cnt = (C.c_int32 * n)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ddb_profile_forward(self.handle, x.data_ptr(), t.data_ptr(), _lib.ptr(y), B,
                                                    eps.data_ptr(), int(ee), ms, cnt, _lib.current_stream_ptr()))
        return {k: dict(ms=float(ms[i]), launches=int(cnt[i])) for i, k in enumerate(self.PROF_CATEGORIES)}
