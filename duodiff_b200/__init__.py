"""duodiff_b200 — B200-native (sm_100a) implementation of DuoDiff's sampling hot path.

Host-side mirror of the reference's module surface (``UViT``, ``EarlyExitUViT``, ``sampler.get_samples``,
``eesampler.get_samples``) over a C-ABI CUDA library (include/duodiff_b200.h).  No CPU / PyTorch compute fallback.
"""
from .early_exit import AttentionProbe, EarlyExitUViT, MLPProbe, OutputHead  # noqa: F401
from .uvit import UViT  # noqa: F401

__all__ = ["UViT", "EarlyExitUViT", "OutputHead", "MLPProbe"]
