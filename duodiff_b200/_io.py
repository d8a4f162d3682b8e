"""Shared host utilities of the two sampler CLIs: yaml config, checkpoint loading, seeding, PNG dumps."""
from __future__ import annotations

import math
import random
from pathlib import Path

import numpy as np
import torch
import yaml

UVIT_KEYS = ("img_size", "patch_size", "in_chans", "embed_dim", "depth", "num_heads", "mlp_ratio", "qkv_bias",
             "mlp_time_embed", "num_classes", "normalize_timesteps")


def load_config(path) -> dict:
    """utils/config_utils.py:5-13."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"Config file {path} does not exist")
    with path.open("r") as f:
        return yaml.safe_load(f)


def uvit_kwargs(config: dict) -> dict:
    """The explicit keyword set of sampler.py:271-283; extra yaml keys (e.g. the stray `classifier_type` of
    configs/uvit_imagenet64.yaml) are dropped instead of raising TypeError."""
    mp = config["model_params"]
    return {k: mp[k] for k in UVIT_KEYS if k in mp}


def seed_everything(seed: int) -> None:
    """utils/train_utils.py:8-12."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def load_checkpoint_into(model: torch.nn.Module, path) -> None:
    """sampler.py:289-293 — bare state_dict or a training checkpoint carrying `model_state_dict` (Q16)."""
    sd = torch.load(path, map_location="cpu")
    if "model_state_dict" in sd:
        sd = sd["model_state_dict"]
    model.load_state_dict(sd)


def _save_png(path: Path, img01: np.ndarray) -> None:
    """What ``plt.imsave(path, img)`` writes for a float RGB array in [0, 1] (sampler.py:174,184): matplotlib's
    ``to_rgba(bytes=True)`` truncates ``x * 255`` to uint8 and appends an opaque alpha channel (RGBA PNG)."""
    from PIL import Image  # local import: only the CLI needs Pillow
    arr = (np.clip(img01, 0, 1) * 255).astype(np.uint8)
    if arr.ndim == 3 and arr.shape[-1] == 3:
        arr = np.concatenate([arr, np.full(arr.shape[:2] + (1,), 255, np.uint8)], axis=-1)
    elif arr.ndim == 3 and arr.shape[-1] == 1:
        arr = arr[..., 0]
    Image.fromarray(arr).save(path)


def dump_samples(samples, output_folder: Path, timestep: int = 1000) -> None:
    """sampler.py:158-184 — one PNG per sample (clipped to [0,1]) plus grid_image.png."""
    n = len(samples)
    side = math.ceil(math.sqrt(n))
    h, w = samples[0].shape[:2]
    grid = np.zeros((side * h, side * w, 3))
    for i, s in enumerate(samples):
        s = np.clip(s, 0, 1)
        name = f"{i}_{timestep}.png" if timestep != 1000 else f"{i}.png"
        _save_png(Path(output_folder) / name, s)
        r, c = divmod(i, side)
        grid[r * h:(r + 1) * h, c * w:(c + 1) * w, :] = s[..., :3]
    _save_png(Path(output_folder) / "grid_image.png", grid)
