#!/usr/bin/env python
"""`python fid.py --dataset D --samples_path DIR` -- the reference's FID evaluation CLI (fid.py:8-56) for samples written
by `sampler.py` / `eesampler.py`.  FID sits AFTER the accelerated path: it needs the pretrained
InceptionV3 of `torchmetrics` and the real datasets, neither of which ships with this repository, so this file only
keeps the reference's command line working where those dependencies exist and fails loudly where they do not:

  * `read_samples` restates `utils/evaluation_utils.py:13-24` (every `*.png` below the folder except the grid image,
    RGB, float32 in [0, 1]);
  * `fid_evaluation` calls `torchmetrics.image.fid.FrechetInceptionDistance(normalize=True)` exactly like fid.py:34-40;
  * the real images come from `--real_path DIR` (a folder of PNGs; an addition, since the reference's dataset loaders
    `datasets/*.py` are training-side code outside the sampling path) or, when the reference's `utils.evaluation_utils`
    is importable, from `get_dataset_samples(dataset, data_path, seed, n)` (fid.py:47-49);
  * `frechet_distance` is the closed form between two Gaussians the metric ends in, for callers that bring their own
    feature statistics.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import torch


def get_args(argv=None):
    p = argparse.ArgumentParser(description="FID evaluation parameters")
    p.add_argument("--dataset", type=str, required=True, choices=["cifar10", "celeba", "imagenet64", "imagenet256"],
                   help="Dataset name.")
    p.add_argument("--seed", type=int, default=0, help="Seed for sampling images from the dataset.")
    p.add_argument("--data_path", type=str, default="data", help="Directory for datasets")
    p.add_argument("--samples_path", type=str, required=True, help="Path to the directory with samples.")
    p.add_argument("--real_path", type=str, default=None,
                   help="Folder of real PNG images (instead of the reference's dataset loaders)")
    return p.parse_args(argv)


def read_samples(path) -> torch.Tensor:
    """utils/evaluation_utils.py:13-24 -> [N, 3, H, W] float32 in [0, 1]."""
    from PIL import Image
    imgs = []
    for p in sorted(Path(path).rglob("*.png")):
        if "grid" in p.name:
            continue
        arr = np.asarray(Image.open(p).convert("RGB"), dtype=np.float32) / 255.0
        imgs.append(torch.from_numpy(arr).permute(2, 0, 1))
    if not imgs:
        raise FileNotFoundError(f"no sample PNGs below {path}")
    out = torch.stack(imgs, dim=0)
    print(f"Read {len(out)} images")
    return out


def frechet_distance(mu1, sigma1, mu2, sigma2) -> float:
    """|mu1 - mu2|^2 + tr(S1 + S2 - 2 (S1 S2)^{1/2}), float64; the trace of the matrix square root is the sum of the
    square roots of the eigenvalues of S1 S2 (real and non-negative for covariance matrices)."""
    mu1, mu2 = np.asarray(mu1, np.float64), np.asarray(mu2, np.float64)
    s1, s2 = np.asarray(sigma1, np.float64), np.asarray(sigma2, np.float64)
    eig = np.linalg.eigvals(s1 @ s2)
    tr_sqrt = np.sqrt(np.clip(eig.real, 0.0, None)).sum()
    d = mu1 - mu2
    return float(d @ d + np.trace(s1) + np.trace(s2) - 2.0 * tr_sqrt)


def fid_evaluation(real_images, generated_images):
    """fid.py:34-40."""
    try:
        from torchmetrics.image.fid import FrechetInceptionDistance
    except ImportError as e:
        raise RuntimeError("fid.py needs `torchmetrics` (and its pretrained InceptionV3 weights), like the reference's "
                           "fid.py:3; it is not part of the sampling path and is not bundled") from e
    fid = FrechetInceptionDistance(normalize=True)
    fid.update(real_images, real=True)
    fid.update(generated_images, real=False)
    print("Evaluating FID")
    value = float(fid.compute())
    print(f"FID: {value}")
    return value


def main(argv=None):
    args = get_args(argv)
    generated = read_samples(args.samples_path)
    n = len(generated)
    print(f"Using {n}")
    if args.real_path:
        real = read_samples(args.real_path)[:n]
    else:
        try:
            from utils.evaluation_utils import get_dataset_samples  # the reference's loaders, when on sys.path
        except ImportError as e:
            raise RuntimeError("no --real_path given and the reference's dataset loaders (utils/evaluation_utils.py, "
                               "datasets/*.py) are not importable") from e
        real = get_dataset_samples(args.dataset, args.data_path, args.seed, n)
    return fid_evaluation(real, generated)


if __name__ == "__main__":
    main()
