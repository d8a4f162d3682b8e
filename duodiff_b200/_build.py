"""In-tree build of libduodiff_b200.so (nvcc, sm_100a only). The .so is git-ignored but travels with gpurun."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libduodiff_b200.so"
SOURCES = ["duodiff_b200.cu", "autoencoder.cu"]
HEADERS = ["ptx.cuh", "gemm.cuh", "gemm2.cuh", "attention.cuh", "elementwise.cuh", "conv_gemm.cuh", "host_common.h",
           "experimental/gemm3.cuh", "experimental/attention2.cuh", "experimental/attention_mma.cuh",
           "experimental/embed_tokens.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def needs_rebuild() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [PKG_DIR.parent / "include" / "duodiff_b200.h"]
    return any(d.stat().st_mtime > built for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False, experimental: bool | None = None) -> Path:
    """Compile the CUDA library for sm_100a. Cross-compiles without a GPU.

    DDB_EXPERIMENTAL=1 (or experimental=True) also compiles the measured-and-rejected kernel variants under
    csrc/experimental/ (A-in-TMEM GEMM, two-threads-per-row attention, generic mma.sync attention, fp32-FMA token assembly,
    single-CTA GEMM for the block linears); the product library does not contain them."""
    if experimental is None:
        experimental = os.environ.get("DDB_EXPERIMENTAL", "0") not in ("", "0")
    if not force and not needs_rebuild() and not experimental:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, *(["-DDDB_EXPERIMENTAL"] if experimental else []), "-o", str(LIB_PATH),
           *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
