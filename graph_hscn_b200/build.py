"""Builds libghscn.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library is compiled straight from graph_hscn_b200/csrc/*.cu -- no torch headers, no JIT cache --
so the built .so travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libghscn.so"
STAMP = LIB_DIR / "libghscn.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",            # default; bit-exact paths use explicit __fmul_rn/__fadd_rn
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "-Xptxas", "-v",
    "-shared",
]


EXTRA = os.environ.get("GHSCN_NVCC_EXTRA", "").split()   # e.g. -DGHSCN_WIDE_ROWS=32 for tuning experiments


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [REPO_ROOT / "include" / "ghscn.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + EXTRA).encode())
    return h.hexdigest()


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libghscn.so")
    return nvcc


def build_library(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, *EXTRA, "-I", str(REPO_ROOT / "include"), "-I", str(CSRC),
           "-o", str(LIB_PATH), *map(str, _sources())]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    (LIB_DIR / "build.log").write_text(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libghscn.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    STAMP.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
