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


def _compile_one(nvcc: str, src: Path, obj: Path, header_fp: str) -> str:
    """Compiles one translation unit unless its object is up to date; returns the compiler output."""
    stamp = obj.with_suffix(".stamp")
    h = hashlib.sha256()
    h.update(src.read_bytes())
    h.update(header_fp.encode())
    fp = h.hexdigest()
    if obj.exists() and stamp.exists() and stamp.read_text().strip() == fp:
        return ""
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    cmd = [nvcc, *flags, *EXTRA, "-I", str(REPO_ROOT / "include"), "-I", str(CSRC), "-c", "-o", str(obj), str(src)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = " ".join(cmd) + "\n" + proc.stdout + proc.stderr
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed compiling {src.name}")
    stamp.write_text(fp)
    return log


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """One object per csrc/*.cu (compiled in parallel, recompiled only when the file, a header or the flags changed),
    linked into lib/libghscn.so."""
    from concurrent.futures import ThreadPoolExecutor
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB_PATH
    nvcc = find_nvcc()
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + [REPO_ROOT / "include" / "ghscn.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + EXTRA).encode())
    header_fp = h.hexdigest()
    srcs = _sources()
    objs = [obj_dir / (s.stem + ".o") for s in srcs]
    if force:
        for o in objs:
            o.with_suffix(".stamp").unlink(missing_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(lambda so: _compile_one(nvcc, so[0], so[1], header_fp), zip(srcs, objs)))
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB_PATH),
            *map(str, objs)]
    proc = subprocess.run(link, capture_output=True, text=True)
    old_log = (LIB_DIR / "build.log").read_text() if (LIB_DIR / "build.log").exists() and not force else ""
    new_log = "".join(logs)
    (LIB_DIR / "build.log").write_text((new_log or old_log) + " ".join(link) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed linking libghscn.so")
    if verbose:
        sys.stderr.write(new_log)
    STAMP.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
