"""In-tree build of ``lib/libtecgat.so`` with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tec_mollm_b200.build [--force] [--verbose]

One object per ``csrc/*.cu`` (compiled in parallel), linked into a plain shared library with a C ABI -- no torch,
no pybind: Python binds it with ctypes (``_lib.py``).
"""
from __future__ import annotations

import argparse
import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libtecgat.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xcompiler", "-fvisibility=default",
    "-I", CSRC, "-I", os.path.join(ROOT, "include"),
]


# experiment switches (e.g. TECGAT_NVCC_EXTRA="-DTG_SW_FAKE_OWN"): part of the digest, so a plain build afterwards rebuilds
EXTRA = os.environ.get("TECGAT_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libtecgat.so")
    return exe


def _digest(paths) -> str:
    """Content hash of the sources (mtimes do not survive the snapshot to the GPU box)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS[:8] + EXTRA).encode())
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    stamp = LIB + ".sha256"
    want = _digest(sources + headers)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return LIB  # up to date
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        # incremental: an object is rebuilt only when its source or any header changed
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        ostamp, owant = obj + ".sha256", _digest([src] + headers)
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read().strip() == owant:
            return obj
        cmd = [nvcc] + NVCC_FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(ostamp, "w") as f:
            f.write(owant)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
