"""Build libmmdiff.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the built
library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmmdiff.so")
SOURCES = ["lib.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newest_mtime(paths):
    return max(os.path.getmtime(p) for p in paths)


def sources():
    files = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".cu"))]
    headers.append(os.path.join(HERE, "..", "include", "mmdiff.h"))
    return files, headers


def build(force: bool = False, verbose: bool = False) -> str:
    files, headers = sources()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_mtime(files + headers):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MMD_NVCC_EXTRA", "").split()
    out = os.environ.get("MMD_LIB_OUT", LIB)
    cmd = [nvcc, *NVCC_FLAGS, *extra, *files, "-o", out]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
