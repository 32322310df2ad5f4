"""Builds libmrag.so (the C-ABI library of include/mrag.h) in-tree with nvcc for sm_100a.

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmrag.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "550",
]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + [
        os.path.join(HERE, "..", "include", "mrag.h")
    ]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def nvcc_path() -> str | None:
    return shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/mrag.cu -> libmrag.so if the sources are newer than the library."""
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libmrag.so (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("MRAG_NVCC_EXTRA", "").split()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    tmp = LIB + ".tmp"
    cmd += ["-o", tmp, os.path.join(CSRC, "mrag.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libmrag.so")
    if verbose:
        sys.stderr.write(r.stderr)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
