"""In-tree build of the native libraries (nvcc / g++), no JIT cache.

    python -m gid_b200.build [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    gid_b200/lib/libgidcuda.so   CUDA kernels + C-ABI shim        (nvcc, sm_100a)
    gid_b200/lib/libgidhost.so   C++ host mirror of GID's JPEG API (g++)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "lib")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unknown-pragmas", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(out: str, deps: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def _sources(d: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libgidcuda.so")
    deps = _sources(CSRC, (".cu", ".cuh", ".h", ".hpp")) + [os.path.join(ROOT, "include", "gidcuda.h")]
    if force or _stale(out, deps):
        srcs = _sources(CSRC, (".cu",))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", out]
        subprocess.run(cmd, check=True, cwd=CSRC)
    return out


def build_host(force: bool = False) -> str | None:
    os.makedirs(LIB, exist_ok=True)
    if not os.path.isdir(HOST):
        return None
    srcs = _sources(HOST, (".cpp",))
    if not srcs:
        return None
    out = os.path.join(LIB, "libgidhost.so")
    deps = _sources(HOST, (".cpp", ".hpp", ".h")) + [os.path.join(ROOT, "include", "gidcuda.h")]
    if force or _stale(out, deps):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra",
               "-I", os.path.join(ROOT, "include")] + srcs + ["-o", out, "-L", LIB, "-lgidcuda",
               "-Wl,-rpath,$ORIGIN", "-lpthread"]
        subprocess.run(cmd, check=True, cwd=HOST)
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", ", ".join(sorted(os.listdir(LIB))))
