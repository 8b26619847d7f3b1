"""Builds the CUDA shared library ``turbomesh_b200/libturbomesh_gpu.so`` for sm_100a (B200) in-tree.

nvcc cross-compiles without a GPU, so this also runs on the CPU-only build box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "turbomesh_gpu.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "kernels.cuh"), os.path.join(HERE, "csrc", "krylov_kernels.cuh"), os.path.join(HERE, "csrc", "krylov.inl"), os.path.join(HERE, "csrc", "krylov_coarse.cuh"), os.path.join(HERE, "csrc", "krylov_phased.cuh"), os.path.join(HERE, "csrc", "mg_kernels.cuh"), os.path.join(HERE, "csrc", "io_kernels.cuh"), os.path.join(HERE, "csrc", "topology.hpp"), os.path.join(HERE, "csrc", "partition.hpp"), os.path.join(HERE, "csrc", "mg_plan.hpp"), os.path.join(HERE, "csrc", "multigrid.inl"), os.path.join(HERE, "csrc", "exchange.inl"), os.path.join(HERE, "csrc", "streamed.inl"),
        os.path.join(HERE, "..", "include", "turbomesh_gpu.h")]
OUT = os.path.join(HERE, "libturbomesh_gpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
