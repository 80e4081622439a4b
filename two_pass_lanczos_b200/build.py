"""In-tree build of libtplanczos.so (hand-written sm_100a CUDA + host C++), no JIT cache involved.

    python -m two_pass_lanczos_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels with the
repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtplanczos.so")
SOURCES = ["tpl_engine.cu", "tpl_loader.cpp", "tpl_ftk.cpp"]
HEADERS = ["tpl_kernels.cuh", "tpl_dense.cuh", "tpl_cells.cuh", "tpl_cells_host.h", "tpl_tiles.cuh", "tpl_tiles_host.h", "tpl_blocks.cuh", "tpl_blocks_host.h", "tpl_build.cuh", "tpl_csr.cuh", "tpl_sharded.cuh", "tpl_internal.h", os.path.join("..", "..", "include", "tplanczos.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtplanczos.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
