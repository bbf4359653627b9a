"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU).

    python -m scfgp_b200.build [--force]

Produces ``scfgp_b200/lib/libscfgp_b200.so`` (git-ignored, travels with gpurun snapshots).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libscfgp_b200.so")
SOURCES = ["scfgp_abi.cu", "scaler_abi.cu"]
DEPS = ["scfgp_abi.cu", "scaler_abi.cu", "kernels.cuh", "gemm_dmma.cuh", "gemm_tma.cuh", "kernels_tma.cuh", "umma_i8.cuh",
        "emu_kernels.cuh", "mma_peak.cuh", "host_state.cuh", "host_util.cuh", "host_buffers.cuh", "host_fp64.cuh", "host_emu.cuh",
        "host_dense.cuh", "host_guard.cuh", os.path.join("..", "..", "include", "scfgp_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-diag-suppress", "177",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libscfgp_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
