"""In-tree build of the CUDA library (sm_100a only): libsysid_b200.so next to this file.

    python -m system_identification_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsysid_b200.so")
SOURCES = ["sysid_api.cu"]
DEPS = ["sysid_api.cu", "gram_kernels.cuh", "kinematics.cuh", "phases.cuh", "model.cuh", "sdp_kernels.cuh", "filter_kernels.cuh", "gram_tiles.inc",
        os.path.join("..", "..", "include", "sysid_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("SYSID_NVCC_EXTRA", "").split()      # e.g. -DSYSID_PHASE_CLOCKS for tools/phase_clocks.py
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
