"""In-tree build of the CUDA library (sm_100a only): libsysid_b200.so next to this file.

    python -m system_identification_b200.build [--force] [-v]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.  Every translation unit is
compiled to its own object (system_identification_b200/_obj/, git-ignored) and only stale ones are rebuilt.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libsysid_b200.so")
HEADER = os.path.join("..", "..", "include", "sysid_b200.h")
# translation unit -> the files it depends on
UNITS = {
    "sysid_api.cu": ["sysid_api.cu", "gram_kernels.cuh", "kinematics.cuh", "phases.cuh", "model.cuh", "sdp_kernels.cuh",
                     "filter_kernels.cuh", "gram_tiles.inc", "tmem_park.cuh", "proj_phase.cuh", "bigmodel.cuh", "gram_struct.cuh", "gram_tiles_struct.inc", HEADER],
    "ingest_api.cu": ["ingest_api.cu", "ingest_kernels.cuh", HEADER],
    "extras_api.cu": ["extras_api.cu", "tsqr_kernels.cuh", HEADER],
}
SOURCES = list(UNITS)
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _obj(src):
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("SYSID_NVCC_EXTRA", "").split()      # e.g. -DSYSID_PHASE_CLOCKS for tools/phase_clocks.py
    global OBJ, LIB
    if extra:
        # diagnostic builds never replace the shipped library: libsysid_b200_diag.so, loaded with SYSID_B200_LIB=<path>
        tag = os.environ.get("SYSID_LIB_TAG", "diag")
        OBJ = os.path.join(HERE, "_obj_" + tag)
        LIB = os.path.join(HERE, "libsysid_b200_" + tag + ".so")
        force = True
    os.makedirs(OBJ, exist_ok=True)
    relink = force or not os.path.exists(LIB)
    for src, deps in UNITS.items():
        obj = _obj(src)
        if not force and not _stale(obj, deps):
            relink = relink or os.path.getmtime(obj) > os.path.getmtime(LIB)
            continue
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
        relink = True
    if relink:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [_obj(s) for s in SOURCES] + ["-o", LIB]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
