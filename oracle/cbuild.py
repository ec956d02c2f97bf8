"""ORACLE (test infrastructure) -- build + ctypes binding of oracle/sysid_oracle.c (the C restatement used as the
CPU baseline).  `python -m oracle.cbuild` compiles it into oracle/_build/libsysid_oracle.so (git-ignored, travels
to the GPU box like the product's .so)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sysid_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libsysid_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # no -march=native: the GPU box may have a different CPU than the build container
    cmd = ["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return LIB


class _Tree(C.Structure):
    _fields_ = [("njoints", C.c_int32), ("n_ee", C.c_int32),
                ("parent", C.POINTER(C.c_int32)), ("jtype", C.POINTER(C.c_int32)),
                ("axis", C.POINTER(C.c_double)), ("place_R", C.POINTER(C.c_double)), ("place_p", C.POINTER(C.c_double)),
                ("ee_joint", C.POINTER(C.c_int32)), ("ee_offset", C.POINTER(C.c_double)),
                ("gravity", C.c_double * 3)]


class COracle:
    """tree: oracle.urdf_tree.Tree; ee_names: end-effector frame names in contact-channel order."""

    def __init__(self, tree, ee_names):
        self.lib = C.CDLL(build())
        self.tree = tree
        self._keep = []
        t = _Tree()
        def arr(a, dt, ct):
            a = np.ascontiguousarray(a, dtype=dt); self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(ct))
        t.njoints = tree.njoints; t.n_ee = len(ee_names)
        t.parent = arr(tree.parent, np.int32, C.c_int32)
        t.jtype = arr(np.where(np.asarray(tree.jtype) < 0, 0, tree.jtype), np.int32, C.c_int32)
        t.axis = arr(np.asarray(tree.axis).reshape(-1), np.float64, C.c_double)
        t.place_R = arr(np.asarray(tree.place_R).reshape(-1), np.float64, C.c_double)
        t.place_p = arr(np.asarray(tree.place_p).reshape(-1), np.float64, C.c_double)
        t.ee_joint = arr([tree.frames[n][0] for n in ee_names], np.int32, C.c_int32)
        t.ee_offset = arr(np.array([tree.frames[n][2] for n in ee_names]).reshape(-1), np.float64, C.c_double)
        t.gravity = (C.c_double * 3)(*[float(g) for g in tree.gravity])
        self.t = t
        self.nv, self.nb = tree.nv, tree.njoints - 1
        self.nd = self.nb - 1
        self.lib.oracle_gram_accumulate.restype = C.c_int
        self.lib.oracle_tau_rmse.restype = C.c_int

    def _p(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def gram(self, q, dq, ddq, tau, cnt, friction=True, nthreads=0):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (q, dq, ddq, tau, cnt)]
        N = arrs[0].shape[1]
        c = 10 * self.nb + (2 * self.nd if friction else 0)
        stats = np.zeros(c * c + c + 2)
        used = self.lib.oracle_gram_accumulate(C.byref(self.t), *[self._p(a) for a in arrs], C.c_int64(N), C.c_int64(N),
                                               C.c_int(1 if friction else 0), C.c_int(nthreads), self._p(stats))
        return stats, used

    def tau_rmse(self, q, dq, ddq, tau, cnt, phi, nthreads=0):
        """(total mean-square, per-joint RMSE) of reference print_tau_prediction_rmse (src/sys_identification.py:421-437)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (q, dq, ddq, tau, cnt)]
        N = arrs[0].shape[1]
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        assert phi.size == 10 * self.nb
        out = np.zeros(1 + self.nd)
        self.lib.oracle_tau_rmse(C.byref(self.t), *[self._p(a) for a in arrs], C.c_int64(N), C.c_int64(N), self._p(phi),
                                 C.c_int(nthreads), self._p(out))
        return float(out[0]), out[1:].copy()

    def blocks(self, q, dq, ddq, tau, cnt, friction=True):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (q, dq, ddq, tau, cnt)]
        N = arrs[0].shape[1]
        c = 10 * self.nb + (2 * self.nd if friction else 0)
        A = np.zeros((N, self.nv, c)); b = np.zeros((N, self.nv)); Y = np.zeros((N, self.nv, 10 * self.nb)); P = np.zeros((N, self.nv, self.nv))
        self.lib.oracle_sample_blocks(C.byref(self.t), *[self._p(a) for a in arrs], C.c_int64(N), C.c_int64(N),
                                      C.c_int(1 if friction else 0), self._p(A), self._p(b), self._p(Y), self._p(P))
        return A, b, Y, P


if __name__ == "__main__":
    print(build(force=True))
