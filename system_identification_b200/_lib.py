"""ctypes binding of libsysid_b200.so (the C-ABI declared in include/sysid_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a CUDA device is not
available, the calls raise.  PyTorch is used only for device memory, streams and (elsewhere)
torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SYSID_B200_LIB") or os.path.join(_HERE, "libsysid_b200.so")   # the override is for diagnostic builds (build.py)

SYSID_OK = 0
SYSID_ERR_NOT_OPTIMAL = -4
REG_TYPES = {"constant_pullback": 0, "euclidean": 1}


class TreeDesc(C.Structure):
    _fields_ = [("njoints", C.c_int32), ("n_ee", C.c_int32),
                ("parent", C.POINTER(C.c_int32)), ("jtype", C.POINTER(C.c_int32)),
                ("axis", C.POINTER(C.c_double)), ("place_R", C.POINTER(C.c_double)), ("place_p", C.POINTER(C.c_double)),
                ("ee_joint", C.POINTER(C.c_int32)), ("ee_offset", C.POINTER(C.c_double)),
                ("gravity", C.c_double * 3)]


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nq", "nv", "nbodies", "ndof", "nparams", "ncols", "n_ee")]


class Limits(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max_bodies", "max_nv", "max_ee", "max_depth", "max_cols_padded")]


class SdpDesc(C.Structure):
    _fields_ = [("num_links", C.c_int32), ("ndof", C.c_int32),
                ("phi_prior", C.POINTER(C.c_double)), ("semi_axes", C.POINTER(C.c_double)), ("centers", C.POINTER(C.c_double)),
                ("total_mass", C.c_double), ("lambda_reg", C.c_double), ("reg_type", C.c_int32),
                ("epsilon", C.c_double), ("tol", C.c_double), ("max_iters", C.c_int32)]


class SdpInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("refactorizations", C.c_int32), ("reserved", C.c_int32),
                ("primal_residual", C.c_double), ("dual_residual", C.c_double), ("rho", C.c_double), ("objective", C.c_double),
                ("min_eig_J", C.c_double), ("min_eig_C", C.c_double), ("mass_residual", C.c_double)]


class Presolve(C.Structure):
    _fields_ = [("desc", C.POINTER(SdpDesc)), ("plan", C.c_void_p), ("sdp_workspace", C.c_void_p), ("sdp_workspace_bytes", C.c_size_t),
                ("stats_snapshot", C.c_void_p), ("x_scratch", C.c_void_p), ("info_scratch", C.c_void_p), ("warm_out", C.c_void_p),
                ("samples", C.c_int64), ("refine_at", C.c_int64), ("stats_snapshot2", C.c_void_p), ("first_tol", C.c_double)]


SDP_INFO_DTYPE = np.dtype([("status", "<i4"), ("iterations", "<i4"), ("refactorizations", "<i4"), ("reserved", "<i4"),
                           ("primal_residual", "<f8"), ("dual_residual", "<f8"), ("rho", "<f8"), ("objective", "<f8"),
                           ("min_eig_J", "<f8"), ("min_eig_C", "<f8"), ("mass_residual", "<f8")])
assert SDP_INFO_DTYPE.itemsize == C.sizeof(SdpInfo)

_P = C.c_void_p
_SIGNATURES = {
    "sysid_abi_version": (C.c_int, []),
    "sysid_last_error": (C.c_char_p, []),
    "sysid_get_limits": (None, [C.POINTER(Limits)]),
    "sysid_model_create": (C.c_int, [C.POINTER(TreeDesc), C.POINTER(_P)]),
    "sysid_model_destroy": (None, [_P]),
    "sysid_model_dims": (C.c_int, [_P, C.POINTER(Dims)]),
    "sysid_regressor_batch": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P]),
    "sysid_projected_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "sysid_stats_len": (C.c_size_t, [_P, C.c_int32]),
    "sysid_gram_workspace_bytes": (C.c_size_t, [_P]),
    "sysid_gram_accumulate": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "sysid_gram_blocks_workspace_bytes": (C.c_size_t, [_P, C.c_int64, C.c_int64]),
    "sysid_gram_blocks": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "sysid_combine_stats": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int64, _P, _P]),
    "sysid_gram_host_workspace_bytes": (C.c_size_t, [_P, C.c_int64]),
    "sysid_gram_accumulate_host": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_int32, _P, _P, _P, C.c_size_t, C.c_int64, _P]),
    "sysid_gram_accumulate_host_ex": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, C.c_int32, _P, _P, _P, C.c_size_t, C.c_int64, _P]),
    "sysid_gram_accumulate_host_presolve": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, C.c_int32, _P, _P, _P, C.c_size_t, C.c_int64, C.POINTER(Presolve), _P]),
    "sysid_gram_from_stack": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P, C.c_size_t, _P]),
    "sysid_gram_from_stack_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "sysid_filtfilt_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64, C.c_int32]),
    "sysid_filtfilt": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, C.c_size_t, _P]),
    "sysid_savgol_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "sysid_savgol": (C.c_int, [C.c_int32, C.c_int32, _P, _P, C.c_int32, C.c_int64, C.c_int64, _P, C.c_size_t, _P]),
    "sysid_dat_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "sysid_dat_scan": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_size_t, _P, _P]),
    "sysid_dat_parse": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_size_t, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P]),
    "sysid_dat_parse_ex": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_size_t, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P]),
    "sysid_fd_rate": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_double, _P]),
    "sysid_contact_from_tau": (C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, _P]),
    "sysid_round_dat": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _P]),
    "sysid_tsqr_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "sysid_tsqr": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P, C.c_size_t, _P]),
    "sysid_physical_consistency": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "sysid_sdp_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "sysid_sdp_solve": (C.c_int, [C.POINTER(SdpDesc), _P, C.c_int64, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "sysid_sdp_plan_bytes": (C.c_size_t, [C.c_int32]),
    "sysid_sdp_plan_create": (C.c_int, [C.POINTER(SdpDesc), _P, C.c_size_t, _P]),
    "sysid_sdp_solve_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "sysid_sdp_warm_len": (C.c_size_t, [C.c_int32, C.c_int32]),
    "sysid_sdp_solve_plan": (C.c_int, [C.POINTER(SdpDesc), _P, _P, C.c_int64, C.c_int32, _P, _P, _P, C.c_size_t, _P, _P, _P]),
    "sysid_predict_rmse": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    "sysid_predict_rmse_workspace_bytes": (C.c_size_t, [_P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (building is NOT attempted here: run __graft_entry__.build() or
    `python -m system_identification_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m system_identification_b200.build` "
                           "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SysidError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"sysid_b200 error {code}: {message}")
        self.code = code


def check(rc):
    if rc != SYSID_OK:
        raise SysidError(rc, load().sysid_last_error().decode("utf-8", "replace"))


def _dptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def create_model(flat):
    """FlatModel -> opaque sysid_model* (as a c_void_p).  Keeps nothing alive: the library copies the tree."""
    lib = load()
    parent = np.ascontiguousarray(flat.parent, dtype=np.int32)
    jtype = np.ascontiguousarray(np.where(np.asarray(flat.jtype) < 0, 0, flat.jtype), dtype=np.int32)
    axis = np.ascontiguousarray(flat.axis, dtype=np.float64).reshape(-1)
    pR = np.ascontiguousarray(flat.place_R, dtype=np.float64).reshape(-1)
    pp = np.ascontiguousarray(flat.place_p, dtype=np.float64).reshape(-1)
    n_ee = len(flat.ee_names)
    eej = np.ascontiguousarray(flat.ee_joint if n_ee else np.zeros(0), dtype=np.int32)
    eeo = np.ascontiguousarray(flat.ee_offset if n_ee else np.zeros(0), dtype=np.float64).reshape(-1)
    d = TreeDesc()
    d.njoints = int(parent.shape[0]); d.n_ee = n_ee
    d.parent = _dptr(parent, C.c_int32); d.jtype = _dptr(jtype, C.c_int32)
    d.axis = _dptr(axis, C.c_double); d.place_R = _dptr(pR, C.c_double); d.place_p = _dptr(pp, C.c_double)
    d.ee_joint = _dptr(eej, C.c_int32) if n_ee else None
    d.ee_offset = _dptr(eeo, C.c_double) if n_ee else None
    d.gravity = (C.c_double * 3)(*[float(g) for g in flat.gravity])
    handle = _P()
    check(lib.sysid_model_create(C.byref(d), C.byref(handle)))
    return handle
