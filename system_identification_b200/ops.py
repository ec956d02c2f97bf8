"""Device-level operators: thin, allocation-explicit wrappers over the C-ABI for torch CUDA tensors.

Every function here launches hand-written sm_100a kernels through libsysid_b200.so on the current
torch CUDA stream.  Inputs are channel-major fp64 CUDA tensors of shape (channels, N), exactly the
arrays the reference's read_data produces (reference demo/solo_identification.py:9-33).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import FlatModel


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("system_identification_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chan(t, channels, name):
    if t.dtype != torch.float64 or not t.is_cuda or t.dim() != 2 or t.shape[0] != channels or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a CUDA float64 tensor of shape ({channels}, N) with unit inner stride, got "
                         f"{tuple(t.shape)} {t.dtype} {t.device}")
    return t


def to_device(a, device=None):
    """numpy/torch (channels, N) of any float dtype -> contiguous CUDA fp64 (exact widening of float32 logs)."""
    _require_cuda()
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    return a.to(device=device or "cuda", dtype=torch.float64, non_blocking=True).contiguous()


class DeviceModel:
    """Owns the immutable library handle for one FlatModel (reference: the pinocchio model/data pair
    built in SystemIdentification.__init__, src/sys_identification.py:16-22)."""

    def __init__(self, flat: FlatModel):
        _require_cuda()
        self.flat = flat
        self.lib = _lib.load()
        self.handle = _lib.create_model(flat)
        d = _lib.Dims()
        _lib.check(self.lib.sysid_model_dims(self.handle, C.byref(d)))
        self.nq, self.nv, self.nb, self.nd, self.nparams, self.n_ee = d.nq, d.nv, d.nbodies, d.ndof, d.nparams, d.n_ee
        self._ws = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.sysid_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def ncols(self, friction=True):
        return self.nparams + (2 * self.nd if friction else 0)

    def stats_len(self, friction=True):
        c = self.ncols(friction)
        return c * c + c + 2

    def _workspace(self, key, nbytes, device):
        buf = self._ws.get((key, device))
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=device)
            self._ws[(key, device)] = buf
        return buf

    # ---------------------------------------------------------------------------------- stage 1
    def regressor_batch(self, q, dq, ddq):
        """pin.computeJointTorqueRegressor for every column: returns (N, nv, 10*nb)."""
        q = _chan(q, self.nq, "q"); dq = _chan(dq, self.nv, "dq"); ddq = _chan(ddq, self.nv, "ddq")
        N = q.shape[1]
        ld = self._common_ld(q, dq, ddq)
        Y = torch.empty((N, self.nv, self.nparams), dtype=torch.float64, device=q.device)
        _lib.check(self.lib.sysid_regressor_batch(self.handle, _ptr(q), _ptr(dq), _ptr(ddq), N, ld, _ptr(Y), _stream()))
        return Y

    def projected_batch(self, q, dq, ddq, tau, cnt, friction=True, want_P=False):
        """Per-sample [P Y | P S^T diag(dq) | P S^T diag(sign dq)] (N, nv, c), P S^T tau (N, nv) and optionally P."""
        q = _chan(q, self.nq, "q"); dq = _chan(dq, self.nv, "dq"); ddq = _chan(ddq, self.nv, "ddq")
        tau = _chan(tau, self.nd, "tau"); cnt = _chan(cnt, self.n_ee, "contact")
        N = q.shape[1]
        ld = self._common_ld(q, dq, ddq, tau, cnt)
        c = self.ncols(friction)
        A = torch.empty((N, self.nv, c), dtype=torch.float64, device=q.device)
        b = torch.empty((N, self.nv), dtype=torch.float64, device=q.device)
        P = torch.empty((N, self.nv, self.nv), dtype=torch.float64, device=q.device) if want_P else None
        _lib.check(self.lib.sysid_projected_batch(self.handle, _ptr(q), _ptr(dq), _ptr(ddq), _ptr(tau), _ptr(cnt), N, ld,
                                                  1 if friction else 0, _ptr(A), _ptr(b), _ptr(P), _stream()))
        return (A, b, P) if want_P else (A, b)

    # ---------------------------------------------------------------------------------- stage 2
    def gram_accumulate(self, q, dq, ddq, tau, cnt, friction=True, weights=None, stats=None, info=None):
        """Fused regressor -> projector -> Gram.  Returns stats = [G (c x c) | r (c) | s | n]; ADDS into `stats` if given."""
        q = _chan(q, self.nq, "q"); dq = _chan(dq, self.nv, "dq"); ddq = _chan(ddq, self.nv, "ddq")
        tau = _chan(tau, self.nd, "tau"); cnt = _chan(cnt, self.n_ee, "contact")
        N = q.shape[1]
        ld = self._common_ld(q, dq, ddq, tau, cnt)
        if stats is None:
            stats = torch.zeros(self.stats_len(friction), dtype=torch.float64, device=q.device)
        if weights is not None and (weights.dtype != torch.float64 or weights.numel() != N or not weights.is_contiguous()):
            raise ValueError("weights: expected contiguous float64 of length N")
        nbytes = self.lib.sysid_gram_workspace_bytes(self.handle)
        ws = self._workspace("gram", nbytes, q.device)
        _lib.check(self.lib.sysid_gram_accumulate(self.handle, _ptr(q), _ptr(dq), _ptr(ddq), _ptr(tau), _ptr(cnt), N, ld,
                                                  _ptr(weights), 1 if friction else 0, _ptr(stats), _ptr(info), _ptr(ws),
                                                  ws.numel(), _stream()))
        return stats

    def gram_blocks(self, q, dq, ddq, tau, cnt, block, friction=True, info=None):
        """One statistics vector per block of `block` consecutive samples, in ONE launch of the fused kernel (segmented mode):
        returns (K, stats_len) with K = ceil(N / block).  Block bootstrap: resample statistics = weights @ this (combine_stats)."""
        q = _chan(q, self.nq, "q"); dq = _chan(dq, self.nv, "dq"); ddq = _chan(ddq, self.nv, "ddq")
        tau = _chan(tau, self.nd, "tau"); cnt = _chan(cnt, self.n_ee, "contact")
        N = q.shape[1]
        ld = self._common_ld(q, dq, ddq, tau, cnt)
        block = int(block)
        if block < 1:
            raise ValueError("block must be >= 1")
        K = (N + block - 1) // block
        slen = self.stats_len(friction)
        out = torch.empty((K, slen), dtype=torch.float64, device=q.device)
        ws = self._workspace("gram_blocks", self.lib.sysid_gram_blocks_workspace_bytes(self.handle, N, block), q.device)
        _lib.check(self.lib.sysid_gram_blocks(self.handle, _ptr(q), _ptr(dq), _ptr(ddq), _ptr(tau), _ptr(cnt), N, ld, block,
                                              1 if friction else 0, _ptr(out), slen, _ptr(info), _ptr(ws), ws.numel(), _stream()))
        return out

    def gram_accumulate_host(self, q, dq, ddq, tau, cnt, friction=True, weights=None, stats=None, info=None, chunk=131072,
                             device=None, presolve=None, presolve_samples=0, presolve_refine_at=0, presolve_tol=None, presolve_first_tol=None):
        """gram_accumulate for arrays still in HOST memory: float64 OR float32 torch CPU tensors / numpy arrays, channel-major
        with unit inner stride (pinned memory for full PCIe speed) -- exactly what the reference's read_data returns
        (float32 q / contact, float64 dq / ddq / tau; demo/solo_identification.py:10-33).  The upload is chunked and
        overlapped with the kernel inside the library (sysid_gram_accumulate_host_presolve); float32 arrays cross PCIe as
        float32 and are widened exactly on the device.  presolve: an SdpPlan whose LMI problem is solved on the first chunk's
        statistics behind the rest of the stream (its record warm-starts the final solve).  Returns the device statistics."""
        hs = []
        for a, ch, name in ((q, self.nq, "q"), (dq, self.nv, "dq"), (ddq, self.nv, "ddq"), (tau, self.nd, "tau"), (cnt, self.n_ee, "contact")):
            t = torch.from_numpy(a) if isinstance(a, np.ndarray) else a
            if t.is_cuda or t.dtype not in (torch.float64, torch.float32) or t.dim() != 2 or t.shape[0] != ch or t.stride(1) != 1:
                raise ValueError(f"{name}: expected a host float64/float32 array of shape ({ch}, N) with unit inner stride")
            hs.append(t)
        N = hs[0].shape[1]
        if any(t.shape[1] != N for t in hs):
            raise ValueError("all sample arrays must have the same number of columns")
        device = torch.device(device or "cuda")
        if stats is None:
            stats = torch.zeros(self.stats_len(friction), dtype=torch.float64, device=device)
        wh = None
        if weights is not None:
            wh = torch.from_numpy(weights) if isinstance(weights, np.ndarray) else weights
            if wh.is_cuda or wh.dtype != torch.float64 or wh.numel() != N or not wh.is_contiguous():
                raise ValueError("weights: expected contiguous host float64 of length N")
        chunk = int(max(1, min(chunk, N)))
        nbytes = self.lib.sysid_gram_host_workspace_bytes(self.handle, chunk)
        ws = self._workspace("gram_host", nbytes, device)
        self._keepalive = (hs, wh)        # the host arrays must outlive the asynchronous copies
        ptrs = (C.c_void_p * 5)(*[t.data_ptr() if t.shape[0] > 0 else None for t in hs])
        dts = (C.c_int32 * 5)(*[0 if t.dtype == torch.float64 else 1 for t in hs])
        lds = (C.c_int64 * 5)(*[max(t.stride(0), N) if t.shape[0] > 1 else N for t in hs])
        pre = None
        if presolve is not None:
            # presolve: an SdpPlan -- the LMI fit of the first chunk is solved behind the stream and left in presolve.warm
            pre = presolve.presolve_struct(presolve_samples, presolve_refine_at, presolve_tol, presolve_first_tol)
        _lib.check(self.lib.sysid_gram_accumulate_host_presolve(self.handle, ptrs, dts, lds, N,
                                                                C.c_void_p(wh.data_ptr()) if wh is not None else None,
                                                                1 if friction else 0, _ptr(stats), _ptr(info), _ptr(ws), ws.numel(), chunk,
                                                                C.byref(pre) if pre is not None else None, _stream()))
        return stats

    def predict_rmse(self, q, dq, ddq, tau, cnt, phi):
        """(total mean-square, per-joint RMSE) of reference print_tau_prediction_rmse (src/sys_identification.py:421-437)."""
        q = _chan(q, self.nq, "q"); dq = _chan(dq, self.nv, "dq"); ddq = _chan(ddq, self.nv, "ddq")
        tau = _chan(tau, self.nd, "tau"); cnt = _chan(cnt, self.n_ee, "contact")
        N = q.shape[1]
        ld = self._common_ld(q, dq, ddq, tau, cnt)
        phi = phi.to(device=q.device, dtype=torch.float64).contiguous()
        if phi.numel() != self.nparams:
            raise ValueError(f"phi: expected {self.nparams} parameters")
        out = torch.empty(1 + self.nd, dtype=torch.float64, device=q.device)
        nbytes = self.lib.sysid_predict_rmse_workspace_bytes(self.handle)
        ws = self._workspace("rmse", nbytes, q.device)
        _lib.check(self.lib.sysid_predict_rmse(self.handle, _ptr(q), _ptr(dq), _ptr(ddq), _ptr(tau), _ptr(cnt), N, ld, _ptr(phi),
                                               _ptr(out), _ptr(ws), ws.numel(), _stream()))
        return out

    @staticmethod
    def _common_ld(*ts):
        ld = ts[0].stride(0)
        N = ts[0].shape[1]
        for t in ts:
            if t.shape[1] != N:
                raise ValueError("all sample arrays must have the same number of columns")
            if t.stride(0) != ld:
                raise ValueError("all sample arrays must share one leading dimension (slice them from equally-shaped parents)")
        return ld


def gram_from_stack(A, b, stats=None):
    """stats of an already stacked system: A (rows, c), b (rows) CUDA fp64 (reference Solver(regressor, tau_vec, ...))."""
    _require_cuda()
    lib = _lib.load()
    if A.dtype != torch.float64 or not A.is_cuda or A.dim() != 2 or not A.is_contiguous():
        raise ValueError("A: expected contiguous CUDA float64 (rows, c)")
    rows, c = A.shape
    b = b.to(device=A.device, dtype=torch.float64).contiguous()
    if b.numel() != rows:
        raise ValueError("b: length must equal the rows of A")
    if stats is None:
        stats = torch.zeros(c * c + c + 2, dtype=torch.float64, device=A.device)
    nbytes = lib.sysid_gram_from_stack_workspace_bytes(c)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    _lib.check(lib.sysid_gram_from_stack(_ptr(A), _ptr(b), rows, c, _ptr(stats), _ptr(ws), ws.numel(), _stream()))
    return stats


def combine_stats(weights, stats_blocks):
    """(B, K) multiplicities x (K, stats_len) per-block statistics -> (B, stats_len) resample statistics, one launch on the fp64
    tensor pipe (sysid_combine_stats)."""
    _require_cuda()
    lib = _lib.load()
    if weights.dtype != torch.float64 or not weights.is_cuda or weights.dim() != 2 or not weights.is_contiguous():
        raise ValueError("weights: expected contiguous CUDA float64 (B, K)")
    if stats_blocks.dtype != torch.float64 or not stats_blocks.is_cuda or stats_blocks.dim() != 2 or not stats_blocks.is_contiguous():
        raise ValueError("stats_blocks: expected contiguous CUDA float64 (K, stats_len)")
    B, K = weights.shape
    if stats_blocks.shape[0] != K:
        raise ValueError("weights and stats_blocks disagree on the number of blocks")
    out = torch.empty((B, stats_blocks.shape[1]), dtype=torch.float64, device=weights.device)
    _lib.check(lib.sysid_combine_stats(_ptr(weights), B, K, _ptr(stats_blocks), stats_blocks.shape[1], _ptr(out), _stream()))
    return out


def tsqr(A, b=None):
    """Triangular factor of the stacked system [A | b] (A (rows, c), b (rows) CUDA fp64): returns the (c+1, c+1) upper
    triangle [R z; 0 rho] (sysid_tsqr).  Used by Solver.solve_llsq_svd (reference src/solver.py:32-39)."""
    _require_cuda()
    lib = _lib.load()
    if A.dtype != torch.float64 or not A.is_cuda or A.dim() != 2 or not A.is_contiguous():
        raise ValueError("A: expected contiguous CUDA float64 (rows, c)")
    rows, c = A.shape
    if b is not None:
        b = b.to(device=A.device, dtype=torch.float64).contiguous()
        if b.numel() != rows:
            raise ValueError("b: length must equal the rows of A")
    R = torch.empty((c + 1, c + 1), dtype=torch.float64, device=A.device)
    nbytes = lib.sysid_tsqr_workspace_bytes(c)
    if nbytes == 0:
        raise ValueError(f"tsqr: c = {c} is outside the compiled envelope")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    _lib.check(lib.sysid_tsqr(_ptr(A), _ptr(b), rows, c, _ptr(R), _ptr(ws), ws.numel(), _stream()))
    return R


def llsq_svd_from_triangle(R_aug, rcond=1e-15):
    """Minimum-norm least-squares solution from the TSQR triangle: SVD of the small c x c R on the device (library call on
    a 130 x 130 matrix), singular values <= rcond * sigma_max dropped (np.linalg.pinv's rule, reference src/solver.py:37)."""
    c = R_aug.shape[0] - 1
    R, z = R_aug[:c, :c], R_aug[:c, c]
    U, S, Vh = torch.linalg.svd(R, full_matrices=False)
    Sinv = torch.where(S > rcond * S.max(), 1.0 / S, torch.zeros_like(S))
    return Vh.T @ (Sinv * (U.T @ z)), S


def physical_consistency(phi, num_links, ellipsoids):
    """get_physical_consistency (reference src/sys_identification.py:324-389) for a batch of parameter vectors on the
    device.  phi: (batch, >= 10 num_links) CUDA fp64 (or 1-D).  Returns a CUDA tensor (batch, 5, num_links):
    min eig I_bar, min eig I (6x6), min eig J, min eig C, tr(J Q)."""
    _require_cuda()
    lib = _lib.load()
    phi = phi if phi.dim() == 2 else phi.unsqueeze(0)
    if phi.dtype != torch.float64 or not phi.is_cuda or phi.stride(1) != 1 or phi.shape[1] < 10 * num_links:
        raise ValueError("phi: expected CUDA float64 (batch, >= 10 * num_links) with unit inner stride")
    sa = torch.tensor(np.array([e["semi_axes"] for e in ellipsoids], dtype=np.float64).reshape(-1), device=phi.device)
    ce = torch.tensor(np.array([e["center"] for e in ellipsoids], dtype=np.float64).reshape(-1), device=phi.device)
    if sa.numel() != 3 * num_links or ce.numel() != 3 * num_links:
        raise ValueError("bounding_ellipsoids do not match num_links")
    out = torch.empty((phi.shape[0], 5, num_links), dtype=torch.float64, device=phi.device)
    _lib.check(lib.sysid_physical_consistency(_ptr(phi), phi.stride(0) if phi.shape[0] > 1 else phi.shape[1], phi.shape[0], num_links, _ptr(sa), _ptr(ce), _ptr(out), _stream()))
    return out


def sdp_solve(stats, num_links, ndof, phi_prior, ellipsoids, total_mass, lambda_reg=1e-1, tol=1e-10, max_iters=0,
              reg_type="constant_pullback", epsilon=1e-6, batch=1):
    """Persistent-kernel (semismooth-Newton augmented Lagrangian) solve of reference Solver.solve_fully_consistent (src/solver.py:123-210).
    stats: CUDA fp64, (batch, c*c+c+2) or flat for batch=1.  Returns (x (batch, c) CUDA, info numpy structured array)."""
    _require_cuda()
    lib = _lib.load()
    if reg_type not in _lib.REG_TYPES:
        raise ValueError(f"reg_type {reg_type!r} is not supported on this path")
    c = 10 * num_links + 2 * ndof
    slen = c * c + c + 2
    stats = stats.contiguous()
    if stats.dtype != torch.float64 or not stats.is_cuda or stats.numel() != batch * slen:
        raise ValueError(f"stats: expected CUDA float64 with {batch}*{slen} elements")
    phi0 = np.ascontiguousarray(np.asarray(phi_prior).astype(np.float64))
    sa = np.ascontiguousarray(np.array([e["semi_axes"] for e in ellipsoids], dtype=np.float64).reshape(-1))
    ce = np.ascontiguousarray(np.array([e["center"] for e in ellipsoids], dtype=np.float64).reshape(-1))
    if phi0.size != 10 * num_links or sa.size != 3 * num_links or ce.size != 3 * num_links:
        raise ValueError("phi_prior / bounding_ellipsoids do not match num_links")
    d = _lib.SdpDesc()
    d.num_links = num_links; d.ndof = ndof
    d.phi_prior = phi0.ctypes.data_as(C.POINTER(C.c_double))
    d.semi_axes = sa.ctypes.data_as(C.POINTER(C.c_double))
    d.centers = ce.ctypes.data_as(C.POINTER(C.c_double))
    d.total_mass = float(total_mass); d.lambda_reg = float(lambda_reg); d.reg_type = _lib.REG_TYPES[reg_type]
    d.epsilon = float(epsilon); d.tol = float(tol); d.max_iters = int(max_iters)
    x = torch.empty((batch, c), dtype=torch.float64, device=stats.device)
    info = torch.zeros(batch * _lib.SDP_INFO_DTYPE.itemsize, dtype=torch.uint8, device=stats.device)
    one = lib.sysid_sdp_workspace_bytes(num_links, ndof)
    plan_bytes = lib.sysid_sdp_plan_bytes(num_links)
    nbytes = plan_bytes + (one - plan_bytes) * batch
    ws = torch.empty(nbytes, dtype=torch.uint8, device=stats.device)
    _lib.check(lib.sysid_sdp_solve(C.byref(d), _ptr(stats), slen, batch, _ptr(x), _ptr(info), _ptr(ws), ws.numel(), _stream()))
    info_np = info.cpu().numpy().view(_lib.SDP_INFO_DTYPE)
    return x, info_np


class SdpPlan:
    """Device-resident plan of the LMI-constrained fit for one (prior, ellipsoids, total mass, lambda, reg_type): the host work
    of the reference's Solver.__init__ + problem build (src/solver.py:6-29,55-121) done ONCE; solve() then launches without
    host work or synchronisation.  Also owns the solver workspace and the warm-start record of the pre-solve."""

    def __init__(self, num_links, ndof, phi_prior, ellipsoids, total_mass, lambda_reg=1e-1, tol=1e-10, max_iters=0,
                 reg_type="constant_pullback", epsilon=1e-6, device=None):
        _require_cuda()
        self.lib = _lib.load()
        if reg_type not in _lib.REG_TYPES:
            raise ValueError(f"reg_type {reg_type!r} is not supported on this path")
        self.L, self.nd = int(num_links), int(ndof)
        self.c = 10 * self.L + 2 * self.nd
        self.slen = self.c * self.c + self.c + 2
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._phi0 = np.ascontiguousarray(np.asarray(phi_prior).astype(np.float64))
        self._sa = np.ascontiguousarray(np.array([e["semi_axes"] for e in ellipsoids], dtype=np.float64).reshape(-1))
        self._ce = np.ascontiguousarray(np.array([e["center"] for e in ellipsoids], dtype=np.float64).reshape(-1))
        if self._phi0.size != 10 * self.L or self._sa.size != 3 * self.L or self._ce.size != 3 * self.L:
            raise ValueError("phi_prior / bounding_ellipsoids do not match num_links")
        d = _lib.SdpDesc()
        d.num_links = self.L; d.ndof = self.nd
        d.phi_prior = self._phi0.ctypes.data_as(C.POINTER(C.c_double))
        d.semi_axes = self._sa.ctypes.data_as(C.POINTER(C.c_double))
        d.centers = self._ce.ctypes.data_as(C.POINTER(C.c_double))
        d.total_mass = float(total_mass); d.lambda_reg = float(lambda_reg); d.reg_type = _lib.REG_TYPES[reg_type]
        d.epsilon = float(epsilon); d.tol = float(tol); d.max_iters = int(max_iters)
        self.desc = d
        self.plan = torch.empty(self.lib.sysid_sdp_plan_bytes(self.L), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.sysid_sdp_plan_create(C.byref(d), _ptr(self.plan), self.plan.numel(), _stream()))
        self.wlen = int(self.lib.sysid_sdp_warm_len(self.L, self.nd))
        self.warm = torch.zeros(self.wlen, dtype=torch.float64, device=self.device)           # record of the pre-solve
        self._ws = {}
        self._pre = None

    def key(self):
        return (self.L, self.nd, self._phi0.tobytes(), self._sa.tobytes(), self._ce.tobytes(), self.desc.total_mass, self.desc.lambda_reg,
                self.desc.reg_type, self.desc.epsilon, self.desc.tol, self.desc.max_iters)

    def _workspace(self, batch, tag=""):
        k = (batch, tag)
        if k not in self._ws:
            self._ws[k] = torch.empty(self.lib.sysid_sdp_solve_workspace_bytes(self.L, self.nd, batch), dtype=torch.uint8, device=self.device)
        return self._ws[k]

    def presolve_struct(self, samples=0, refine_at=0, tol=None, first_tol=None):
        """tol: tolerance of the pre-solves (default: the plan's own).  They only produce a starting point, so a looser one shortens the
        chain hidden behind the stream without changing where the final solve converges."""
        if self._pre is None:
            self._pre_bufs = (self._workspace(1, "pre"), torch.empty(self.slen, dtype=torch.float64, device=self.device),
                              torch.empty(self.c, dtype=torch.float64, device=self.device),
                              torch.zeros(_lib.SDP_INFO_DTYPE.itemsize, dtype=torch.uint8, device=self.device),
                              torch.empty(self.slen, dtype=torch.float64, device=self.device))
            p = _lib.Presolve()
            self._pre_desc = _lib.SdpDesc.from_buffer_copy(self.desc)        # same problem; its own tolerance (the pointers are shared)
            p.desc = C.pointer(self._pre_desc); p.plan = self.plan.data_ptr()
            p.sdp_workspace = self._pre_bufs[0].data_ptr(); p.sdp_workspace_bytes = self._pre_bufs[0].numel()
            p.stats_snapshot = self._pre_bufs[1].data_ptr(); p.x_scratch = self._pre_bufs[2].data_ptr()
            p.info_scratch = self._pre_bufs[3].data_ptr(); p.warm_out = self.warm.data_ptr()
            p.stats_snapshot2 = self._pre_bufs[4].data_ptr()
            self._pre = p
        self._pre.samples = int(samples)
        self._pre.refine_at = int(refine_at)
        self._pre_desc.tol = float(tol) if tol else self.desc.tol
        self._pre.first_tol = float(first_tol) if first_tol else 0.0
        return self._pre

    def presolve_info(self):
        """Solver record of the last pre-solve (numpy structured scalar); synchronises."""
        return self._pre_bufs[3].cpu().numpy().view(_lib.SDP_INFO_DTYPE)[0] if self._pre is not None else None

    def solve(self, stats, batch=1, warm=None, x_out=None, info_out=None, sync_info=True):
        """stats: CUDA fp64 (batch, c*c+c+2) or flat.  warm: None, or a CUDA fp64 tensor of batch * warm_len doubles (e.g.
        self.warm after a pre-solve).  Returns (x (batch, c) CUDA, info) -- info is the numpy record array when sync_info, else
        the raw device buffer (no synchronisation)."""
        stats = stats.contiguous()
        if stats.dtype != torch.float64 or not stats.is_cuda or stats.numel() != batch * self.slen:
            raise ValueError(f"stats: expected CUDA float64 with {batch}*{self.slen} elements")
        x = x_out if x_out is not None else torch.empty((batch, self.c), dtype=torch.float64, device=stats.device)
        info = info_out if info_out is not None else torch.zeros(batch * _lib.SDP_INFO_DTYPE.itemsize, dtype=torch.uint8, device=stats.device)
        ws = self._workspace(batch)
        if warm is not None and (warm.dtype != torch.float64 or not warm.is_cuda or warm.numel() != batch * self.wlen):
            raise ValueError("warm: expected CUDA float64 with batch * warm_len elements")
        _lib.check(self.lib.sysid_sdp_solve_plan(C.byref(self.desc), _ptr(self.plan), _ptr(stats), self.slen, batch, _ptr(x), _ptr(info),
                                                 _ptr(ws), ws.numel(), _ptr(warm), None, _stream()))
        if sync_info:
            return x, info.cpu().numpy().view(_lib.SDP_INFO_DTYPE)
        return x, info
