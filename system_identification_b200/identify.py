"""identify(): the end-to-end call BASELINE.json's metric names.  Five host arrays in, phi (and friction) out:

    upload (pinned -> HBM)  ->  fused regressor+projector+Gram kernel  ->  [all-reduce over ranks]  ->  on-device LMI solve

It computes what the reference demo's main() computes between read_data and the printers
(reference demo/solo_identification.py:67-88) without ever forming the stacked regressor.
"""
from __future__ import annotations

import numpy as np
import torch

from . import distributed as D
from .ops import to_device, sdp_solve


def _host_streamable(arrays, weights):
    """True when all five arrays are host float64 / float32 (what the reference's read_data returns: float32 q and
    contact, float64 dq / ddq / tau -- demo/solo_identification.py:10-33), 2-D with unit inner stride."""
    for a in arrays:
        if isinstance(a, np.ndarray):
            if a.dtype not in (np.float64, np.float32) or a.ndim != 2 or a.strides[1] != a.itemsize:
                return False
        elif isinstance(a, torch.Tensor):
            if a.is_cuda or a.dtype not in (torch.float64, torch.float32) or a.dim() != 2 or a.stride(1) != 1:
                return False
        else:
            return False
    if weights is not None:
        w = weights
        if isinstance(w, np.ndarray):
            return w.dtype == np.float64 and w.ndim == 1 and w.flags.c_contiguous
        return isinstance(w, torch.Tensor) and (not w.is_cuda) and w.dtype == torch.float64 and w.is_contiguous()
    return True


def identify(sysid, q, dq, ddq, tau, cnt, lambda_reg=1e-1, tol=1e-10, max_iters=1000, reg_type="constant_pullback",
             friction=True, return_info=False, sharded=False, weights=None):
    """sysid: SystemIdentification.  Arrays: (channels x N) numpy or torch (host or device).
    With torch.distributed initialised and sharded=False every rank passes the FULL log and takes its own
    contiguous shard; with sharded=True each rank passes only its shard.  Returns phi (10 L,) [, b_v, b_c, info]."""
    from .solver import NEWTON_STEPS_PER_IPM_ITER
    dm = sysid.device_model
    rank, ws = D.world()
    N = q.shape[1]
    if ws > 1 and not sharded:
        lo, hi = D.shard_bounds(N, rank, ws)
        q, dq, ddq, tau, cnt = (a[:, lo:hi] for a in (q, dq, ddq, tau, cnt))
        if weights is not None:
            weights = weights[lo:hi]
    arrays = (q, dq, ddq, tau, cnt)
    if _host_streamable(arrays, weights):
        # host float64 arrays: chunked upload overlapped with the kernel inside the library
        stats = dm.gram_accumulate_host(*arrays, friction=friction, weights=weights)
    else:
        dev = [a if (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64) else to_device(a) for a in arrays]
        dev = [a if a.stride(1) == 1 else a.contiguous() for a in dev]
        if len({a.stride(0) for a in dev}) != 1:
            dev = [a.contiguous() for a in dev]
        if weights is not None and not (isinstance(weights, torch.Tensor) and weights.is_cuda):
            weights = torch.as_tensor(np.asarray(weights, dtype=np.float64)).cuda()
        stats = dm.gram_accumulate(*dev, friction=friction, weights=weights)
    D.allreduce_stats(stats)
    L, nd = sysid.get_num_links(), (sysid.joints_dof if friction else 0)
    c = 10 * L + 2 * nd
    if rank == 0:
        x, info = sdp_solve(stats, L, nd, sysid.get_phi_prior(), sysid.get_bounding_ellipsoids(), sysid.get_robot_mass(),
                            lambda_reg=lambda_reg, tol=tol, max_iters=int(max_iters) * NEWTON_STEPS_PER_IPM_ITER, reg_type=reg_type)
        status = int(info[0]["status"])
        x = x[0]
    else:
        x = torch.empty(c, dtype=torch.float64, device=stats.device)
        info, status = None, 0
    if ws > 1:
        st = torch.tensor([status], dtype=torch.int32, device=stats.device)
        D.broadcast_solution(st)
        status = int(st.item())
        D.broadcast_solution(x)
    if status not in (0, 1):   # 1 = optimal_inaccurate, accepted like the reference accepts cp.OPTIMAL_INACCURATE
        print("The problem did not solve to optimality. Status:", status)
        raise ValueError("The problem did not solve to optimality.")
    xh = x.cpu().numpy()
    phi = xh[:10 * L].copy()
    if not return_info:
        return phi
    info_d = None if info is None else {k: info[0][k].item() for k in info.dtype.names}
    return phi, xh[10 * L:10 * L + nd].copy(), xh[10 * L + nd:].copy(), info_d
