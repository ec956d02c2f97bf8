"""identify(): the end-to-end call BASELINE.json's metric names.  Five host arrays in, phi (and friction) out:

    upload (pinned -> HBM)  ->  fused regressor+projector+Gram kernel  ->  [all-reduce over ranks]  ->  on-device LMI solve

It computes what the reference demo's main() computes between read_data and the printers
(reference demo/solo_identification.py:67-88) without ever forming the stacked regressor.
"""
from __future__ import annotations

import numpy as np
import torch

from . import distributed as D
from .ops import to_device


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


PRESOLVE_MIN_SAMPLES = 16384      # shortest first chunk worth a pre-solve (below that the warm start is too noisy to pay for itself)
PRESOLVE_MIN_LOG = 600_000        # the pre-solve (a cold solve, ~9 ms) only pays when the rest of the stream hides it: the statistics of
                                  # 600 k samples take ~9-11 ms; on shorter (or sharded) logs the final solve would just wait for it
                                  # (measured with two ranks on the 1 M-sample log, threshold lowered to 380 k: 19.1 ms instead of 17.6 --
                                  # the first chunk of one shard is 6 % of the log and its fit a poorer start: 37 final Newton steps, not 23)


# A SECOND pre-solve stage: the statistics of the first PRESOLVE_REFINE_FRACTION of the shard, solved behind the first stage and
# warm-started from it; its record starts the final solve (13 Newton steps instead of 23 on the 1 M-sample G1 log: 18.6 instead of
# 20.1 ms).  Only where the stream hides both stages -- ~1 + 9 + 4 ms of solves against 14 ms of statistics per 1 M samples; at a
# fraction of 0.5 and more the chain overruns the stream and the final solve waits for it (tools/presolve_refine_study.py).
PRESOLVE_REFINE_MIN_LOG = 900_000
PRESOLVE_REFINE_FRACTION = 0.45
PRESOLVE_FIRST_TOL = 1e-7          # tolerance of the first pre-solve when a second stage follows (the second stage and the final solve keep
                                   # the caller's): the chain of pre-solves, not the stream, bounds identify() on the 1 M-sample log, and the
                                   # last decades of the first stage buy nothing -- 17.6 instead of 18.7 ms; 1e-6: the same; 1e-4: 22-26 ms
PRESOLVE_TOL = None                # tolerance of the pre-solves (None: the final solve's; looser ones were measured: 1e-6 costs the final
                                   # solve 7 more Newton steps, 1e-5 twenty)


def _plan_for(sysid, L, nd, lambda_reg, tol, max_iters, reg_type):
    """Device plan of the LMI problem, built once per (prior, ellipsoids, mass, lambda, ...) and cached on the sysid object."""
    from .ops import SdpPlan
    from .solver import NEWTON_STEPS_PER_IPM_ITER
    prior = np.asarray(sysid.get_phi_prior()).astype(np.float64)
    ell = sysid.get_bounding_ellipsoids()
    key = (L, nd, float(lambda_reg), float(tol), int(max_iters), reg_type, float(sysid.get_robot_mass()), prior.tobytes(),
           np.array([e["semi_axes"] for e in ell], dtype=np.float64).tobytes(), np.array([e["center"] for e in ell], dtype=np.float64).tobytes(),
           torch.cuda.current_device())
    cache = sysid.__dict__.setdefault("_sdp_plans", {})
    if key not in cache:
        if len(cache) >= 8:
            cache.clear()
        cache[key] = SdpPlan(L, nd, prior, ell, sysid.get_robot_mass(), lambda_reg=lambda_reg, tol=tol,
                             max_iters=int(max_iters) * NEWTON_STEPS_PER_IPM_ITER, reg_type=reg_type)
    return cache[key]


def identify(sysid, q, dq, ddq, tau, cnt, lambda_reg=1e-1, tol=1e-10, max_iters=1000, reg_type="constant_pullback",
             friction=True, return_info=False, sharded=False, weights=None, presolve=True, chunk=131072):
    """sysid: SystemIdentification.  Arrays: (channels x N) numpy or torch (host or device).
    With torch.distributed initialised and sharded=False every rank passes the FULL log and takes its own
    contiguous shard; with sharded=True each rank passes only its shard.  Returns phi (10 L,) [, b_v, b_c, info].

    Host arrays (what the reference's read_data returns) are streamed: chunked upload overlapped with the fused kernel, and
    -- presolve=True, logs of at least PRESOLVE_MIN_LOG samples per rank ("force": any length) -- the LMI fit of the first
    chunk's statistics solved behind the rest of the stream; its point and multipliers warm-start the final solve (same unique
    optimum, about half the Newton steps).

    Deviations from the reference that are REPORTED, never silent (a RuntimeWarning each, counts in info):
      * samples with a non-finite input are skipped (the reference would propagate NaN into the whole stack);
      * a contact Jacobian that loses row rank has the dependent row dropped at 1e-13 of the largest row norm squared
        (numpy's pinv cuts singular values at 1e-15 sigma_max; the two differ only where cond(J_c) > 3e6).
    A log in which EVERY sample is skipped raises ValueError."""
    import warnings
    from . import _lib
    dm = sysid.device_model
    rank, ws = D.world()
    N = q.shape[1]
    if ws > 1 and not sharded:
        lo, hi = D.shard_bounds(N, rank, ws)
        q, dq, ddq, tau, cnt = (a[:, lo:hi] for a in (q, dq, ddq, tau, cnt))
        if weights is not None:
            weights = weights[lo:hi]
    fixed = not sysid._floating_base
    if fixed:                                   # floating_base=False: emulated on the free-flyer kernels (urdf.py::flatten)
        q, dq, ddq, tau, cnt = sysid._to_floating(q, dq, ddq, tau, cnt)
    arrays = (q, dq, ddq, tau, cnt)
    n_loc = arrays[0].shape[1]
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    L, nd = sysid.get_num_links(), (sysid.joints_dof if friction else 0)
    c = 10 * L + 2 * nd                         # columns of the problem that is solved
    cdev = dm.ncols(friction)                   # columns of the statistics the kernels produce (fixed base: + the pinned root body)
    slen = cdev * cdev + cdev + 2
    err = None
    plan = None
    if rank == 0:
        try:
            plan = _plan_for(sysid, L, nd, lambda_reg, tol, max_iters, reg_type)
        except Exception as e:                                           # noqa: BLE001 -- re-raised below, on every rank
            err = e
    # [stats (c*c + c + 2) | rank-loss count, skipped count] in ONE buffer so that a single all-reduce merges both
    buf = torch.zeros(slen + 2, dtype=torch.float64, device=device)
    stats = buf[:slen]
    counts = torch.zeros(2, dtype=torch.int64, device=device)
    warm = None
    if _host_streamable(arrays, weights):
        # host float64 / float32 arrays: chunked upload overlapped with the kernel inside the library
        n0 = min(int(chunk), max(PRESOLVE_MIN_SAMPLES, n_loc // 8))
        use_pre = (presolve is True and n_loc >= PRESOLVE_MIN_LOG or presolve == "force") and plan is not None and n_loc >= 2 * n0
        refine = int(PRESOLVE_REFINE_FRACTION * n_loc) if (use_pre and n_loc >= PRESOLVE_REFINE_MIN_LOG) else 0
        dm.gram_accumulate_host(*arrays, friction=friction, weights=weights, stats=stats, info=counts, chunk=chunk,
                                presolve=plan if use_pre else None, presolve_samples=n0, presolve_refine_at=refine, presolve_tol=PRESOLVE_TOL,
                                presolve_first_tol=PRESOLVE_FIRST_TOL if refine else None)
        warm = plan.warm if use_pre else None
    else:
        dev = [a if (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64) else to_device(a) for a in arrays]
        dev = [a if a.stride(1) == 1 else a.contiguous() for a in dev]
        if len({a.stride(0) for a in dev}) != 1:
            dev = [a.contiguous() for a in dev]
        if weights is not None and not (isinstance(weights, torch.Tensor) and weights.is_cuda):
            weights = torch.as_tensor(np.asarray(weights, dtype=np.float64)).cuda()
        dm.gram_accumulate(*dev, friction=friction, weights=weights, stats=stats, info=counts)
    buf[slen:] = counts.to(torch.float64)
    D.allreduce_stats(buf)
    # rank 0 solves; its outcome (status and solution, or the fact that it raised) reaches every rank through ONE broadcast, so
    # that no rank is left waiting in a collective when the solve fails on the host.  No host synchronisation before the launch.
    out = torch.zeros(c + 5, dtype=torch.float64, device=device)      # [x (c) | status | failed | rank-loss | skipped | rows]
    info_dev = None
    if rank == 0:
        try:
            if err is not None:
                raise err
            x, info_dev = plan.solve(sysid._stats_to_fixed(stats, friction) if fixed else stats, warm=warm, sync_info=False)
            out[:c] = x[0]
            out[c] = info_dev[:4].view(torch.int32)[0].to(torch.float64)
        except Exception as e:                                           # noqa: BLE001 -- re-raised below, on every rank
            err = e
            out[c + 1] = 1.0
        out[c + 2:c + 4] = buf[slen:]
        out[c + 4] = buf[slen - 1]
    if ws > 1:
        D.broadcast_solution(out)
    outh = out.cpu().numpy()                                              # the one synchronisation of the call
    status, failed, n_rankloss, n_skipped, n_rows = int(outh[c]), outh[c + 1] != 0.0, int(outh[c + 2]), int(outh[c + 3]), outh[c + 4]
    if failed:
        if err is not None:
            raise err
        raise RuntimeError("identify(): the LMI solve failed on rank 0 (see that rank's exception)")
    if not n_rows > 0.0:
        raise ValueError("identify(): no usable sample (every sample of the log has a non-finite input or zero weight)")
    if n_skipped:
        warnings.warn(f"identify(): {n_skipped} sample(s) with a non-finite input were skipped (the reference would propagate NaN)", RuntimeWarning, stacklevel=2)
    if n_rankloss and not fixed:                # (the anchors of a fixed base are rank deficient by construction: 9 rows, rank 6)
        warnings.warn(f"identify(): the contact Jacobian lost row rank in {n_rankloss} sample(s); dependent rows were dropped "
                      "(pinv semantics, cutoff 1e-13 of the largest squared row norm)", RuntimeWarning, stacklevel=2)
    if status not in (0, 1):   # 1 = optimal_inaccurate, accepted like the reference accepts cp.OPTIMAL_INACCURATE
        print("The problem did not solve to optimality. Status:", status)
        raise ValueError("The problem did not solve to optimality.")
    xh = outh[:c]
    phi = xh[:10 * L].copy()
    if not return_info:
        return phi
    info_d = {}
    if info_dev is not None:
        rec = info_dev.cpu().numpy().view(_lib.SDP_INFO_DTYPE)[0]
        info_d = {k: rec[k].item() for k in rec.dtype.names}
        if warm is not None:
            pre = plan.presolve_info()
            info_d["presolve_iterations"] = int(pre["iterations"]); info_d["presolve_status"] = int(pre["status"])
    info_d.update(status=status, rank_deficient_samples=n_rankloss, skipped_samples=n_skipped)
    return phi, xh[10 * L:10 * L + nd].copy(), xh[10 * L + nd:].copy(), info_d
