"""Bootstrap resampling of the identification (BASELINE.json configs[4]: B resamples of one log, many Gram + LMI solves).

The reference has no bootstrap; this is the batched use of its path that the drop-in makes affordable: every resample is
the reference's identification (demo/solo_identification.py:67-88) of a log in which sample (or block) i appears w_bi
times.  Because the fused kernel returns ADDITIVE sufficient statistics, a resample never re-reads the log more than
once:

  block == 1   sample-level bootstrap: one weighted sysid_gram_accumulate launch per resample (multinomial weights),
  block  > 1   moving-block bootstrap for time series: ONE launch of the fused kernel in segmented mode (sysid_gram_blocks: one
               statistics vector per block of `block` consecutive samples), then stats_b = sum_k w_bk stats_k for all resamples
               in ONE launch on the fp64 tensor pipe (sysid_combine_stats),

followed by ONE batched sysid_sdp_solve_plan launch (one thread block per resample) -- three launches for B fits.  Under torch.distributed the resamples
shard by problem over the ranks (no data-path collective) and are gathered on every rank.
"""
from __future__ import annotations

import numpy as np
import torch

from . import distributed as D
from .identify import _plan_for
from .ops import combine_stats, to_device


def bootstrap_weights(n_units, B, seed):
    """(B, n_units) multinomial multiplicities, rows summing to n_units (numpy default_rng(seed): reproducible on the host)."""
    rng = np.random.default_rng(seed)
    return rng.multinomial(n_units, np.full(n_units, 1.0 / n_units), size=B).astype(np.float64)


def bootstrap_identify(sysid, q, dq, ddq, tau, cnt, B=1024, block=1, seed=1005, lambda_reg=1e-1, tol=1e-10, max_iters=1000,
                       reg_type="constant_pullback", friction=True, return_stats=False):
    """Returns x (B, c) = [phi | b_v | b_c] per resample (numpy), the solver info array, and optionally the (B, slen) stats."""
    from .solver import NEWTON_STEPS_PER_IPM_ITER
    dm = sysid.device_model
    rank, ws = D.world()
    dev = [a if (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64) else to_device(a) for a in (q, dq, ddq, tau, cnt)]
    N = dev[0].shape[1]
    L, nd = sysid.get_num_links(), (sysid.joints_dof if friction else 0)
    c = 10 * L + 2 * nd
    slen = dm.stats_len(friction)
    block = int(max(1, block))
    K = (N + block - 1) // block
    W = bootstrap_weights(K, B, seed)                       # identical on every rank
    lo, hi = D.shard_bounds(B, rank, ws)                    # this rank's resamples
    Wl = torch.from_numpy(W[lo:hi]).to(dev[0].device)
    nb = hi - lo
    if block == 1:
        # sample-level bootstrap: B weighted passes (the work is inherently B Grams; prefer block > 1 for time series)
        stats = torch.zeros((nb, slen), dtype=torch.float64, device=dev[0].device)
        for b in range(nb):
            dm.gram_accumulate(*dev, friction=friction, weights=Wl[b].contiguous(), stats=stats[b])
    else:
        # moving-block bootstrap: ONE launch for the K per-block statistics, ONE for all resamples (stats_b = sum_k w_bk stats_k)
        dev = [a if a.stride(1) == 1 else a.contiguous() for a in dev]
        if len({a.stride(0) for a in dev}) != 1:
            dev = [a.contiguous() for a in dev]
        per_block = dm.gram_blocks(*dev, block, friction=friction)
        stats = combine_stats(Wl.contiguous(), per_block) if nb > 0 else torch.zeros((0, slen), dtype=torch.float64, device=dev[0].device)
    x = torch.empty((nb, c), dtype=torch.float64, device=dev[0].device)
    info, err = None, None
    try:
        if nb > 0:
            plan = _plan_for(sysid, L, nd, lambda_reg, tol, max_iters, reg_type)       # host work once, then ONE batched launch
            x, info = plan.solve(stats, batch=nb)
    except Exception as e:                                  # noqa: BLE001 -- every rank must still reach the collective below
        err = e
    if ws > 1:
        import torch.distributed as dist
        full = torch.zeros((B + 1, c), dtype=torch.float64, device=dev[0].device)
        if err is None:
            full[lo:hi] = x
        else:
            full[B, 0] = 1.0                                # failure flag, summed over ranks
        dist.all_reduce(full, op=dist.ReduceOp.SUM)         # disjoint row ranges: a gather
        if err is None and float(full[B, 0].item()) != 0.0:
            err = RuntimeError("bootstrap_identify(): the batched LMI solve failed on another rank")
        x = full[:B]
    if err is not None:
        raise err
    out = (x.cpu().numpy(), info)
    return out + (stats,) if return_stats else out
