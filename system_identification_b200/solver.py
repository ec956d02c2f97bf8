"""Host-side mirror of the reference's Solver class (reference src/solver.py:5-210): same constructor and
methods, so the demos' `Solver(Y_proj, tau_proj, num_links, phi_prior, total_mass, bounding_ellipsoids,
B_v=..., B_c=...).solve_fully_consistent()` runs unmodified against `src.solver`.

No cvxpy, no MOSEK: the stacked system is reduced to its Gram statistics on the GPU
(sysid_gram_from_stack) and the LMI-constrained fit runs as one persistent semismooth-Newton augmented-Lagrangian kernel (sysid_sdp_solve).
`Solver.from_stats` skips the stack entirely (used by the fused identify() path).
"""
from __future__ import annotations

import numpy as np

NEWTON_STEPS_PER_IPM_ITER = 1   # the reference's max_iters caps MOSEK interior-point iterations; here it caps Newton steps


class _Value:
    """Stand-in for the cvxpy Variables the reference keeps in _phi/_b_v/_b_c (only `.value` is ever read)."""

    def __init__(self, value=None):
        self.value = value


class Solver():
    def __init__(self, regressor, tau_vec, num_links, phi_prior, total_mass, bounding_ellipsoids, B_v=None, B_c=None):
        self._Y = regressor
        self._tau = tau_vec
        self._nx = self._Y.shape[1]
        self._num_samples = self._Y.shape[0]           # ROWS of the stack (reference quirk Q4)
        self._num_links = num_links
        self._num_inertial_params = self._Y.shape[1] // self._num_links
        self._phi_prior = phi_prior
        self.total_mass = total_mass
        self._bounding_ellipsoids = bounding_ellipsoids
        self._phi = _Value(np.asarray(phi_prior))
        self._identify_fric = (B_v is not None) and (B_c is not None)
        self.ndof = 0
        if self._identify_fric:
            self._B_v = B_v
            self._B_c = B_c
            self.ndof = B_v.shape[1]
            self._b_v = _Value()
            self._b_c = _Value()
        self._stats = None
        self._problem = None
        self.info = None

    @classmethod
    def from_stats(cls, stats, num_links, phi_prior, total_mass, bounding_ellipsoids, ndof=0):
        """Solver on precomputed device statistics [G | r | s | n] (sysid_gram_accumulate output)."""
        self = cls.__new__(cls)
        self._Y = None; self._tau = None
        self._num_links = num_links
        self._num_inertial_params = 10
        self._nx = 10 * num_links
        self._phi_prior = phi_prior
        self.total_mass = total_mass
        self._bounding_ellipsoids = bounding_ellipsoids
        self._phi = _Value(np.asarray(phi_prior))
        self.ndof = ndof
        self._identify_fric = ndof > 0
        if self._identify_fric:
            self._b_v = _Value(); self._b_c = _Value()
        self._stats = stats
        self._num_samples = None
        self._problem = None
        self.info = None
        return self

    # -------------- Unconstrained Solver -------------- #
    def solve_llsq_svd(self):
        """Minimum-norm least squares of the full stack (reference src/solver.py:32-39: thin SVD, pinv of diag(Sigma),
        V Sigma^+ U^T tau).  The rows x c stack is reduced on the device to its c x c triangular factor R and z = Q^T tau
        (sysid_tsqr: blocked Householder, no Gram, so the 1e-15 cutoff of pinv means what it means in the reference);
        R has the singular values and right singular vectors of the stack, and the answer is V Sigma^+ U_R^T z."""
        import torch
        from .ops import _require_cuda, llsq_svd_from_triangle, tsqr
        _require_cuda()
        if self._Y is None:
            raise ValueError("solve_llsq_svd needs the stacked regressor (Solver built from_stats has none)")
        Y = torch.as_tensor(np.ascontiguousarray(np.asarray(self._Y, dtype=np.float64)), device="cuda")
        tau = torch.as_tensor(np.ascontiguousarray(np.asarray(self._tau, dtype=np.float64).reshape(-1)), device="cuda")
        x, self._singular_values = llsq_svd_from_triangle(tsqr(Y, tau))
        return x.cpu().numpy()

    # ------------ Constrained Solver (LMI) ------------ #
    def _device_stats(self):
        import torch
        from .ops import gram_from_stack, _require_cuda
        _require_cuda()
        if self._stats is not None:
            return self._stats
        blocks = [np.asarray(self._Y, dtype=np.float64)]
        if self._identify_fric:
            blocks += [np.asarray(self._B_v, dtype=np.float64), np.asarray(self._B_c, dtype=np.float64)]
        A = np.ascontiguousarray(np.hstack(blocks)) if len(blocks) > 1 else np.ascontiguousarray(blocks[0])
        b = np.ascontiguousarray(np.asarray(self._tau, dtype=np.float64).reshape(-1))
        if A.shape[0] != b.shape[0]:
            raise ValueError("regressor rows and tau_vec length differ")
        self._stats = gram_from_stack(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda())
        return self._stats

    def solve_fully_consistent(self, lambda_reg=1e-1, tol=1e-10, max_iters=1000, reg_type="constant_pullback"):
        """Constrained least squares with per-link LMIs (reference src/solver.py:123-210).  Returns phi (10 L,)."""
        from .ops import sdp_solve
        if reg_type == "entropic":
            raise ValueError("reg_type 'entropic' is marked non-converging in the reference (src/solver.py:164-172) and is not supported")
        stats = self._device_stats()
        x, info = sdp_solve(stats, self._num_links, self.ndof, self._phi_prior, self._bounding_ellipsoids, self.total_mass,
                            lambda_reg=lambda_reg, tol=tol, max_iters=int(max_iters) * NEWTON_STEPS_PER_IPM_ITER,
                            reg_type=reg_type)
        self.info = {k: info[0][k].item() for k in info.dtype.names}
        # 0 = optimal, 1 = optimal_inaccurate (accepted, as the reference accepts cp.OPTIMAL_INACCURATE, src/solver.py:206)
        self._problem = _Value({0: "optimal", 1: "optimal_inaccurate"}.get(self.info["status"], "not_optimal"))
        if self.info["status"] not in (0, 1):
            print("The problem did not solve to optimality. Status:", self._problem.value, self.info)
            raise ValueError("The problem did not solve to optimality.")
        x = x[0].cpu().numpy()
        p = self._num_links * 10
        self._phi.value = x[:p].copy()
        if self._identify_fric:
            self._b_v.value = x[p:p + self.ndof].copy()
            self._b_c.value = x[p + self.ndof:p + 2 * self.ndof].copy()
        return self._phi.value
