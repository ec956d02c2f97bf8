"""ORACLE (test infrastructure, not product code) -- stage 3: the LMI-constrained fit of
src/solver.py restated in matrix form and solved by a log-barrier Newton method in numpy fp64.

PARITY UNPINNED against cvxpy+MOSEK (not installable / licence-bound; the reference states no
fallback solver, src/solver.py:203-210).  This is the STATED CPU reference solve: the problem below
is strictly convex (H > 0), so its optimum is unique and any solver that certifies the KKT
conditions (kkt_certificate) is within sqrt(2*gap/lambda_min(H)) of what MOSEK (rel-gap 1e-10)
converges to.

Restates (reference file:line):
  * Solver.__init__ sizes: nx = Y.shape[1], num_samples = Y.shape[0] (ROWS, quirk Q4)   src/solver.py:6-29
  * _construct_pseudo_inertia_matrix  J(phi)                                           src/solver.py:55-65
  * _construct_ellipsoid_matrix       Q (float32, +Q top-left, quirk Q2)               src/solver.py:67-75
  * _construct_com_constraint_matrix  C(phi)                                           src/solver.py:77-93
  * _pullback_metric                  M[a,b] = tr(P^-1 V_a P^-1 V_b)                   src/solver.py:95-121
  * solve_fully_consistent: objective, m>=0, J+1e-6 I >= 0, C+1e-6 I >= 0, tr(JQ)>=0,
    sum m == total_mass, b_v, b_c >= 0; ValueError when not optimal                    src/solver.py:123-210
Variable x = [phi (10 L, REFERENCE order m,hx,hy,hz,Ixx,Ixy,Ixz,Iyy,Iyz,Izz) ; b_v (d) ; b_c (d)].
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

EPS_LMI = 1e-6  # src/solver.py:145


def pseudo_inertia(phi):
    """src/solver.py:55-65."""
    m, hx, hy, hz, Ixx, Ixy, Ixz, Iyy, Iyz, Izz = [float(t) for t in phi]
    t = 0.5 * (Ixx + Iyy + Izz)
    return np.array([[t - Ixx, -Ixy, -Ixz, hx],
                     [-Ixy, t - Iyy, -Iyz, hy],
                     [-Ixz, -Iyz, t - Izz, hz],
                     [hx, hy, hz, m]], dtype=np.float64)


def com_matrix(phi, semi_axes, center):
    """src/solver.py:77-93."""
    m = float(phi[0]); h = np.asarray(phi[1:4], dtype=np.float64)
    C = np.zeros((4, 4))
    C[0, 0] = m
    C[0, 1:] = h - m * np.asarray(center, float)
    C[1:, 0] = C[0, 1:]
    C[1:, 1:] = m * np.diag(np.asarray(semi_axes, float)) ** 2
    return C


def ellipsoid_matrix(semi_axes, center):
    """src/solver.py:67-75: stored as float32, +Q in the top-left block (quirk Q2)."""
    semi_axes = np.asarray(semi_axes, dtype=np.float64); center = np.asarray(center, dtype=np.float64)
    Qf = np.zeros((4, 4), dtype=np.float32)
    Q = np.linalg.inv(np.diag(semi_axes) ** 2)
    Qf[:3, :3] = Q
    Qf[:3, 3] = Q @ center
    Qf[3, :3] = (Q @ center).T
    Qf[3, 3] = 1 - (center.T @ Q @ center)
    return Qf


def pullback_metric(phi0):
    """src/solver.py:95-121 (P = J(phi0) without epsilon; symmetrise; shift if an eigenvalue < 0)."""
    P_inv = np.linalg.inv(pseudo_inertia(phi0))
    V = [pseudo_inertia(np.eye(10)[a]) for a in range(10)]
    M = np.zeros((10, 10))
    for a in range(10):
        for b in range(10):
            M[a, b] = np.trace(P_inv @ V[a] @ P_inv @ V[b])
    M = (M + M.T) / 2
    ev = np.linalg.eigvals(M)
    if np.any(ev < 0):
        M = M + (-np.min(ev) + 1e-5) * np.eye(10)
    assert np.min(np.linalg.eigvals(M)) > 0, "Matrix is not positive definite."
    return M


@dataclass
class SdpProblem:
    """min 1/2 x^T H x - g^T x  s.t.  F_k(x) = F0_k + sum_a x[idx_k[a]] T_k[a] >= 0 (4x4, k < 2L),
    Ain x + bin >= 0, aeq^T x = beq."""
    H: np.ndarray
    g: np.ndarray
    const: float
    lmi_idx: np.ndarray     # (K,10) indices into x
    lmi_T: np.ndarray       # (K,10,4,4)
    lmi_F0: np.ndarray      # (K,4,4)
    Ain: np.ndarray         # (m_in, c)
    bin: np.ndarray
    aeq: np.ndarray
    beq: float
    num_links: int
    ndof: int

    @property
    def nx(self):
        return self.H.shape[0]


def build_problem(G, r, s, n, num_links, phi_prior, total_mass, bounding_ellipsoids, ndof=0,
                  lambda_reg=1e-1, reg_type="constant_pullback") -> SdpProblem:
    """G = A^T A, r = A^T b, s = b^T b with A = [Y_proj | B_v | B_c]; n = A.shape[0] (quirk Q4)."""
    L = num_links
    c = 10 * L + 2 * ndof
    G = np.asarray(G, dtype=np.float64); r = np.asarray(r, dtype=np.float64)
    assert G.shape == (c, c) and r.shape == (c,)
    phi0 = np.asarray(phi_prior).astype(np.float64)   # float32 values widened (quirk Q3)
    H = G / n
    g = r / n
    const = 0.5 * s / n
    Vs = np.array([pseudo_inertia(np.eye(10)[a]) for a in range(10)])
    lmi_idx, lmi_T, lmi_F0, Ain, bin_ = [], [], [], [], []
    for i in range(L):
        sl = slice(10 * i, 10 * i + 10)
        ell = bounding_ellipsoids[i]
        sa, ce = np.asarray(ell["semi_axes"], float), np.asarray(ell["center"], float)
        # m >= 0
        row = np.zeros(c); row[10 * i] = 1.0
        Ain.append(row); bin_.append(0.0)
        # J + eps I >= 0
        lmi_idx.append(np.arange(10 * i, 10 * i + 10)); lmi_T.append(Vs); lmi_F0.append(EPS_LMI * np.eye(4))
        # C + eps I >= 0
        Ws = np.array([com_matrix(np.eye(10)[a], sa, ce) for a in range(10)])
        lmi_idx.append(np.arange(10 * i, 10 * i + 10)); lmi_T.append(Ws); lmi_F0.append(EPS_LMI * np.eye(4))
        # tr(J Q) >= 0
        Q = ellipsoid_matrix(sa, ce).astype(np.float64)
        row = np.zeros(c); row[sl] = [np.trace(Vs[a] @ Q) for a in range(10)]
        Ain.append(row); bin_.append(0.0)
        if reg_type == "constant_pullback":
            M = pullback_metric(phi0[sl])
            H[sl, sl] += lambda_reg * M
            g[sl] += lambda_reg * (M @ phi0[sl])
            const += 0.5 * lambda_reg * phi0[sl] @ M @ phi0[sl]
        elif reg_type == "euclidean":
            H[sl, sl] += 2 * lambda_reg * np.eye(10)      # quad_form without the 1/2 (src/solver.py:175-177)
            g[sl] += 2 * lambda_reg * phi0[sl]
            const += lambda_reg * phi0[sl] @ phi0[sl]
        else:
            raise ValueError(f"reg_type {reg_type!r} is out of scope (entropic is marked non-converging upstream)")
    for k in range(2 * ndof):
        row = np.zeros(c); row[10 * L + k] = 1.0
        Ain.append(row); bin_.append(0.0)
    aeq = np.zeros(c); aeq[0:10 * L:10] = 1.0
    H = 0.5 * (H + H.T)
    return SdpProblem(H=H, g=g, const=float(const), lmi_idx=np.array(lmi_idx), lmi_T=np.array(lmi_T),
                      lmi_F0=np.array(lmi_F0), Ain=np.array(Ain), bin=np.array(bin_), aeq=aeq,
                      beq=float(total_mass), num_links=L, ndof=ndof)


def objective(prob: SdpProblem, x):
    return 0.5 * x @ prob.H @ x - prob.g @ x + prob.const


def lmi_values(prob: SdpProblem, x):
    return prob.lmi_F0 + np.einsum("ka,kaij->kij", x[prob.lmi_idx], prob.lmi_T)


# ------------------------------------------------------------------ barrier Newton
def _barrier_terms(prob, x, extra_t=None):
    """value, gradient, Hessian of  -sum logdet F_k - sum log(lin)  (optionally with + t I / + t)."""
    c = prob.nx
    F = lmi_values(prob, x)
    lin = prob.Ain @ x + prob.bin
    if extra_t is not None:
        F = F + extra_t * np.eye(4)
        lin = lin + extra_t
    if not (np.all(np.isfinite(F)) and np.all(np.isfinite(lin))) or np.any(lin <= 0):
        return None
    try:
        Ls = np.linalg.cholesky(F)
        Finv = np.linalg.inv(F)
    except np.linalg.LinAlgError:
        return None
    dg = np.diagonal(Ls, axis1=1, axis2=2)
    if not np.all(np.isfinite(dg)) or np.any(dg <= 0):
        return None
    val = -2.0 * np.sum(np.log(dg)) - np.sum(np.log(lin))
    nt = c + (1 if extra_t is not None else 0)
    grad = np.zeros(nt); Hs = np.zeros((nt, nt))
    # LMI part: d/dx_a = -tr(Finv T_a); d2 = tr(Finv T_a Finv T_b)
    FT = np.einsum("kij,kajl->kail", Finv, prob.lmi_T)          # Finv T_a
    gk = -np.einsum("kaii->ka", FT)
    Hk = np.einsum("kaij,kbji->kab", FT, FT)
    for k in range(F.shape[0]):
        idx = prob.lmi_idx[k]
        np.add.at(grad, idx, gk[k])
        Hs[np.ix_(idx, idx)] += Hk[k]
        if extra_t is not None:
            tF = np.trace(Finv[k])
            grad[c] += -tF
            Hs[c, c] += np.sum(Finv[k] * Finv[k].T)
            col = np.einsum("aij,ji->a", FT[k], Finv[k])        # tr(Finv T_a Finv)
            Hs[idx, c] += col; Hs[c, idx] += col
    il = 1.0 / lin
    grad[:c] += -prob.Ain.T @ il
    Hs[:c, :c] += prob.Ain.T @ (il[:, None] ** 2 * prob.Ain)
    if extra_t is not None:
        grad[c] += -np.sum(il)
        Hs[c, c] += np.sum(il ** 2)
        v = prob.Ain.T @ (il ** 2)
        Hs[:c, c] += v; Hs[c, :c] += v
    return val, grad, Hs, F, lin


def _newton_centering(fun, x, aeq, beq, max_newton=60, tol=1e-13):
    """Minimise fun(x) -> (val, grad, Hess) or None if infeasible, s.t. aeq^T x = beq (infeasible start
    OK).  Stops on the Newton decrement lambda^2/2 <= tol (barrier units, self-concordant).  Inside the
    quadratic-convergence region (lambda^2 < 0.05) full steps are taken without a function-value
    test, because at large barrier parameter the objective value no longer resolves the decrease."""
    n = x.size
    steps = 0
    for _ in range(max_newton):
        out = fun(x)
        if out is None:
            return x, steps, False
        val, grad, Hs = out
        K = np.zeros((n + 1, n + 1))
        K[:n, :n] = Hs; K[:n, n] = aeq; K[n, :n] = aeq
        rhs = np.concatenate([-grad, [beq - aeq @ x]])
        try:
            sol = np.linalg.solve(K, rhs)
        except np.linalg.LinAlgError:
            return x, steps, False
        if not np.all(np.isfinite(sol)):
            return x, steps, False
        dx = sol[:n]
        lam2 = float(dx @ Hs @ dx)
        feas_eq = abs(beq - aeq @ x) <= 1e-13 * max(1.0, abs(beq))
        if feas_eq and lam2 / 2 <= tol:
            break
        t = 1.0
        while True:
            xn = x + t * dx
            o2 = fun(xn)
            if o2 is not None and (not feas_eq or lam2 < 0.05 or o2[0] <= val + 0.25 * t * (grad @ dx)):
                break
            t *= 0.5
            if t < 1e-14:
                return x, steps, False
        x = xn
        steps += 1
    else:
        return x, steps, False
    return x, steps, True


def _scaling(prob: SdpProblem):
    """Block-Jacobi scaling x = T y: Cholesky of each 10x10 link block of H, 1/sqrt(diag) for friction
    (cond(H) 1.5e14 raw -> ~10 scaled)."""
    c = prob.nx
    T = np.zeros((c, c))
    for i in range(prob.num_links):
        sl = slice(10 * i, 10 * i + 10)
        Lc = np.linalg.cholesky(prob.H[sl, sl])
        T[sl, sl] = np.linalg.inv(Lc).T
    for k in range(10 * prob.num_links, c):
        T[k, k] = 1.0 / np.sqrt(prob.H[k, k]) if prob.H[k, k] > 0 else 1.0
    return T


def _scaled_problem(prob: SdpProblem, T):
    lmi_T = np.zeros_like(prob.lmi_T)
    for k in range(prob.lmi_T.shape[0]):
        idx = prob.lmi_idx[k]
        Tb = T[np.ix_(idx, idx)]                 # block structure: link blocks only mix within themselves
        lmi_T[k] = np.einsum("ab,aij->bij", Tb, prob.lmi_T[k])
    return SdpProblem(H=T.T @ prob.H @ T, g=T.T @ prob.g, const=prob.const, lmi_idx=prob.lmi_idx, lmi_T=lmi_T,
                      lmi_F0=prob.lmi_F0, Ain=prob.Ain @ T, bin=prob.bin, aeq=T.T @ prob.aeq, beq=prob.beq,
                      num_links=prob.num_links, ndof=prob.ndof)


def solve_barrier(prob: SdpProblem, x0=None, mu=8.0, t0=1.0, gap_tol=1e-12, verbose=False):
    """Returns (x, info).  Raises ValueError('The problem did not solve to optimality.') like
    src/solver.py:209-210 when no strictly feasible point exists."""
    c = prob.nx
    T = _scaling(prob)
    sp = _scaled_problem(prob, T)
    Tinv = np.linalg.inv(T)
    if x0 is None:
        x0 = np.zeros(c)
    y = Tinv @ np.asarray(x0, dtype=np.float64)
    # make friction start strictly positive
    n_newton = 0
    F = lmi_values(sp, y); lin = sp.Ain @ y + sp.bin
    min_slack = min(np.min(np.linalg.eigvalsh(F)), np.min(lin))
    eq_ok = abs(sp.aeq @ y - sp.beq) <= 1e-12 * max(1.0, abs(sp.beq))
    if min_slack <= 0 or not eq_ok:
        # phase I: min t  s.t. F_k + t I >= 0, lin + t >= 0, equality
        # first satisfy the equality by the minimum-norm correction, then pick t strictly feasible
        y = y + sp.aeq * (sp.beq - sp.aeq @ y) / (sp.aeq @ sp.aeq)
        F = lmi_values(sp, y); lin = sp.Ain @ y + sp.bin
        min_slack = min(np.min(np.linalg.eigvalsh(F)), np.min(lin))
        if min_slack <= 0:
            z = np.concatenate([y, [max(1.0, 2.0 * -min_slack) + 1e-3]])
            aeq1 = np.concatenate([sp.aeq, [0.0]])
            tt = 1.0
            found = False
            for _outer in range(60):
                def fun(zz, tt=tt):
                    out = _barrier_terms(sp, zz[:c], extra_t=zz[c])
                    if out is None:
                        return None
                    val, grad, Hs, _, _ = out
                    # tiny proximal term keeps the phase-I Hessian nonsingular in unbounded directions
                    val = val + tt * zz[c] + 0.5e-8 * (zz[:c] @ zz[:c])
                    grad = grad.copy(); grad[c] += tt; grad[:c] += 1e-8 * zz[:c]
                    Hs = Hs + 1e-8 * np.diag(np.concatenate([np.ones(c), [0.0]]))
                    return val, grad, Hs
                z, k, _ = _newton_centering(fun, z, aeq1, sp.beq, tol=1e-8)
                n_newton += k
                if z[c] < 0:
                    F = lmi_values(sp, z[:c]); lin = sp.Ain @ z[:c] + sp.bin
                    if min(np.min(np.linalg.eigvalsh(F)), np.min(lin)) > 0:
                        found = True
                        break
                tt *= mu
            if not found:
                raise ValueError("The problem did not solve to optimality.")
            y = z[:c]
    m_total = 4 * sp.lmi_T.shape[0] + sp.Ain.shape[0]
    tt = t0
    y_good, t_good = None, None
    while True:
        def fun(yy, tt=tt):
            out = _barrier_terms(sp, yy)
            if out is None:
                return None
            val, grad, Hs, _, _ = out
            return (tt * (0.5 * yy @ sp.H @ yy - sp.g @ yy) + val, tt * (sp.H @ yy - sp.g) + grad, tt * sp.H + Hs)
        y_new, k, ok = _newton_centering(fun, y, sp.aeq, sp.beq)
        n_newton += k
        if verbose:
            print(f"  barrier t={tt:.3e} newton={k} ok={ok} gap={m_total / tt:.3e}")
        if not ok:
            # fp64 ran out on the central path (active LMIs make F^-1 ~ t): keep the last centred point
            if y_good is None:
                raise ValueError("The problem did not solve to optimality.")
            y, tt = y_good, t_good
            break
        y = y_new
        y_good, t_good = y, tt
        fval = abs(0.5 * y @ sp.H @ y - sp.g @ y + sp.const)
        if m_total / tt < gap_tol * max(1.0, fval):
            break
        tt *= mu
    x = T @ y
    info = kkt_certificate(prob, x, tt)
    info["newton_steps"] = n_newton
    info["t_final"] = tt
    return x, info


def kkt_certificate(prob: SdpProblem, x, tt):
    """Solver-independent optimality certificate from the central-path duals Z_k = F_k^-1 / t,
    z = 1/(t lin): primal/dual feasibility, complementarity (= duality gap) and stationarity, plus
    the strong-convexity distance bound ||x - x*||_H <= sqrt(2 gap)."""
    F = lmi_values(prob, x)
    lin = prob.Ain @ x + prob.bin
    Z = np.linalg.inv(F) / tt
    z = 1.0 / (tt * lin)
    dual_lmi = np.zeros(prob.nx)
    for k in range(F.shape[0]):
        np.add.at(dual_lmi, prob.lmi_idx[k], np.einsum("aij,ji->a", prob.lmi_T[k], Z[k]))
    grad_f = prob.H @ x - prob.g
    resid0 = grad_f - dual_lmi - prob.Ain.T @ z
    nu = (prob.aeq @ resid0) / (prob.aeq @ prob.aeq)
    stat = resid0 - nu * prob.aeq
    # measure stationarity in block-Jacobi-scaled coordinates: raw H has cond ~1e14, so an unscaled
    # norm is dominated by the stiffest (lower-leg inertia) directions and says nothing about the rest
    Tsc = _scaling(prob)
    stat_s, grad_s = Tsc.T @ stat, Tsc.T @ grad_f
    gap = float(np.sum(np.einsum("kij,kji->k", Z, F)) + z @ lin)
    return {
        "primal_min_eig": float(np.min(np.linalg.eigvalsh(F))),
        "primal_min_lin": float(np.min(lin)),
        "eq_residual": float(abs(prob.aeq @ x - prob.beq)),
        "dual_min_eig": float(np.min(np.linalg.eigvalsh(Z))),
        "dual_min_lin": float(np.min(z)),
        "gap": gap,
        "stationarity_rel": float(np.linalg.norm(stat) / max(1e-300, np.linalg.norm(grad_f))),
        "stationarity_rel_scaled": float(np.linalg.norm(stat_s) / max(1e-300, np.linalg.norm(grad_s))),
        "dist_bound_H": float(np.sqrt(2 * max(gap, 0.0))),
        "objective": float(objective(prob, x)),
        "nu": float(nu),
    }


def solve_fully_consistent(A, b, num_links, phi_prior, total_mass, bounding_ellipsoids, ndof=0,
                           lambda_reg=1e-1, reg_type="constant_pullback"):
    """Reference-shaped entry (stacked A = [Y|B_v|B_c], b): returns (phi, b_v, b_c, info)."""
    G, r, s, n = A.T @ A, A.T @ b, float(b @ b), A.shape[0]
    prob = build_problem(G, r, s, n, num_links, phi_prior, total_mass, bounding_ellipsoids, ndof, lambda_reg, reg_type)
    x0 = np.concatenate([np.asarray(phi_prior, dtype=np.float64), np.ones(2 * ndof)])
    x, info = solve_barrier(prob, x0)
    L = num_links
    return x[:10 * L], x[10 * L:10 * L + ndof], x[10 * L + ndof:], info


# ------------------------------------------------------------------ second, independent solve: SSN augmented Lagrangian
# The barrier method above stalls near gap ~1e-8 in fp64 when an LMI is active (F^-1 ~ t).  polish_alm restarts from
# any point and drives the KKT residuals of the SAME problem to ~1e-10 with a semismooth-Newton augmented-Lagrangian
# iteration (numpy, dense).  Tests use barrier-vs-ALM agreement as the oracle's own accuracy estimate.
_SV_IJ = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (3, 2), (3, 3)]
_SQ2 = np.sqrt(2.0)


def _svec(F):
    return np.array([F[i, j] * (1.0 if i == j else _SQ2) for (i, j) in _SV_IJ])


def _smat(v):
    F = np.zeros((4, 4))
    for r, (i, j) in enumerate(_SV_IJ):
        F[i, j] = F[j, i] = v[r] / (1.0 if i == j else _SQ2)
    return F


def solve_alm(prob: SdpProblem, x0=None, tol=1e-10, max_newton=400, verbose=False):
    """Returns (x, info) with info = {'kkt', 'dual_residual', 'newton_steps', 'sigma', 'objective'}."""
    c = prob.nx
    T = _scaling(prob)
    H = T.T @ prob.H @ T; g = T.T @ prob.g; a = T.T @ prob.aeq; beq = prob.beq
    K = prob.lmi_T.shape[0]; mlin = prob.Ain.shape[0]; m = 10 * K + mlin
    A = np.zeros((m, c)); c0 = np.zeros(m)
    for k in range(K):
        idx = prob.lmi_idx[k]
        Mk = np.array([[prob.lmi_T[k][b_][i, j] * (1.0 if i == j else _SQ2) for b_ in range(10)] for (i, j) in _SV_IJ])
        A[10 * k:10 * k + 10][:, idx] = Mk @ T[np.ix_(idx, idx)]
        c0[10 * k:10 * k + 10] = _svec(prob.lmi_F0[k])
    A[10 * K:] = prob.Ain @ T; c0[10 * K:] = prob.bin
    for k in range(K):
        sl = slice(10 * k, 10 * k + 10)
        sc = 1.0 / np.sqrt((A[sl] ** 2).sum() / 10)
        A[sl] *= sc; c0[sl] *= sc
    for r in range(10 * K, m):
        sc = 1.0 / np.linalg.norm(A[r])
        A[r] *= sc; c0[r] *= sc

    def project(w):
        out = w.copy(); eig = []
        for k in range(K):
            lam, V = np.linalg.eigh(_smat(w[10 * k:10 * k + 10]))
            out[10 * k:10 * k + 10] = _svec((V * np.maximum(lam, 0)) @ V.T); eig.append((lam, V))
        out[10 * K:] = np.maximum(w[10 * K:], 0)
        return out, eig

    def hess_term(eig, w):
        Hh = np.zeros((c, c))
        for k in range(K):
            lam, V = eig[k]; lp = np.maximum(lam, 0)
            Om = np.zeros((4, 4))
            for i in range(4):
                for j in range(4):
                    if abs(lam[i] - lam[j]) > 1e-14 * max(1.0, abs(lam[i]), abs(lam[j])):
                        Om[i, j] = (lp[i] - lp[j]) / (lam[i] - lam[j])
                    else:
                        Om[i, j] = 1.0 if lam[i] > 0 else 0.0
            idx = prob.lmi_idx[k]
            Ht = [V.T @ _smat(A[10 * k:10 * k + 10, col]) @ V for col in idx]
            for ia, ca in enumerate(idx):
                for ib, cb in enumerate(idx):
                    Hh[ca, cb] += np.sum(Om * Ht[ia] * Ht[ib])
        act = (w[10 * K:] > 0).astype(float)
        Al = A[10 * K:]
        return Hh + Al.T @ (act[:, None] * Al)

    y = np.linalg.solve(T, np.asarray(x0, dtype=np.float64)) if x0 is not None else a * beq / (a @ a)
    y = y + a * (beq - a @ y) / (a @ a)
    lam = np.zeros(m); sigma = 1.0; n_newton = 0; kkt_prev = 1.0
    gnorm = np.linalg.norm(g); kkt = gn = np.inf
    for outer in range(200):
        tol_in = max(0.5 * tol * (1 + gnorm), 1e-2 * min(1.0, kkt_prev))
        for inner in range(40):
            w = lam - sigma * (A @ y + c0); pw, eig = project(w)
            grad = H @ y - g - A.T @ pw
            pg = grad - a * (a @ grad) / (a @ a)
            gn = np.linalg.norm(pg)
            if gn <= tol_in or n_newton >= max_newton:
                break
            Kin = np.linalg.inv(H + sigma * hess_term(eig, w))
            v1 = Kin @ grad; Ka = Kin @ a
            dy = -(v1 - Ka * (a @ v1) / (a @ Ka))
            val0 = pw @ pw / (2 * sigma); lin = (H @ y - g) @ dy; q2 = dy @ H @ dy; gd = grad @ dy
            t = 1.0; ok = False
            for _ in range(40):
                pwt, _e = project(lam - sigma * (A @ (y + t * dy) + c0))
                if t * lin + 0.5 * t * t * q2 + pwt @ pwt / (2 * sigma) <= val0 + 1e-4 * t * gd + 1e-14 * (abs(val0) + 1):
                    ok = True; break
                t *= 0.5
            n_newton += 1
            if not ok:
                break
            y = y + t * dy
        gy = A @ y + c0
        lam_new, _e = project(lam - sigma * gy)
        kkt = np.linalg.norm(lam_new - lam) / sigma
        lam = lam_new
        if verbose:
            print(f"  alm outer {outer} sigma {sigma:.0e} newton {n_newton} kkt {kkt:.2e} dual {gn:.2e}")
        if (kkt <= tol * (1 + np.linalg.norm(gy)) and gn <= tol * (1 + gnorm)) or n_newton >= max_newton:
            break
        if kkt > 0.25 * kkt_prev:
            sigma = min(sigma * 10, 1e6)
        kkt_prev = kkt
    x = T @ y
    return x, {"kkt": float(kkt), "dual_residual": float(gn), "newton_steps": n_newton, "sigma": sigma,
               "objective": float(objective(prob, x))}
