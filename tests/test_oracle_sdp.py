"""Stage-3 oracle: the matrix restatement of src/solver.py and its two independent solves (barrier, SSN-ALM)."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import sdp


def _problem(name, lam=0.1, reg="constant_pullback"):
    g = np.load(os.path.join(H.GOLDEN_DIR, f"{name}_N48.npz"))
    flat = H.flat_model(name)
    prob = sdp.build_problem(g["sdp_G"], g["sdp_r"], float(g["sdp_s"]), float(g["sdp_n"]), flat.nbodies, flat.phi_prior,
                             flat.robot_mass, flat.ellipsoids, flat.joints_dof, lambda_reg=lam, reg_type=reg)
    return flat, prob, g


def test_constraint_builders_match_reference_formulas():
    phi = np.array([2.0, 0.1, -0.2, 0.05, 0.3, 0.01, -0.02, 0.4, 0.03, 0.5])
    J = sdp.pseudo_inertia(phi)
    t = 0.5 * (0.3 + 0.4 + 0.5)
    assert np.allclose(J, [[t - 0.3, -0.01, 0.02, 0.1], [-0.01, t - 0.4, -0.03, -0.2], [0.02, -0.03, t - 0.5, 0.05], [0.1, -0.2, 0.05, 2.0]])
    sa, ce = np.array([0.1, 0.2, 0.3]), np.array([0.01, -0.02, 0.03])
    C = sdp.com_matrix(phi, sa, ce)
    assert np.allclose(C[0, 1:], phi[1:4] - 2.0 * ce) and np.allclose(np.diag(C)[1:], 2.0 * sa ** 2) and C[0, 0] == 2.0
    Q = sdp.ellipsoid_matrix(sa, ce)
    assert Q.dtype == np.float32 and Q[0, 0] > 0                      # quirk Q2: +Q, float32
    assert abs(Q[3, 3] - (1 - np.sum(ce ** 2 / sa ** 2))) < 1e-6


def test_pullback_metric_is_spd_and_scale_covariant():
    flat = H.flat_model("solo12")
    M = sdp.pullback_metric(flat.phi_prior[10:20].astype(float))
    assert np.abs(M - M.T).max() == 0 and np.linalg.eigvalsh(M).min() > 0
    # tr(P^-1 V P^-1 V) with V = J(phi0) itself is tr(I_4) = 4
    p = flat.phi_prior[10:20].astype(float)
    assert abs(p @ M @ p - 4.0) < 1e-6 * 4


@pytest.mark.parametrize("name", H.ROBOTS)
def test_barrier_and_alm_agree_and_certify(name):
    flat, prob, g = _problem(name)
    x0 = np.concatenate([flat.phi_prior.astype(float), np.ones(2 * flat.joints_dof)])
    xb, ib = sdp.solve_barrier(prob, x0)
    xa, ia = sdp.solve_alm(prob)
    assert H.rel(xb, xa) < 1e-6
    assert ia["kkt"] < 1e-8 and ia["dual_residual"] < 1e-8
    assert ib["eq_residual"] < 1e-10 and ib["primal_min_eig"] > -1e-12 and ib["primal_min_lin"] > -1e-12
    assert ib["stationarity_rel_scaled"] < 1e-5
    assert abs(ib["objective"] - ia["objective"]) <= 1e-7 * max(1.0, abs(ia["objective"]))
    # pinned against the committed fixture
    assert H.rel(xa, g["sdp_x_alm"]) < 1e-8
    # feasibility of the ALM point
    F = sdp.lmi_values(prob, xa)
    assert np.linalg.eigvalsh(F).min() > -1e-8 and (prob.Ain @ xa + prob.bin).min() > -1e-8
    assert abs(prob.aeq @ xa - prob.beq) < 1e-9


def test_active_lmi_case_and_euclidean_regulariser():
    flat, prob, _ = _problem("solo12", lam=1e-3)
    xa, ia = sdp.solve_alm(prob)
    F = sdp.lmi_values(prob, xa)
    assert np.linalg.eigvalsh(F).min() < 1e-7                          # at least one LMI is active
    flat, prob, _ = _problem("solo12", lam=1e-2, reg="euclidean")
    x0 = np.concatenate([flat.phi_prior.astype(float), np.ones(24)])
    xb, _ = sdp.solve_barrier(prob, x0)
    xa, _ = sdp.solve_alm(prob)
    assert H.rel(xb, xa) < 1e-5


def test_infeasible_problem_raises_like_reference():
    flat, prob, _ = _problem("solo12")
    prob.beq = -1.0                                                     # sum of non-negative masses cannot be negative
    with pytest.raises(ValueError, match="did not solve to optimality"):
        sdp.solve_barrier(prob, np.concatenate([flat.phi_prior.astype(float), np.ones(24)]))


def test_objective_matches_stacked_form():
    """1/2 ||A x - b||^2 / n + lambda sum 1/2 (phi-phi0)^T M (phi-phi0) evaluated from the stack equals the Gram form."""
    flat, data = H.small_log("solo12", 12, seed=21)
    _, _, A, b = H.oracle_blocks(flat, data)
    G, r, s, n = A.T @ A, A.T @ b, float(b @ b), A.shape[0]
    prob = sdp.build_problem(G, r, s, n, 13, flat.phi_prior, flat.robot_mass, flat.ellipsoids, 12)
    x = np.random.default_rng(0).normal(size=154) * 0.01 + np.concatenate([flat.phi_prior.astype(float), np.full(24, 0.02)])
    phi0 = flat.phi_prior.astype(float)
    direct = 0.5 * np.sum((A @ x - b) ** 2) / n
    for i in range(13):
        d = x[10 * i:10 * i + 10] - phi0[10 * i:10 * i + 10]
        direct += 0.1 * 0.5 * d @ sdp.pullback_metric(phi0[10 * i:10 * i + 10]) @ d
    assert abs(sdp.objective(prob, x) - direct) <= 1e-9 * abs(direct)
