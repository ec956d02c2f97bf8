"""Pins the stage-1/2 oracle WITHOUT pinocchio: the regressor identity Y pi == RNEA(pi), energy-based checks of the
forward pass, finite differences for the contact Jacobian, projector identities, and the committed golden vectors."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import dynamics as dy


def _rand_state(t, rng, unit_quat=True):
    q = rng.normal(size=t.nq) * 0.6
    if unit_quat:
        q[3:7] /= np.linalg.norm(q[3:7])
    return q, rng.normal(size=t.nv), rng.normal(size=t.nv)


@pytest.mark.parametrize("name", H.ROBOTS)
def test_regressor_equals_rnea(name):
    flat = H.flat_model(name); t = H.oracle_tree(flat)
    rng = np.random.default_rng(0)
    for _ in range(5):
        q, v, a = _rand_state(t, rng)
        pi = rng.normal(size=10 * t.nbodies)
        Y = dy.joint_torque_regressor(t, q, v, a)
        tau = dy.rnea(t, q, v, a, pi)
        assert np.abs(Y @ pi - tau).max() <= 1e-13 * np.abs(tau).max()
    assert Y.shape == (t.nv, 10 * t.nbodies)


@pytest.mark.parametrize("name", H.ROBOTS)
def test_regressor_structural_sparsity(name):
    flat = H.flat_model(name); t = H.oracle_tree(flat)
    q, v, a = _rand_state(t, np.random.default_rng(1))
    Y = dy.joint_torque_regressor(t, q, v, a)
    expect = {"solo12": 756, "spot": 756, "g1_12dof": 936}[name]      # SURVEY section 6 (scratch figures, reproduced)
    assert np.count_nonzero(Y) == expect
    # rows touched by body i's 10 columns: the 6 base rows plus one per revolute ancestor (itself included)
    depth = flat.depth()
    for i in range(1, t.njoints):
        rows = np.nonzero(np.abs(Y[:, 10 * (i - 1):10 * i]).sum(axis=1))[0]
        assert len(rows) == 6 + (depth[i] - 1)


@pytest.mark.parametrize("name", H.ROBOTS)
def test_gravity_torque_is_potential_gradient(name):
    """Independent check of the forward pass: at v = a = 0 the joint rows of RNEA are dU/dq, with the potential
    U = -sum_i m_i g . c_i evaluated through world_placements (a code path the regressor does not use)."""
    flat = H.flat_model(name); t = H.oracle_tree(flat)
    rng = np.random.default_rng(2)
    q, _, _ = _rand_state(t, rng)

    def potential(qq):
        oR, op = dy.world_placements(t, qq)
        U = 0.0
        for i in range(1, t.njoints):
            m, h = t.dyn_params[i, 0], t.dyn_params[i, 1:4]
            U -= t.gravity @ (m * op[i] + oR[i] @ h)
        return U

    tau = dy.rnea(t, q, np.zeros(t.nv), np.zeros(t.nv))
    h = 1e-6
    for k in range(t.nv - 6):
        qp, qm = q.copy(), q.copy()
        qp[7 + k] += h; qm[7 + k] -= h
        fd = (potential(qp) - potential(qm)) / (2 * h)
        assert abs(fd - tau[6 + k]) <= 1e-6 * max(1.0, abs(tau[6 + k]))


@pytest.mark.parametrize("name", H.ROBOTS)
def test_mass_matrix_symmetric_positive(name):
    flat = H.flat_model(name); t = H.oracle_tree(flat)
    q, _, _ = _rand_state(t, np.random.default_rng(3))
    g = dy.rnea(t, q, np.zeros(t.nv), np.zeros(t.nv))
    M = np.array([dy.rnea(t, q, np.zeros(t.nv), e) - g for e in np.eye(t.nv)]).T
    assert np.abs(M - M.T).max() <= 1e-12 * np.abs(M).max()
    assert np.linalg.eigvalsh(0.5 * (M + M.T)).min() > 0
    assert abs(M[0, 0] - t.dyn_params[:, 0].sum()) <= 1e-12 * M[0, 0]      # total mass on the base translation


@pytest.mark.parametrize("name", H.ROBOTS)
def test_contact_jacobian_matches_finite_differences(name):
    flat = H.flat_model(name); t = H.oracle_tree(flat)
    q, _, _ = _rand_state(t, np.random.default_rng(4))
    foot = flat.ee_names[0]
    J = dy.frame_jacobian_lwa_linear(t, q, foot)
    h = 1e-6
    Jfd = np.zeros_like(J)
    R = dy.quat_to_matrix(*q[3:7])
    for k in range(t.nv - 6):
        qp, qm = q.copy(), q.copy(); qp[7 + k] += h; qm[7 + k] -= h
        Jfd[:, 6 + k] = (dy.frame_position(t, qp, foot) - dy.frame_position(t, qm, foot)) / (2 * h)
    for k in range(3):   # base translation in the LOCAL frame
        qp, qm = q.copy(), q.copy(); qp[:3] += h * R[:, k]; qm[:3] -= h * R[:, k]
        Jfd[:, k] = (dy.frame_position(t, qp, foot) - dy.frame_position(t, qm, foot)) / (2 * h)

    def qmul(a, b):
        x1, y1, z1, w1 = a; x2, y2, z2, w2 = b
        return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                         w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    for k in range(3):   # base rotation about local axes
        w = np.zeros(3); w[k] = h
        dqq = np.concatenate([w / 2, [1.0]]); dqq /= np.linalg.norm(dqq)
        dqm = dqq.copy(); dqm[:3] *= -1
        qp, qm = q.copy(), q.copy(); qp[3:7] = qmul(q[3:7], dqq); qm[3:7] = qmul(q[3:7], dqm)
        Jfd[:, 3 + k] = (dy.frame_position(t, qp, foot) - dy.frame_position(t, qm, foot)) / (2 * h)
    assert np.abs(J - Jfd).max() <= 1e-8


@pytest.mark.parametrize("name", H.ROBOTS)
def test_projector_identities_and_quirks(name):
    flat = H.flat_model(name); t = H.oracle_tree(flat)
    rng = np.random.default_rng(5)
    q, v, a = _rand_state(t, rng)
    n_ee = len(flat.ee_names)
    cnt = np.ones(n_ee)
    P = dy.null_space_projector(t, q, cnt, flat.ee_names)
    assert np.abs(P - P.T).max() <= 1e-13 and np.abs(P @ P - P).max() <= 1e-13
    assert abs(np.trace(P) - (t.nv - 3 * n_ee)) <= 1e-11
    # quirk Q5: state 2 counts as stance, and the over-allocated zero rows do not change P
    cnt2 = cnt.copy(); cnt2[0] = 2.0
    assert np.abs(dy.null_space_projector(t, q, cnt2, flat.ee_names) - P).max() <= 1e-13
    assert dy.contact_jacobian(t, q, cnt2, flat.ee_names).shape[0] == 3 * (n_ee + 1)
    # flight: P = I
    assert np.array_equal(dy.null_space_projector(t, q, np.zeros(n_ee), flat.ee_names), np.eye(t.nv))
    # (P A)^T (P A) == A^T P A
    Y = dy.joint_torque_regressor(t, q, v, a)
    assert np.abs((P @ Y).T @ (P @ Y) - Y.T @ P @ Y).max() <= 1e-10 * np.abs(Y.T @ Y).max()


def test_friction_blocks_sign_of_zero():
    flat = H.flat_model("solo12"); t = H.oracle_tree(flat)
    q, v, a = _rand_state(t, np.random.default_rng(6))
    v[6] = 0.0
    Bv, Bc = dy.proj_friction_regressors(t, q, v, a, np.array([1.0, 0, 0, 1.0]), flat.ee_names)
    assert np.all(Bc[:, 0] == 0.0) and np.all(Bv[:, 0] == 0.0)
    assert Bv.shape == (t.nv, 12) and Bc.shape == (t.nv, 12)


@pytest.mark.parametrize("name", H.ROBOTS)
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(H.GOLDEN_DIR, f"{name}_N48.npz"))
    flat = H.flat_model(name)
    data = (g["q"], g["dq"], g["ddq"], g["tau"], g["cnt"])
    sub = tuple(a[:, :8] for a in data)
    Y, P, A, b = H.oracle_blocks(flat, sub)
    assert np.abs(Y[:6] - g["Y"]).max() <= 1e-13 * np.abs(g["Y"]).max()
    assert np.abs(P - g["P"][:8]).max() <= 1e-13
    assert np.abs(A[:flat.nv * 4] - g["A_first"]).max() <= 1e-12 * np.abs(g["A_first"]).max()
    assert np.abs(b - g["b"][:flat.nv * 8]).max() <= 1e-12 * np.abs(g["b"]).max()


@pytest.mark.parametrize("name", ["solo12", "g1_12dof"])
def test_c_twin_matches_numpy_oracle(name):
    from oracle.cbuild import COracle
    flat, data = H.small_log(name, 24, seed=11)
    t = H.oracle_tree(flat)
    co = COracle(t, flat.ee_names)
    A, b, Y, P = co.blocks(*data)
    Yo, Po, Ao, bo = H.oracle_blocks(flat, data)
    assert np.abs(Y - Yo).max() <= 1e-13 * np.abs(Yo).max()
    assert np.abs(P - Po).max() <= 1e-12
    assert np.abs(A.reshape(-1, A.shape[-1]) - Ao).max() <= 1e-12 * np.abs(Ao).max()
    stats, used = co.gram(*data, nthreads=2)
    G, r, s, n = dy.gram_from_stack(Ao, bo)
    Gc, rc, sc, nc = H.split_stats(stats, Ao.shape[1])
    assert H.rel(Gc, G) <= 1e-13 and H.rel(rc, r) <= 1e-13 and abs(sc - s) <= 1e-12 * s and nc == n


def test_rmse_formulas_quirk_q7():
    flat, data = H.small_log("solo12", 6, seed=12)
    t = H.oracle_tree(flat)
    phi = flat.phi_prior.astype(float)
    tot, pj = dy.tau_prediction_rmse(t, *data, phi, flat.ee_names)
    _, _, A, b = H.oracle_blocks(flat, data)
    err = (A[:, :130] @ phi - b).reshape(6, t.nv)[:, 6:]
    assert abs(tot - np.mean(np.sum(err ** 2, axis=1))) <= 1e-12 * tot          # mean squared norm, no root
    assert np.abs(pj - np.sqrt(np.mean(err ** 2, axis=0))).max() <= 1e-12 * pj.max()


@pytest.mark.parametrize("name", H.ROBOTS)
def test_c_twin_rmse_matches_numpy_oracle(name):
    """oracle_tau_rmse (the C twin the full-size GPU tests compare against) == dy.tau_prediction_rmse, the numpy
    restatement of reference src/sys_identification.py:421-437."""
    from oracle.cbuild import COracle
    flat, data = H.small_log(name, 40, seed=13)
    t = H.oracle_tree(flat)
    co = COracle(t, flat.ee_names)
    rng = np.random.default_rng(5)
    for phi in (flat.phi_prior.astype(float), flat.phi_prior.astype(float) * (1 + 0.1 * rng.standard_normal(130))):
        tot, pj = dy.tau_prediction_rmse(t, *data, phi, flat.ee_names)
        tot_c, pj_c = co.tau_rmse(*data, phi, nthreads=2)
        assert abs(tot_c - tot) <= 1e-12 * tot and np.abs(pj_c - pj).max() <= 1e-12 * pj.max()
