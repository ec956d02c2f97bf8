"""floating_base=False (reference src/sys_identification.py:15-18,29-37: pin.buildModelFromUrdf(path), S = I, nq = nv = number of
joints): emulated on the free-flyer kernels with the base pinned by three always-in-stance anchor contacts on the root body
(urdf.py::flatten).  CPU: the host-side model; GPU: every stage against the oracle's fixed-base tree."""
import os

import numpy as np
import pytest
import yaml

import helpers as H

URDF_CANDIDATES = (os.path.join(H.ROOT, "baseline", "_ref", "files", "solo_description", "solo12.urdf"),
                   os.path.join(H.REFERENCE_FILES, "solo_description", "solo12.urdf"))


def _urdf():
    for p in URDF_CANDIDATES:
        if os.path.exists(p):
            return p
    pytest.skip("solo12.urdf is neither staged under baseline/_ref nor present under /root/reference")


def _fixed_sysid(tmp_path):
    from oracle import urdf_tree as ut
    from system_identification_b200.sys_identification import SystemIdentification
    urdf = _urdf()
    with open(os.path.join(os.path.dirname(urdf), "solo12_config.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["robot"]["link_names"] = cfg["robot"]["link_names"][1:]            # the base link is welded to the universe: 12 moving bodies
    cfg["robot"]["end_effectors_frame_names"] = []
    prior = ut.phi_prior(urdf, cfg["robot"]["link_names"])
    cfg["robot"]["mass"] = float(sum(prior[10 * i] for i in range(12)))
    path = str(tmp_path / "solo12_fixed.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return SystemIdentification(urdf, path, floating_base=False), ut.build_tree(urdf, floating_base=False), prior, cfg["robot"]["mass"]


def _log(N, seed=4):
    rng = np.random.default_rng(seed)
    t = np.arange(N) / 500.0
    ph = rng.uniform(0, 6.28, (12, 1)); f = rng.uniform(0.3, 2.0, (12, 1))
    q = 0.6 * np.sin(2 * np.pi * f * t + ph); dq = 0.6 * 2 * np.pi * f * np.cos(2 * np.pi * f * t + ph)
    ddq = -0.6 * (2 * np.pi * f) ** 2 * np.sin(2 * np.pi * f * t + ph)
    return q, dq, ddq, np.zeros((0, N))


def test_fixed_base_host_model(tmp_path):
    si, tree, prior, mass = _fixed_sysid(tmp_path)
    assert si.nq == si.nv == si.joints_dof == 12 and si._base_dof == 0 and np.array_equal(si._S, np.eye(12))
    assert si.get_num_links() == 12 and si._nb_ee == 0 and tree.nv == 12
    assert np.array_equal(si.get_phi_prior(), prior.astype(np.float32)) and si.get_robot_mass() == mass
    qF, dqF, ddqF, tauF, cntF = si._to_floating(np.ones((12, 3)), np.ones((12, 3)), np.ones((12, 3)), np.ones((12, 3)), np.zeros((0, 3)))
    assert qF.shape == (19, 3) and np.array_equal(qF[:7, 0], [0, 0, 0, 0, 0, 0, 1]) and np.array_equal(dqF[:6], np.zeros((6, 3)))
    assert cntF.shape == (3, 3) and np.all(cntF == 1.0)


@pytest.mark.gpu
def test_fixed_base_emulation_vs_oracle(tmp_path):
    torch = pytest.importorskip("torch")
    from oracle import dynamics as dy, sdp as osdp
    from system_identification_b200.solver import Solver
    si, tree, prior, mass = _fixed_sysid(tmp_path)
    N = 240
    q, dq, ddq, cnt = _log(N)
    rng = np.random.default_rng(9)
    # torques from a perturbed ground truth (P = I on a fixed base without contacts) + friction + noise
    phi_true = np.concatenate([(tree.dyn_params[i] if hasattr(tree, "dyn_params") else np.zeros(10)) for i in range(1, 13)]) * (1 + 0.1 * rng.standard_normal(120))
    Y = np.array([dy.joint_torque_regressor(tree, q[:, i], dq[:, i], ddq[:, i]) for i in range(N)])
    tau = (Y @ phi_true).T + 0.01 * dq + 0.02 * np.sign(dq) + 0.01 * rng.standard_normal((12, N))
    # per-sample API
    for i in (0, 7, N - 1):
        y, t = si.get_proj_regressor_torque(q[:, i], dq[:, i], ddq[:, i], tau[:, i], cnt[:, i])
        bv, bc = si.get_proj_friction_regressors(q[:, i], dq[:, i], ddq[:, i], cnt[:, i])
        yo, to = dy.proj_regressor_torque(tree, q[:, i], dq[:, i], ddq[:, i], tau[:, i], cnt[:, i], [], floating_base=False)
        bvo, bco = dy.proj_friction_regressors(tree, q[:, i], dq[:, i], ddq[:, i], cnt[:, i], [], floating_base=False)
        assert y.shape == (12, 120) and np.abs(y - yo).max() <= 1e-10 * np.abs(yo).max() and np.abs(t - to).max() <= 1e-10 * np.abs(to).max()
        assert np.abs(bv - bvo).max() <= 1e-12 and np.abs(bc - bco).max() <= 1e-12
    # statistics
    A, b = dy.stacked_system(tree, q, dq, ddq, tau, cnt, [], floating_base=False)
    G, r, s, n = H.split_stats(si.gram(q, dq, ddq, tau, cnt).cpu().numpy(), 144)
    assert H.rel(G, A.T @ A) <= 1e-11 and H.rel(r, A.T @ b) <= 1e-11 and abs(s - b @ b) <= 1e-11 * (b @ b) and n == A.shape[0] == 12 * N
    # LMI fit: fused path, reference-shaped path, oracle
    ell = si.get_bounding_ellipsoids()
    prob = osdp.build_problem(A.T @ A, A.T @ b, float(b @ b), A.shape[0], 12, si.get_phi_prior(), mass, ell, 12)
    xo, _ = osdp.solve_alm(prob)
    phi, bv, bc, info = si.identify(q, dq, ddq, tau, cnt, return_info=True)
    assert info["status"] in (0, 1) and H.rel(np.concatenate([phi, bv, bc]), xo) <= 1e-4
    sol = Solver(A[:, :120], b, 12, si.get_phi_prior(), mass, ell, B_v=A[:, 120:132], B_c=A[:, 132:144])
    assert H.rel(sol.solve_fully_consistent(), xo[:120]) <= 1e-4
    # evaluation pass with the reference's [6:] slice (joints 6..11 on a fixed base)
    tot, pj = si.tau_prediction_rmse(q, dq, ddq, tau, cnt, phi)
    tot_o, pj_o = dy.tau_prediction_rmse(tree, q, dq, ddq, tau, cnt, phi, [], floating_base=False)
    assert pj.shape == (6,) and abs(tot - tot_o) <= 1e-9 * tot_o and np.abs(pj - pj_o).max() <= 1e-9 * pj_o.max()
