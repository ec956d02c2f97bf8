"""Parity against the reference's REAL dependencies, whenever their fixtures exist.

tests/golden/make_pinocchio_golden.py writes tests/golden/pinocchio_<robot>.npz on any box where pinocchio (and, for
stage 3, cvxpy + a conic solver) is installed.  These tests compare the oracle (CPU) and the CUDA path (-m gpu) with
those files and are SKIPPED while the files are absent -- which is the state of this repository as long as the build
image has no pinocchio ("parity unpinned")."""
import os

import numpy as np
import pytest

import helpers as H

TOL_Y = 1e-10       # north_star: regressor within 1e-10 relative of pinocchio's
TOL_PHI = 1e-4      # north_star: parameters within 1e-4 relative of the reference cvxpy solution


def _fixture(name):
    path = os.path.join(H.GOLDEN_DIR, f"pinocchio_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} not generated yet (needs a box with pinocchio: tests/golden/make_pinocchio_golden.py)")
    return np.load(path, allow_pickle=False)


def test_generator_script_is_committed_and_importable():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_pinocchio_golden", os.path.join(H.GOLDEN_DIR, "make_pinocchio_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert set(mod.ROBOTS) == set(H.ROBOTS)
    flat, data = mod.seeded_log("solo12")
    g = np.load(os.path.join(H.GOLDEN_DIR, "solo12_N48.npz"))
    assert np.array_equal(data[0], g["q"]) and np.array_equal(data[3], g["tau"])       # the same seeded log as the oracle fixtures


@pytest.mark.parametrize("name", H.ROBOTS)
def test_oracle_vs_pinocchio_fixture(name):
    from oracle import dynamics as dy
    g = _fixture(name)
    flat = H.flat_model(name)
    t = H.oracle_tree(flat)
    # model flattening: joint order, parents, placements, foot frames
    assert list(g["joint_names"]) == list(flat.joint_names)
    assert np.array_equal(g["parents"][1:], flat.parent[1:])
    assert np.abs(g["place_R"] - flat.place_R).max() <= 1e-12 and np.abs(g["place_p"] - flat.place_p).max() <= 1e-12
    assert np.array_equal(g["ee_parent"], flat.ee_joint) and np.abs(g["ee_offset"] - flat.ee_offset).max() <= 1e-12
    # quirk Q1: pinocchio's parameter order [m, mc, Ixx, Ixy, Iyy, Ixz, Iyz, Izz]
    assert np.allclose(g["q1_dyn_params_of_known_inertia"][0], 1.5) and np.allclose(g["q1_dyn_params_of_known_inertia"][1:4], 1.5 * np.array([0.1, -0.2, 0.3]))
    # inertia about the body-frame ORIGIN in the order xx, xy, yy, xz, yz, zz: I_C + m (|c|^2 I - c c^T)
    c = np.array([0.1, -0.2, 0.3]); Io = np.array([[1.0, 0.2, 0.3], [0.2, 2.0, 0.5], [0.3, 0.5, 3.0]]) + 1.5 * ((c @ c) * np.eye(3) - np.outer(c, c))
    assert np.allclose(g["q1_dyn_params_of_known_inertia"][4:], [Io[0, 0], Io[0, 1], Io[1, 1], Io[0, 2], Io[1, 2], Io[2, 2]], atol=1e-12)
    q, dq, ddq, cnt = g["q"], g["dq"], g["ddq"], g["cnt"]
    for i in range(q.shape[1]):
        Y = dy.joint_torque_regressor(t, q[:, i], dq[:, i], ddq[:, i])
        assert np.abs(Y - g["Y"][i]).max() <= TOL_Y * np.abs(g["Y"][i]).max()
        P = dy.null_space_projector(t, q[:, i], cnt[:, i], flat.ee_names)
        assert np.abs(P - g["P"][i]).max() <= 1e-10
    if "phi_identified" in g.files:
        from oracle import sdp as osdp
        A = np.hstack([g["Y_proj"], g["B_v"], g["B_c"]]); b = g["tau_proj"]
        prob = osdp.build_problem(A.T @ A, A.T @ b, float(b @ b), A.shape[0], 13, g["phi_prior"], float(g["robot_mass"]),
                                  [{"semi_axes": s, "center": c} for s, c in zip(g["ellipsoid_semi_axes"], g["ellipsoid_centers"])], 12)
        x, _ = osdp.solve_alm(prob)
        assert H.rel(x[:130], g["phi_identified"]) <= TOL_PHI


@pytest.mark.gpu
@pytest.mark.parametrize("name", H.ROBOTS)
def test_cuda_path_vs_pinocchio_fixture(name):
    torch = pytest.importorskip("torch")
    from system_identification_b200.ops import DeviceModel, to_device, sdp_solve
    g = _fixture(name)
    flat = H.flat_model(name)
    dm = DeviceModel(flat)
    dev = tuple(to_device(g[k]) for k in ("q", "dq", "ddq", "tau", "cnt"))
    Y = dm.regressor_batch(*dev[:3]).cpu().numpy()
    for i in range(Y.shape[0]):
        assert np.abs(Y[i] - g["Y"][i]).max() <= TOL_Y * np.abs(g["Y"][i]).max()
    _, _, P = dm.projected_batch(*dev, want_P=True)
    assert np.abs(P.cpu().numpy() - g["P"]).max() <= 1e-10
    if "phi_identified" in g.files:
        stats = dm.gram_accumulate(*dev)
        ell = [{"semi_axes": s, "center": c} for s, c in zip(g["ellipsoid_semi_axes"], g["ellipsoid_centers"])]
        x, info = sdp_solve(stats, 13, 12, g["phi_prior"], ell, float(g["robot_mass"]))
        assert int(info[0]["status"]) in (0, 1) and H.rel(x[0, :130].cpu().numpy(), g["phi_identified"]) <= TOL_PHI
