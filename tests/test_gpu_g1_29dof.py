"""Unitree G1 with all 29 actuated joints (reference files/g1_description/g1_29dof.urdf: 30 bodies, nv = 35, c = 358): outside the
fused kernel's compile-time envelope, served by the large-model path (csrc/bigmodel.cuh).  Every stage against the oracle."""
import numpy as np
import pytest

import helpers as H

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _setup(N, seed=7, robot="g1_29dof"):
    from system_identification_b200.ops import DeviceModel, to_device
    flat = H.flat_model(robot)
    q, dq, ddq, cnt = H.synth.make_trajectory(flat, N, seed)
    tau = H.synth.synth_tau(flat, N, 3, scale=5.0)
    data = (q, dq, ddq, tau, cnt)
    return flat, data, DeviceModel(flat), tuple(to_device(a) for a in data)


def test_g1_29dof_regressor_projector_blocks_vs_oracle():
    flat, data, dm, dev = _setup(40)
    assert (dm.nv, dm.nb, dm.nd, dm.ncols(True)) == (35, 30, 29, 358)
    Y = dm.regressor_batch(*dev[:3]).cpu().numpy()
    A, b, P = dm.projected_batch(*dev, want_P=True)
    A, b, P = A.cpu().numpy(), b.cpu().numpy(), P.cpu().numpy()
    Yo, Po, Ao, bo = H.oracle_blocks(flat, tuple(a[:, :10] for a in data))
    for i in range(10):
        assert np.abs(Y[i] - Yo[i]).max() <= 1e-10 * np.abs(Yo[i]).max()
        assert np.abs(P[i] - Po[i]).max() <= 1e-11
    assert np.array_equal(Y[:10] == 0, Yo == 0)                              # same structural zeros
    assert np.abs(A[:10].reshape(350, 358) - Ao).max() <= 1e-10 * np.abs(Ao).max()
    assert np.abs(b[:10].reshape(-1) - bo).max() <= 1e-10 * np.abs(bo).max()
    A2, _ = dm.projected_batch(*dev, friction=False)
    assert A2.shape[-1] == 300 and np.abs(A2.cpu().numpy() - A[..., :300]).max() <= 1e-13 * np.abs(A).max()


@pytest.mark.parametrize("robot,c", [("g1_29dof", 358), ("g1_29dof_lock_waist", 334)])
def test_g1_29dof_gram_and_rmse_vs_oracle_twin(robot, c):
    """Both large trees the reference ships (the lock-waist URDF merges two waist joints: 28 bodies, nv = 33): the SYRK's structural
    row masks are derived from the tree, so a second tree is the check that they are derived, not fitted."""
    from oracle.cbuild import COracle
    N = 20000                                                                 # three chunks (8 192 samples) of the large-model path, ragged tail
    flat, data, dm, dev = _setup(N, seed=8, robot=robot)
    nv, npar = dm.nv, 10 * dm.nb
    co = COracle(H.oracle_tree(flat), flat.ee_names)
    so, _ = co.gram(*data)
    st = dm.gram_accumulate(*dev)
    G, r, s, n = H.split_stats(st.cpu().numpy(), c)
    Go, ro, s_o, n_o = H.split_stats(so, c)
    assert H.rel(G, Go) <= 1e-12 and H.rel(r, ro) <= 1e-12 and abs(s - s_o) <= 1e-12 * s_o and n == n_o == nv * N
    assert np.array_equal(G, G.T) and torch.equal(dm.gram_accumulate(*dev), st)         # symmetric, bit-reproducible
    acc = dm.gram_accumulate(*(a[:, :1234] for a in dev))
    dm.gram_accumulate(*(a[:, 1234:] for a in dev), stats=acc)
    assert H.rel(acc.cpu().numpy(), so) <= 1e-12                                       # additive over shards
    G3 = H.split_stats(dm.gram_accumulate(*dev, friction=False).cpu().numpy(), npar)[0]
    assert H.rel(G3, Go[:npar, :npar]) <= 1e-12
    phi = flat.phi_prior.astype(np.float64)
    M = 600
    out = dm.predict_rmse(*(a[:, :M] for a in dev), torch.from_numpy(phi)).cpu().numpy()
    tot_o, pj_o = co.tau_rmse(*(a[:, :M] for a in data), phi)
    assert abs(out[0] - tot_o) <= 1e-9 * tot_o and np.abs(out[1:] - pj_o).max() <= 1e-9 * pj_o.max()
    with pytest.raises(Exception):
        dm.gram_accumulate(*dev, weights=torch.ones(N, dtype=torch.float64, device="cuda"))   # not on the large-model path


def test_g1_29dof_identify_vs_oracle():
    """Stage 3 at c = 358 (30 links, 29 joints: 60 LMIs, 718 constraint rows): the large-problem instantiation of the LMI solver
    against the oracle's solve of the oracle's own statistics, through identify() from host arrays."""
    import sys
    sys.path.insert(0, H.ROOT)
    from oracle import sdp as osdp
    from oracle.cbuild import COracle
    from src.sys_identification import SystemIdentification
    N, c = 4000, 358
    flat, data, dm, dev = _setup(N, seed=9)
    si = SystemIdentification.from_flat_model(flat)
    phi, bv, bc, info = si.identify(*data, return_info=True)
    assert info["status"] in (0, 1) and phi.shape == (300,) and bv.shape == (29,) and bc.shape == (29,)
    co = COracle(H.oracle_tree(flat), flat.ee_names)
    so, _ = co.gram(*data)
    G, r, s, n = H.split_stats(so, c)
    prob = osdp.build_problem(G, r, float(s), n, 30, flat.phi_prior, flat.robot_mass, flat.ellipsoids, 29)
    xo, _ = osdp.solve_alm(prob)
    x = np.concatenate([phi, bv, bc])
    assert H.rel(x, xo) <= 1e-4
    for i in range(30):
        assert H.rel(phi[10 * i:10 * i + 10], xo[10 * i:10 * i + 10]) <= 1e-4
    assert abs(phi[0::10].sum() - flat.robot_mass) <= 1e-8 * flat.robot_mass and bv.min() >= -1e-9 and bc.min() >= -1e-9
