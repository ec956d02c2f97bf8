"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes -> libsysid_b200.so), against the oracle on
the same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full sizes -- through
size-independent properties.  Tolerances: regressor 1e-10 relative (north_star), Gram 1e-12 relative Frobenius
(SURVEY section 8d), identified parameters 1e-4 relative, torque-prediction RMSE 0.5 %."""
import os

import numpy as np
import pytest

import helpers as H

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL_Y = 1e-10          # max-abs entry error / max-abs entry, per sample (north_star)
TOL_GRAM = 1e-12       # relative Frobenius
TOL_PHI = 1e-4         # global l2 and per-link-block l2, relative
TOL_RMSE = 1e-9         # the evaluation pass is a deterministic fp64 quantity (north_star's 0.5 % gate is far looser)


def _dev(flat):
    from system_identification_b200.ops import DeviceModel
    return DeviceModel(flat)


def _up(data):
    from system_identification_b200.ops import to_device
    return tuple(to_device(a) for a in data)


def _loaded_native():
    with open("/proc/self/maps") as f:
        return "libsysid_b200.so" in f.read()


@pytest.mark.parametrize("name", H.ROBOTS)
def test_regressor_vs_oracle_and_golden(name):
    flat, data = H.small_log(name, 48)
    g = np.load(os.path.join(H.GOLDEN_DIR, f"{name}_N48.npz"))
    assert np.array_equal(g["q"], data[0])                         # same seeded inputs as the fixture
    dm = _dev(flat)
    q, dq, ddq, tau, cnt = _up((g["q"], g["dq"], g["ddq"], g["tau"], g["cnt"]))
    Y = dm.regressor_batch(q, dq, ddq).cpu().numpy()
    assert _loaded_native()
    Yo, _, _, _ = H.oracle_blocks(flat, tuple(a[:, :12] for a in (g["q"], g["dq"], g["ddq"], g["tau"], g["cnt"])))
    for i in range(12):
        assert np.abs(Y[i] - Yo[i]).max() <= TOL_Y * np.abs(Yo[i]).max()
    assert np.array_equal(Y[:12] == 0, Yo == 0)                    # same structural zeros
    assert np.abs(Y[:6] - g["Y"]).max() <= TOL_Y * np.abs(g["Y"]).max()


@pytest.mark.parametrize("name", H.ROBOTS)
def test_projected_blocks_vs_golden(name):
    flat = H.flat_model(name)
    g = np.load(os.path.join(H.GOLDEN_DIR, f"{name}_N48.npz"))
    dm = _dev(flat)
    dev = _up((g["q"], g["dq"], g["ddq"], g["tau"], g["cnt"]))
    A, b, P = dm.projected_batch(*dev, friction=True, want_P=True)
    A, b, P = A.cpu().numpy(), b.cpu().numpy(), P.cpu().numpy()
    assert np.abs(P - g["P"]).max() <= 1e-12
    assert np.abs(P - np.transpose(P, (0, 2, 1))).max() <= 1e-14
    nv = flat.nv
    assert np.abs(A[:4].reshape(4 * nv, -1) - g["A_first"]).max() <= TOL_Y * np.abs(g["A_first"]).max()
    assert np.abs(b.reshape(-1) - g["b"]).max() <= TOL_Y * np.abs(g["b"]).max()
    # without friction columns
    A2, b2 = dm.projected_batch(*dev, friction=False)
    assert A2.shape[-1] == 130 and torch.equal(A2, torch.as_tensor(A[..., :130]).cuda())


@pytest.mark.parametrize("name", H.ROBOTS)
def test_fused_gram_vs_golden_and_stack_path(name):
    from system_identification_b200.ops import gram_from_stack
    flat = H.flat_model(name)
    g = np.load(os.path.join(H.GOLDEN_DIR, f"{name}_N48.npz"))
    dm = _dev(flat)
    dev = _up((g["q"], g["dq"], g["ddq"], g["tau"], g["cnt"]))
    c = 154
    G, r, s, n = H.split_stats(dm.gram_accumulate(*dev).cpu().numpy(), c)
    assert H.rel(G, g["G"]) <= TOL_GRAM and H.rel(r, g["r"]) <= TOL_GRAM and abs(s - g["s"]) <= TOL_GRAM * g["s"]
    assert n == g["n"] == 48 * 18 and np.array_equal(G, G.T)
    # compat path: Gram of the per-sample blocks stacked on the device == fused Gram
    A, b = dm.projected_batch(*dev, friction=True)
    G2, r2, s2, n2 = H.split_stats(gram_from_stack(A.reshape(-1, c).contiguous(), b.reshape(-1)).cpu().numpy(), c)
    assert H.rel(G2, G) <= TOL_GRAM and H.rel(r2, r) <= TOL_GRAM and n2 == n
    # no-friction variant: 130 columns, tau column right after the body columns
    G3, r3, s3, n3 = H.split_stats(dm.gram_accumulate(*dev, friction=False).cpu().numpy(), 130)
    assert H.rel(G3, g["G"][:130, :130]) <= TOL_GRAM and H.rel(r3, g["r"][:130]) <= TOL_GRAM and abs(s3 - s) <= 1e-12 * s


def test_gram_edge_cases_ragged_empty_accumulate_weights_nan():
    flat, data = H.small_log("solo12", 101, seed=41)             # 101 = 3 super-batches + ragged tail, not a multiple of 4
    dm = _dev(flat)
    dev = _up(data)
    full = dm.gram_accumulate(*dev)
    # streaming accumulation over uneven chunks equals one call (the call ADDS into stats)
    acc = torch.zeros_like(full)
    for lo, hi in [(0, 1), (1, 34), (34, 34), (34, 101)]:
        if hi > lo:
            dm.gram_accumulate(*(a[:, lo:hi] for a in dev), stats=acc)
    assert H.rel(acc.cpu().numpy(), full.cpu().numpy()) <= 1e-13
    # integer weights == repeating samples
    w = torch.zeros(101, dtype=torch.float64, device="cuda"); w[[3, 50, 100]] = torch.tensor([2.0, 1.0, 3.0], dtype=torch.float64, device="cuda")
    ws = dm.gram_accumulate(*dev, weights=w).cpu().numpy()
    idx = [3, 3, 50, 100, 100, 100]
    rep = dm.gram_accumulate(*(a[:, idx].contiguous() for a in dev)).cpu().numpy()
    assert H.rel(ws, rep) <= 1e-13
    # a NaN sample is skipped and counted, the rest is unchanged
    bad = [a.clone() for a in dev]
    bad[1][4, 7] = float("nan")
    info = torch.zeros(2, dtype=torch.int64, device="cuda")
    st = dm.gram_accumulate(*bad, info=info).cpu().numpy()
    keep = [i for i in range(101) if i != 7]
    ref = dm.gram_accumulate(*(a[:, keep].contiguous() for a in dev)).cpu().numpy()
    assert info.cpu().tolist() == [0, 1] and np.all(np.isfinite(st)) and H.rel(st, ref) <= 1e-13
    # all-flight and all-stance contact patterns, contact state 2 (quirk Q5)
    q, dq, ddq, tau, cnt = data
    for pattern in (np.zeros_like(cnt), np.ones_like(cnt), 2 * np.ones_like(cnt)):
        d2 = (q[:, :8], dq[:, :8], ddq[:, :8], tau[:, :8], pattern[:, :8])
        _, Po, Ao, bo = H.oracle_blocks(flat, d2)
        _, _, P = dm.projected_batch(*_up(d2), want_P=True)
        assert np.abs(P.cpu().numpy() - Po).max() <= 1e-12


def test_host_streaming_entry_equals_device_entry():
    """sysid_gram_accumulate_host (chunked upload overlapped with the kernel) == the device-resident call; ragged last
    chunk, pinned and pageable sources, weights, and identify() taking host arrays directly."""
    flat, data = H.small_log("g1_12dof", 1000, seed=5)
    dm = _dev(flat)
    full = dm.gram_accumulate(*_up(data)).cpu().numpy()
    host = [np.ascontiguousarray(a, dtype=np.float64) for a in data]
    for chunk in (1000, 384, 37):
        st = dm.gram_accumulate_host(*host, chunk=chunk)
        torch.cuda.synchronize()
        assert H.rel(st.cpu().numpy(), full) <= 1e-13
    pinned = [torch.from_numpy(a).pin_memory() for a in host]
    st = dm.gram_accumulate_host(*pinned, chunk=256)
    torch.cuda.synchronize()
    assert H.rel(st.cpu().numpy(), full) <= 1e-13
    w = np.zeros(1000); w[::3] = 2.0
    sw = dm.gram_accumulate_host(*host, weights=w, chunk=300)
    swd = dm.gram_accumulate(*_up(data), weights=torch.from_numpy(w).cuda())
    torch.cuda.synchronize()
    assert H.rel(sw.cpu().numpy(), swd.cpu().numpy()) <= 1e-13
    # a column slice of a wider host array (leading dimension > N)
    wide = [np.ascontiguousarray(np.concatenate([a, a], axis=1)) for a in host]
    st = dm.gram_accumulate_host(*(a[:, :1000] for a in wide), chunk=512)
    torch.cuda.synchronize()
    assert H.rel(st.cpu().numpy(), full) <= 1e-13


def test_bootstrap_resamples_equal_replicated_logs():
    """BASELINE configs[4] in miniature: every bootstrap resample's statistics equal those of the log with its samples
    (or blocks) physically repeated, and the batched LMI solve of the resamples matches the oracle's solve."""
    from oracle import sdp as osdp
    from system_identification_b200.bootstrap import bootstrap_identify, bootstrap_weights
    from system_identification_b200.sys_identification import SystemIdentification
    flat, data = H.small_log("solo12", 300, seed=9)
    si = SystemIdentification.from_flat_model(flat)
    dm = si.device_model
    dev = _up(data)
    for block, B in ((1, 5), (25, 6)):
        x, info, stats = bootstrap_identify(si, *dev, B=B, block=block, seed=1005, return_stats=True)
        K = (300 + block - 1) // block
        W = bootstrap_weights(K, B, 1005)
        assert x.shape == (B, 154) and np.allclose(W.sum(1), K) and all(int(i["status"]) in (0, 1) for i in info)
        for b in (0, B - 1):
            idx = np.concatenate([np.arange(k * block, min(300, (k + 1) * block)) for k in range(K) for _ in range(int(W[b, k]))])
            rep = dm.gram_accumulate(*(a[:, idx].contiguous() for a in dev)).cpu().numpy()
            assert H.rel(stats[b].cpu().numpy(), rep) <= 1e-12
        G, r, s, n = H.split_stats(stats[0].cpu().numpy(), 154)
        prob = osdp.build_problem(G, r, s, n, 13, flat.phi_prior, flat.robot_mass, flat.ellipsoids, 12)
        xo, _ = osdp.solve_alm(prob)
        assert H.rel(x[0], xo) <= TOL_PHI
    assert np.abs(x[0] - x[1]).max() > 0           # resamples differ


def test_rank_deficient_contact_jacobian_matches_pinv():
    """Two contact frames on the same point make J_c rank deficient whenever both are in stance: numpy's pinv (the reference,
    src/sys_identification.py:134) projects onto the true null space; the fused path (Householder QR with the pinv rank
    rule) and the per-sample path (Cholesky with dropped pivots) must agree with it and count the samples in info[0]."""
    import copy
    flat = copy.deepcopy(H.flat_model("solo12"))
    flat.ee_joint = np.array(flat.ee_joint).copy(); flat.ee_offset = np.array(flat.ee_offset).copy()
    flat.ee_joint[1] = flat.ee_joint[0]; flat.ee_offset[1] = flat.ee_offset[0]          # foot 1 duplicates foot 0
    N = 64
    q, dq, ddq, cnt = synth_log = H.synth.make_trajectory(flat, N, 77)
    tau = H.synth.synth_tau(flat, N, 3)
    cnt = np.array(cnt); cnt[0, :] = 1.0; cnt[1, ::2] = 1.0; cnt[1, 1::2] = 0.0             # both duplicates down on even samples
    data = (q, dq, ddq, tau, cnt)
    dm = _dev(flat)
    dev = _up(data)
    info = torch.zeros(2, dtype=torch.int64, device="cuda")
    G, r, s, n = H.split_stats(dm.gram_accumulate(*dev, info=info).cpu().numpy(), 154)
    t = H.oracle_tree(flat)
    A, b = H.dy.stacked_system(t, *data, flat.ee_names)
    assert H.rel(G, A.T @ A) <= 1e-11 and H.rel(r, A.T @ b) <= 1e-11 and abs(s - b @ b) <= 1e-11 * (b @ b)
    assert info.cpu().tolist()[0] == N // 2 and info.cpu().tolist()[1] == 0
    _, _, P = dm.projected_batch(*dev, want_P=True)
    Po = np.array([H.dy.null_space_projector(t, q[:, i], cnt[:, i], flat.ee_names) for i in range(8)])
    assert np.abs(P[:8].cpu().numpy() - Po).max() <= 1e-11


def test_api_errors_are_reported_not_thrown_across_the_abi():
    from system_identification_b200 import _lib
    from system_identification_b200.ops import DeviceModel
    flat = H.flat_model("solo12")
    dm = DeviceModel(flat)
    q = torch.zeros((19, 4), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        dm.regressor_batch(q[:18], q[:18], q[:18])
    lib = _lib.load()
    rc = lib.sysid_regressor_batch(dm.handle, None, None, None, 4, 4, None, None)
    assert rc == -1 and b"null" in lib.sysid_last_error()
    import copy
    big = copy.deepcopy(flat)
    big.parent = np.concatenate([flat.parent, [13]]).astype(np.int32); big.jtype = np.concatenate([flat.jtype, [1]]).astype(np.int32)
    big.axis = np.vstack([flat.axis, [1, 0, 0]]); big.place_R = np.concatenate([flat.place_R, np.eye(3)[None]]); big.place_p = np.vstack([flat.place_p, [0, 0, 0]])
    h = _lib.create_model(big)                                      # 15 joints: outside the fused kernels' envelope, served by the large-model path
    d = _lib.Dims(); lib.sysid_model_dims(h, __import__("ctypes").byref(d))
    assert (d.nv, d.nbodies) == (19, 14)
    lib.sysid_model_destroy(h)
    huge = copy.deepcopy(flat)
    extra = 30
    huge.parent = np.concatenate([flat.parent, 13 + np.arange(extra)]).astype(np.int32); huge.jtype = np.concatenate([flat.jtype, np.ones(extra)]).astype(np.int32)
    huge.axis = np.vstack([flat.axis] + [[1, 0, 0]] * extra); huge.place_R = np.concatenate([flat.place_R, np.tile(np.eye(3)[None], (extra, 1, 1))]); huge.place_p = np.vstack([flat.place_p, np.zeros((extra, 3))])
    with pytest.raises(_lib.SysidError) as e:
        _lib.create_model(huge)
    assert e.value.code == -2                                       # 44 joints: outside every compiled envelope


@pytest.mark.parametrize("name", H.ROBOTS)
def test_sdp_solve_vs_oracle_fixture(name):
    from system_identification_b200.ops import sdp_solve
    flat = H.flat_model(name)
    g = np.load(os.path.join(H.GOLDEN_DIR, f"{name}_N48.npz"))
    stats = torch.from_numpy(np.concatenate([g["sdp_G"].reshape(-1), g["sdp_r"], [g["sdp_s"], g["sdp_n"]]])).cuda()
    x, info = sdp_solve(stats, 13, 12, flat.phi_prior, flat.ellipsoids, flat.robot_mass)
    x = x[0].cpu().numpy()
    assert info[0]["status"] == 0
    xo = g["sdp_x_alm"]
    assert H.rel(x, xo) <= TOL_PHI and H.rel(x[:130], xo[:130]) <= TOL_PHI
    for i in range(13):
        assert H.rel(x[10 * i:10 * i + 10], xo[10 * i:10 * i + 10]) <= TOL_PHI
    assert H.rel(x, g["sdp_x_barrier"]) <= TOL_PHI                  # and the barrier solve, independently
    assert abs(info[0]["objective"] - g["sdp_obj"]) <= 1e-7 * max(1.0, abs(g["sdp_obj"]))
    assert abs(info[0]["mass_residual"]) <= 1e-9 and info[0]["min_eig_J"] > -1e-7 and info[0]["min_eig_C"] > -1e-7
    assert np.all(x[130:] > -1e-8)


def test_sdp_active_lmi_euclidean_batched_and_failure():
    from system_identification_b200.ops import sdp_solve
    from oracle import sdp as osdp
    flat = H.flat_model("solo12")
    g = np.load(os.path.join(H.GOLDEN_DIR, "solo12_N48.npz"))
    G, r, s, n = g["sdp_G"], g["sdp_r"], float(g["sdp_s"]), float(g["sdp_n"])
    stats = torch.from_numpy(np.concatenate([G.reshape(-1), r, [s, n]])).cuda()
    for lam, reg in [(1e-3, "constant_pullback"), (1e-2, "euclidean")]:
        prob = osdp.build_problem(G, r, s, n, 13, flat.phi_prior, flat.robot_mass, flat.ellipsoids, 12, lambda_reg=lam, reg_type=reg)
        xo, _ = osdp.solve_alm(prob)
        x, info = sdp_solve(stats, 13, 12, flat.phi_prior, flat.ellipsoids, flat.robot_mass, lambda_reg=lam, reg_type=reg)
        assert info[0]["status"] in (0, 1) and H.rel(x[0].cpu().numpy(), xo) <= TOL_PHI
    # batch of 3 scaled problems == 3 separate solves
    batch = torch.stack([stats, stats * 1.0, stats]).contiguous()
    batch[1, :154 * 154 + 154 + 1] *= 1.5                            # a different (still PSD) Gram
    xb, ib = sdp_solve(batch, 13, 12, flat.phi_prior, flat.ellipsoids, flat.robot_mass, batch=3)
    x1, _ = sdp_solve(batch[1].contiguous(), 13, 12, flat.phi_prior, flat.ellipsoids, flat.robot_mass)
    assert torch.allclose(xb[1], x1[0], rtol=0, atol=1e-9) and torch.allclose(xb[0], xb[2], rtol=0, atol=0)
    assert all(int(s_) == 0 for s_ in ib["status"])
    # infeasible: negative total mass -> not optimal -> ValueError through the reference-shaped Solver
    from src.solver import Solver
    sol = Solver.from_stats(stats, 13, flat.phi_prior, -1.0, flat.ellipsoids, ndof=12)
    with pytest.raises(ValueError, match="did not solve to optimality"):
        sol.solve_fully_consistent()


@pytest.mark.parametrize("name", H.ROBOTS)
def test_end_to_end_identify_vs_oracle(name):
    """identify() (fused path) == reference-shaped path (per-sample API -> stack -> Solver) == oracle, on one log."""
    import sys
    sys.path.insert(0, H.ROOT)
    from oracle import dynamics as dy, sdp as osdp
    from src.solver import Solver
    from src.sys_identification import SystemIdentification
    flat = H.flat_model(name)
    si = SystemIdentification.from_flat_model(flat)
    dm = si.device_model

    def regress(q, dq, ddq, cnt):
        dev = _up((q, dq, ddq, np.zeros((12, q.shape[1])), cnt))
        Y = dm.regressor_batch(*dev[:3]).cpu().numpy()
        _, _, P = dm.projected_batch(*dev, want_P=True)
        return Y, P.cpu().numpy()
    data, _, _, _ = H.identifiable_log(flat, 300, 91, regress)
    phi, bv, bc, info = si.identify(*data, return_info=True)
    # oracle on the oracle's own stack
    t = H.oracle_tree(flat)
    A, b = dy.stacked_system(t, *data, flat.ee_names)
    prob = osdp.build_problem(A.T @ A, A.T @ b, float(b @ b), A.shape[0], 13, flat.phi_prior, flat.robot_mass, flat.ellipsoids, 12)
    xo, io = osdp.solve_alm(prob)
    x = np.concatenate([phi, bv, bc])
    assert H.rel(x, xo) <= TOL_PHI
    for i in range(13):
        assert H.rel(phi[10 * i:10 * i + 10], xo[10 * i:10 * i + 10]) <= TOL_PHI
    # torque-prediction metric within 0.5 % of the oracle's, for prior and identified parameters
    for p in (flat.phi_prior.astype(float), phi):
        tot, pj = si.tau_prediction_rmse(*data, p)
        tot_o, pj_o = dy.tau_prediction_rmse(t, *(a[:, :60] for a in data), p, flat.ee_names)
        tot_s, pj_s = si.tau_prediction_rmse(*(a[:, :60] for a in data), p)
        assert abs(tot_s - tot_o) <= TOL_RMSE * tot_o and np.abs(pj_s - pj_o).max() <= TOL_RMSE * pj_o.max()
    # reference-shaped path on a prefix: the demo's loops, vstack, Solver(...).solve_fully_consistent()
    n = 40
    Ys, Ts, Bv, Bc = [], [], [], []
    q, dq, ddq, tau, cnt = (a[:, :n] for a in data)
    for i in range(n):
        y, tp = si.get_proj_regressor_torque(q[:, i], dq[:, i], ddq[:, i], tau[:, i], cnt[:, i])
        b_v, b_c = si.get_proj_friction_regressors(q[:, i], dq[:, i], ddq[:, i], cnt[:, i])
        Ys.append(y); Ts.append(tp); Bv.append(b_v); Bc.append(b_c)
    Ys, Ts, Bv, Bc = np.vstack(Ys), np.hstack(Ts), np.vstack(Bv), np.vstack(Bc)
    assert np.abs(np.hstack([Ys, Bv, Bc]) - A[:18 * n]).max() <= TOL_Y * np.abs(A[:18 * n]).max()
    sol = Solver(Ys, Ts, 13, si.get_phi_prior(), si.get_robot_mass(), si.get_bounding_ellipsoids(), B_v=Bv, B_c=Bc)
    phi_s = sol.solve_fully_consistent()
    phi_f = si.identify(q, dq, ddq, tau, cnt)
    assert H.rel(phi_s, phi_f) <= 1e-6 and phi_s.shape == (130,) and sol._b_v.value.shape == (12,)


def _device_identifiable_tau(flat, dm, dev, seed):
    from system_identification_b200.synth import identifiable_tau_device
    return identifiable_tau_device(flat, dm, dev, seed)


TOL_RMSE_FULL = 1e-9   # deterministic fp64 quantity: the evaluation pass vs the oracle's C twin


@pytest.mark.parametrize("name", H.ROBOTS)
def test_full_size_20k_vs_oracle(name):
    """BASELINE configs[0-2] at their real size (N = 20 000): the fused Gram against the oracle's C twin at 1e-12,
    identify() against BOTH oracle solves (semismooth-Newton ALM and log-barrier) of the ORACLE's own statistics at 1e-4
    global and per link, and the evaluation pass against the oracle's at 1e-9."""
    from oracle import sdp as osdp
    from oracle.cbuild import COracle
    from system_identification_b200.sys_identification import SystemIdentification
    N, c = 20000, 154
    flat = H.flat_model(name)
    q, dq, ddq, cnt = H.synth.make_trajectory(flat, N, H.synth.SEEDS[name])
    si = SystemIdentification.from_flat_model(flat)
    dm = si.device_model
    dev = list(_up((q, dq, ddq, np.zeros((12, N)), cnt)))
    dev[3] = _device_identifiable_tau(flat, dm, dev, seed=23)
    tau = dev[3].cpu().numpy()
    assert np.isfinite(tau).all() and np.abs(tau).max() < 1e4            # a well-posed log: no blown-up least-squares torques
    data = (q, dq, ddq, tau, cnt)
    co = COracle(H.oracle_tree(flat), flat.ee_names)
    so, _ = co.gram(*data)
    Go, ro, s_o, n_o = H.split_stats(so, c)
    st = dm.gram_accumulate(*dev)
    assert _loaded_native()
    G, r, s, n = H.split_stats(st.cpu().numpy(), c)
    assert H.rel(G, Go) <= TOL_GRAM and H.rel(r, ro) <= TOL_GRAM and abs(s - s_o) <= TOL_GRAM * s_o and n == n_o == 18 * N
    assert np.array_equal(G, G.T) and torch.equal(dm.gram_accumulate(*dev), st)       # symmetric, bit-reproducible
    # host-array entry (what read_data returns: float32 q / contact, float64 dq / ddq / tau) == device-resident entry
    sh = dm.gram_accumulate_host(q.astype(np.float32), dq, ddq, tau, cnt.astype(np.float32), chunk=8192)
    torch.cuda.synchronize()
    assert H.rel(sh.cpu().numpy(), st.cpu().numpy()) <= 1e-13
    # stage 3 on the ORACLE's statistics, two independent solves
    prob = osdp.build_problem(Go, ro, s_o, n_o, 13, flat.phi_prior, flat.robot_mass, flat.ellipsoids, 12)
    xa, _ = osdp.solve_alm(prob)
    xb, _ = osdp.solve_barrier(prob)
    phi, bv, bc, info = si.identify(*dev, return_info=True)
    x = np.concatenate([phi, bv, bc])
    for xo in (xa, xb):
        assert H.rel(x, xo) <= TOL_PHI and H.rel(phi, xo[:130]) <= TOL_PHI
        for i in range(13):
            assert H.rel(phi[10 * i:10 * i + 10], xo[10 * i:10 * i + 10]) <= TOL_PHI
    assert info["status"] in (0, 1)
    # evaluation pass (reference print_tau_prediction_rmse) at full size, prior and identified parameters
    for p in (flat.phi_prior.astype(np.float64), phi):
        tot, pj = si.tau_prediction_rmse(*dev, p)
        tot_o, pj_o = co.tau_rmse(*data, p)
        assert abs(tot - tot_o) <= TOL_RMSE_FULL * tot_o and np.abs(pj - pj_o).max() <= TOL_RMSE_FULL * pj_o.max()
    out1 = dm.predict_rmse(*dev, torch.from_numpy(phi))
    assert torch.equal(out1, dm.predict_rmse(*dev, torch.from_numpy(phi)))            # no atomics: bit-reproducible
    # checksum of checksums: x^T G x - 2 x^T r + s is the stack's squared residual; its joint rows are what the
    # evaluation pass sums when friction is left out (quirk Q7), so that part is bounded by the whole
    x0 = np.concatenate([phi, np.zeros(24)])
    ssr = float(x0 @ G @ x0 - 2 * x0 @ r + s)
    tot, _ = si.tau_prediction_rmse(*dev, phi)
    assert 0 < tot * N <= ssr * (1 + 1e-9)


def test_full_size_1m_g1_vs_oracle():
    """BASELINE configs[3], the headline configuration: the 1 000 000-sample G1-12dof log of bench.py.  The fused Gram of
    the whole log against the oracle's C twin (all host cores, a few seconds) at 1e-12; additivity over the 8-rank shard
    boundaries; and the evaluation pass of the first 150 000 samples against the oracle's at 1e-9."""
    from oracle.cbuild import COracle
    N, c = 1_000_000, 154
    flat = H.flat_model("g1_12dof")
    q, dq, ddq, cnt = H.synth.make_trajectory(flat, N, H.synth.SEEDS["g1_1m"])
    tau = H.synth.synth_tau(flat, N, 11, scale=10.0)
    data = (q, dq, ddq, tau, cnt)
    dm = _dev(flat)
    dev = _up(data)
    co = COracle(H.oracle_tree(flat), flat.ee_names)
    so, _ = co.gram(*data)
    st = dm.gram_accumulate(*dev)
    G, r, s, n = H.split_stats(st.cpu().numpy(), c)
    Go, ro, s_o, n_o = H.split_stats(so, c)
    assert H.rel(G, Go) <= TOL_GRAM and H.rel(r, ro) <= TOL_GRAM and abs(s - s_o) <= TOL_GRAM * s_o and n == n_o == 18 * N
    from system_identification_b200.distributed import shard_bounds
    acc = torch.zeros_like(st)
    for rk in range(8):
        lo, hi = shard_bounds(N, rk, 8)
        dm.gram_accumulate(*(a[:, lo:hi] for a in dev), stats=acc)
    assert H.rel(acc.cpu().numpy(), so) <= TOL_GRAM
    M = 150_000
    phi = flat.phi_prior.astype(np.float64)
    out = dm.predict_rmse(*(a[:, :M] for a in dev), torch.from_numpy(phi)).cpu().numpy()
    tot_o, pj_o = co.tau_rmse(*(a[:, :M] for a in data), phi)
    assert abs(out[0] - tot_o) <= TOL_RMSE_FULL * tot_o and np.abs(out[1:] - pj_o).max() <= TOL_RMSE_FULL * pj_o.max()


def test_full_size_properties_20k():
    """Size-independent properties at N = 20 000 (additivity over shards, linearity in tau)."""
    flat, data = H.small_log("spot", 20000)
    dm = _dev(flat)
    dev = _up(data)
    c = 154
    st = dm.gram_accumulate(*dev)
    G, r, s, n = H.split_stats(st.cpu().numpy(), c)
    assert n == 18 * 20000 and np.array_equal(G, G.T) and np.linalg.eigvalsh(G).min() >= -1e-9 * np.abs(G).max()
    # additivity over shards (the multi-GPU reduction in miniature)
    halves = dm.gram_accumulate(*(a[:, :9000] for a in dev))
    dm.gram_accumulate(*(a[:, 9000:] for a in dev), stats=halves)
    assert H.rel(halves.cpu().numpy(), st.cpu().numpy()) <= 1e-13
    # linearity in tau: r(tau1 + tau2) = r(tau1) + r(tau2); G does not depend on tau
    tau2 = torch.roll(dev[3], 7, dims=1)
    r1 = dm.gram_accumulate(*dev)[c * c:c * c + c]
    r2 = dm.gram_accumulate(dev[0], dev[1], dev[2], tau2, dev[4])[c * c:c * c + c]
    r12 = dm.gram_accumulate(dev[0], dev[1], dev[2], (dev[3] + tau2).contiguous(), dev[4])[c * c:c * c + c]
    assert H.rel((r1 + r2).cpu().numpy(), r12.cpu().numpy()) <= 1e-12


TOL_FILT = 1e-10       # relative to the max-abs of the scipy result (fp64 recursions / fits in a different summation order)


def test_filters_vs_scipy_fixture_and_live():
    """sysid_filtfilt / sysid_savgol == scipy.signal.filtfilt / savgol_filter as read_data calls them
    (reference demo/solo_identification.py:15-32).  Checked against the scipy-generated fixture (always) and against scipy
    itself on longer, multi-chunk signals (scipy is importable in this image); plus properties at full size."""
    from system_identification_b200 import filters as F
    g = np.load(os.path.join(H.GOLDEN_DIR, "filters_scipy.npz"))
    b, a = F.butter_lowpass(5, 0.15)
    up = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda()
    for xk, yk in (("x", "filtfilt"), ("x_short", "filtfilt_short")):
        y = F.filtfilt(b, a, up(g[xk]), float32_input=True).cpu().numpy()      # the fixture filters a float32 log, as the reference does
        assert np.abs(y - g[yk]).max() <= TOL_FILT * np.abs(g[yk]).max()
        y64 = F.filtfilt(b, a, up(g[xk])).cpu().numpy()                       # fp64 pads: differs only by the float32 rounding of the pads
        assert 0 < np.abs(y64 - g[yk]).max() <= 1e-6 * np.abs(g[yk]).max()
    for xk, yk in (("x", "savgol_f64"), ("x_w21", "savgol_w21")):
        y = F.savgol_filter(up(g[xk]), 21, 5).cpu().numpy()
        assert np.abs(y - g[yk]).max() <= TOL_FILT * np.abs(g[yk]).max()
    # on a float32 log scipy runs savgol_filter IN float32 (and returns float32): the reference's own result carries that
    # rounding; the fp64 device result agrees with it to float32 precision
    assert g["savgol"].dtype == np.float32
    y = F.savgol_filter(up(g["x"]), 21, 5).cpu().numpy()
    assert np.abs(y - g["savgol"]).max() <= 3e-5 * np.abs(g["savgol"]).max()
    # float32_input=True returns float32-representable values like scipy does for a float32 log; scipy's own float32 path is
    # not a rounding of the fp64 result (1.15+ runs the edge polynomial fit in float32), so float32 precision is all there is
    y32 = F.savgol_filter(up(g["x"]), 21, 5, float32_input=True).cpu().numpy()
    assert np.array_equal(y32, y32.astype(np.float32).astype(np.float64))
    assert np.abs(y32 - g["savgol"]).max() <= 3e-5 * np.abs(g["savgol"]).max()
    assert _loaded_native()
    # reference error behaviour
    with pytest.raises(ValueError):
        F.filtfilt(b, a, up(g["x"][:, :18]))                       # N <= padlen
    with pytest.raises(ValueError):
        F.savgol_filter(up(g["x"][:, :20]), 21, 5)                 # window longer than the signal
    # in place, and a column slice of a wider array (leading dimension > N)
    xd = up(g["x"])
    wide = torch.cat([xd, xd], dim=1)
    y = F.filtfilt(b, a, wide[:, :700], out=torch.empty_like(wide)[:, :700], float32_input=True).cpu().numpy()
    assert np.abs(y - g["filtfilt"]).max() <= TOL_FILT * np.abs(g["filtfilt"]).max()
    xi = xd.clone(); F.filtfilt(b, a, xi, out=xi, float32_input=True)
    assert np.abs(xi.cpu().numpy() - g["filtfilt"]).max() <= TOL_FILT * np.abs(g["filtfilt"]).max()
    # multi-chunk signals against scipy itself, other filter orders / windows
    signal = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(7)
    x = (np.cumsum(rng.standard_normal((3, 20000)), axis=1) * 0.01 + rng.standard_normal((3, 20000))).astype(np.float32).astype(np.float64)
    for order, wn in ((5, 0.15), (2, 0.3), (8, 0.1)):
        bb, aa = F.butter_lowpass(order, wn)
        ref = signal.filtfilt(bb, aa, x, axis=1)
        y = F.filtfilt(bb, aa, up(x)).cpu().numpy()
        assert np.abs(y - ref).max() <= TOL_FILT * np.abs(ref).max()
    for W, p in ((21, 5), (5, 2), (51, 3)):
        ref = signal.savgol_filter(x, W, p)
        y = F.savgol_filter(up(x), W, p).cpu().numpy()
        assert np.abs(y - ref).max() <= TOL_FILT * np.abs(ref).max()
    # size-independent properties at 1 M samples: unit DC gain, linearity, time-reversal symmetry of the zero-phase filter
    N = 1_000_000
    xl = torch.randn((4, N), dtype=torch.float64, device="cuda")
    const = torch.full((1, N), 3.25, dtype=torch.float64, device="cuda")
    assert (F.filtfilt(b, a, const) - 3.25).abs().max().item() <= 1e-11
    assert (F.savgol_filter(const, 21, 5) - 3.25).abs().max().item() <= 1e-12
    y1, y2 = F.filtfilt(b, a, xl[:2]), F.filtfilt(b, a, xl[2:])
    y12 = F.filtfilt(b, a, (2.0 * xl[:2] - 0.5 * xl[2:]).contiguous())
    assert (y12 - (2.0 * y1 - 0.5 * y2)).abs().max().item() <= 1e-11
    yr = F.filtfilt(b, a, torch.flip(xl[:2], dims=[1]).contiguous())
    # forward-backward == backward-forward away from the edge transients (the two passes commute; only the padding differs)
    assert (torch.flip(yr, dims=[1]) - y1)[:, 5000:-5000].abs().max().item() <= 1e-9


def test_presolve_warm_start_same_optimum_fewer_steps():
    """identify() on host arrays with the LMI pre-solve hidden behind the stream == the cold solve (the optimum is unique), with
    fewer Newton steps in the final solve; an invalid / non-finite warm record is ignored; the plan-based solve equals the
    one-shot C-ABI entry."""
    from system_identification_b200.ops import SdpPlan, sdp_solve
    from system_identification_b200.sys_identification import SystemIdentification
    flat = H.flat_model("g1_12dof")
    N = 80000
    q, dq, ddq, cnt = H.synth.make_trajectory(flat, N, 77)
    si = SystemIdentification.from_flat_model(flat)
    dm = si.device_model
    dev = list(_up((q, dq, ddq, np.zeros((12, N)), cnt)))
    tau = _device_identifiable_tau(flat, dm, dev, seed=5).cpu().numpy()
    host = (q.astype(np.float32), dq, ddq, tau, cnt.astype(np.float32))
    from system_identification_b200.identify import identify
    phi_c, bv_c, bc_c, info_c = identify(si, *host, return_info=True, presolve=False, chunk=16384)
    phi_w, bv_w, bc_w, info_w = identify(si, *host, return_info=True, presolve="force", chunk=16384)
    assert info_c["status"] == 0 and info_w["status"] == 0 and info_w["presolve_status"] in (0, 1)
    xw, xc = np.concatenate([phi_w, bv_w, bc_w]), np.concatenate([phi_c, bv_c, bc_c])
    assert H.rel(xw, xc) <= 1e-6
    for i in range(13):
        assert H.rel(phi_w[10 * i:10 * i + 10], phi_c[10 * i:10 * i + 10]) <= 1e-5
    assert info_w["iterations"] < info_c["iterations"]
    # second pre-solve stage (statistics of the first 45 % of the log, warm-started from the first stage): same optimum again
    import system_identification_b200.identify as idm
    keep = idm.PRESOLVE_REFINE_MIN_LOG
    try:
        idm.PRESOLVE_REFINE_MIN_LOG = 0
        phi_r, bv_r, bc_r, info_r = identify(si, *host, return_info=True, presolve="force", chunk=16384)
    finally:
        idm.PRESOLVE_REFINE_MIN_LOG = keep
    assert info_r["status"] == 0 and H.rel(np.concatenate([phi_r, bv_r, bc_r]), xc) <= 1e-6 and info_r["iterations"] < info_c["iterations"]
    # plan-based solve == one-shot entry; garbage warm records are ignored
    st = dm.gram_accumulate(*_up((q, dq, ddq, tau, cnt)))
    plan = SdpPlan(13, 12, flat.phi_prior, flat.ellipsoids, flat.robot_mass)
    x1, i1 = plan.solve(st)
    x0, i0 = sdp_solve(st, 13, 12, flat.phi_prior, flat.ellipsoids, flat.robot_mass)
    assert torch.equal(x1, x0) and int(i1[0]["iterations"]) == int(i0[0]["iterations"])
    bad = torch.full((plan.wlen,), float("nan"), dtype=torch.float64, device="cuda")
    x2, i2 = plan.solve(st, warm=bad)
    assert torch.equal(x2, x0)
    zero = torch.zeros(plan.wlen, dtype=torch.float64, device="cuda")           # valid flag 0
    x3, _ = plan.solve(st, warm=zero)
    assert torch.equal(x3, x0)


def test_bootstrap_1024_solo_20k_full_size():
    """BASELINE configs[4] at full size: 1024 moving-block resamples of the 20 000-sample Solo-12 log in three launches
    (segmented fused kernel -> resample combination on the tensor pipe -> batched LMI solve).  Per-block statistics against the
    oracle's C twin, resample statistics against physically replicated logs, fits against the oracle's solve."""
    from oracle import sdp as osdp
    from oracle.cbuild import COracle
    from system_identification_b200.bootstrap import bootstrap_identify, bootstrap_weights
    from system_identification_b200.ops import combine_stats
    from system_identification_b200.sys_identification import SystemIdentification
    N, B, block, c = 20000, 1024, 100, 154
    flat = H.flat_model("solo12")
    q, dq, ddq, cnt = H.synth.make_trajectory(flat, N, H.synth.SEEDS["solo12_bootstrap"])
    si = SystemIdentification.from_flat_model(flat)
    dm = si.device_model
    dev = list(_up((q, dq, ddq, np.zeros((12, N)), cnt)))
    dev[3] = H.synth.identifiable_tau_device(flat, dm, dev, seed=31, perturb=0.1, bv_max=0.02, bc_max=0.05, noise=0.05)
    data = (q, dq, ddq, dev[3].cpu().numpy(), cnt)
    # per-block statistics: one launch; blocks 0, 77 and the last (ragged: 20 000 = 200 x 100, so also try block = 96)
    co = COracle(H.oracle_tree(flat), flat.ee_names)
    for blk in (100, 96):
        pb = dm.gram_blocks(*dev, blk).cpu().numpy()
        K = (N + blk - 1) // blk
        assert pb.shape == (K, c * c + c + 2)
        for k in (0, 77, K - 1):
            so, _ = co.gram(*(a[:, k * blk:min(N, (k + 1) * blk)] for a in data))
            assert H.rel(pb[k], so) <= 1e-12
        assert H.rel(pb.sum(0), dm.gram_accumulate(*dev).cpu().numpy()) <= 1e-13
    x, info, stats = bootstrap_identify(si, *dev, B=B, block=block, seed=1005, return_stats=True)
    assert x.shape == (B, c) and set(int(i["status"]) for i in info) <= {0, 1}
    K = N // block
    W = bootstrap_weights(K, B, 1005)
    for b in (0, 511, 1023):
        idx = np.concatenate([np.arange(k * block, (k + 1) * block) for k in range(K) for _ in range(int(W[b, k]))])
        rep = dm.gram_accumulate(*(a[:, idx].contiguous() for a in dev)).cpu().numpy()
        assert H.rel(stats[b].cpu().numpy(), rep) <= 1e-12
        G, r, s, n = H.split_stats(rep, c)
        prob = osdp.build_problem(G, r, s, n, 13, flat.phi_prior, flat.robot_mass, flat.ellipsoids, 12)
        xo, _ = osdp.solve_alm(prob)
        assert H.rel(x[b], xo) <= TOL_PHI
    # the combination kernel against a float64 matmul
    Wd = torch.from_numpy(W[:37]).cuda()
    pb = dm.gram_blocks(*dev, block)
    assert H.rel(combine_stats(Wd, pb).cpu().numpy(), (Wd @ pb).cpu().numpy()) <= 1e-14
    # spread of the resampled fits around the full-log fit: non-degenerate and centred
    phi_full = si.identify(*dev)
    assert np.abs(x[:, :130].mean(0) - phi_full).max() <= 5 * x[:, :130].std(0).max() and x[:, 0].std() > 0
