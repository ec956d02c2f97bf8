"""No-GPU checks: the C-ABI library loads and exports every symbol the header declares; host-side logic
(synthetic logs, sharding, reference-shaped API objects) behaves; the N>1 reduction path works over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from system_identification_b200 import _lib, synth
from system_identification_b200 import distributed as D

HEADER = os.path.join(H.ROOT, "include", "sysid_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sysid_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from system_identification_b200.build import build
    lib_path = build()
    names = _declared_functions()
    assert len(names) >= 17
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sysid_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes signatures and header are out of sync"
    lib.sysid_abi_version.restype = ctypes.c_int
    assert lib.sysid_abi_version() == 1
    lim = _lib.Limits()
    lib.sysid_get_limits(ctypes.byref(lim))
    assert (lim.max_bodies, lim.max_nv, lim.max_ee, lim.max_cols_padded) == (13, 18, 4, 160)


def test_struct_layouts_match_header_sizes():
    # sysid_sdp_info: 4 x int32 + 7 x double
    assert ctypes.sizeof(_lib.SdpInfo) == 16 + 7 * 8 == _lib.SDP_INFO_DTYPE.itemsize
    assert ctypes.sizeof(_lib.Dims) == 7 * 4


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from system_identification_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.DeviceModel(H.flat_model("solo12"))


def test_product_does_not_import_oracle():
    pkg = os.path.join(H.ROOT, "system_identification_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
    for f in ("src/solver.py", "src/sys_identification.py"):
        assert "oracle" not in open(os.path.join(H.ROOT, f)).read()


@pytest.mark.parametrize("name", H.ROBOTS)
def test_synthetic_log_shapes_and_conventions(name):
    flat = H.flat_model(name)
    q, dq, ddq, cnt = synth.make_trajectory(flat, 2000, synth.SEEDS[name])
    tau = synth.synth_tau(flat, 2000, 1)
    assert q.shape == (19, 2000) and dq.shape == ddq.shape == (18, 2000) and tau.shape == (12, 2000)
    assert cnt.shape == (flat.n_ee, 2000)
    assert np.all(np.isfinite(q)) and np.all(np.isfinite(ddq))
    assert np.array_equal(q, q.astype(np.float32).astype(np.float64))          # quirk Q8: q went through float32
    assert np.abs(np.linalg.norm(q[3:7], axis=0) - 1).max() < 1e-6
    vals = set(np.unique(cnt))
    assert vals <= ({0.0, 1.0} if flat.n_ee == 4 else {0.0, 1.0, 2.0})
    lo, hi = flat.lower[2:, None], flat.upper[2:, None]
    assert np.all(q[7:] >= lo - 1e-6) and np.all(q[7:] <= hi + 1e-6)
    # analytic joint velocity is the derivative of the (pre-rounding) joint position
    fd = (q[7:, 2:] - q[7:, :-2]) * 500.0 / 2
    assert np.abs(fd - dq[6:, 1:-1]).max() < 5e-2 * max(1.0, np.abs(dq[6:]).max())


def test_shard_bounds_cover_and_balance():
    for N, ws in [(1000000, 8), (20000, 3), (7, 8), (0, 4)]:
        b = [D.shard_bounds(N, r, ws) for r in range(ws)]
        assert b[0][0] == 0 and b[-1][1] == N
        assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import helpers as H
from oracle import dynamics as dy
from system_identification_b200 import distributed as D
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, ws = D.world()
flat, data = H.small_log("solo12", 10, seed=31)
lo, hi = D.shard_bounds(10, rank, ws)
t = H.oracle_tree(flat)
A, b = dy.stacked_system(t, *(a[:, lo:hi] for a in data), flat.ee_names)      # CPU stand-in for the rank-local kernel
G, r, s, n = dy.gram_from_stack(A, b)
stats = torch.from_numpy(np.concatenate([G.reshape(-1), r, [s, n]]))
D.allreduce_stats(stats)
x = torch.full((3,), float(rank)); D.broadcast_solution(x)
if rank == 0:
    Af, bf = dy.stacked_system(t, *data, flat.ee_names)
    Gf, rf, sf, nf = dy.gram_from_stack(Af, bf)
    c = Af.shape[1]
    Gs, rs, ss, ns = H.split_stats(stats.numpy(), c)
    assert H.rel(Gs, Gf) < 1e-13 and H.rel(rs, rf) < 1e-13 and abs(ss - sf) < 1e-12 * sf and ns == nf, "sharded stats differ"
assert float(x[0]) == 0.0
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
'''


def test_two_rank_gloo_allreduce_of_sharded_statistics(tmp_path):
    """World size 2 on CPU (gloo): shard the log, reduce the packed [G|r|s|n], compare with the unsharded statistics."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), H.ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o


def test_reference_shaped_classes_import_without_gpu():
    sys.path.insert(0, H.ROOT)
    from src.solver import Solver
    from src.sys_identification import SystemIdentification
    flat = H.flat_model("spot")
    si = SystemIdentification.from_flat_model(flat)
    assert si.get_num_links() == 13 and si.get_robot_mass() == 34.0 and si.nq == 19 and si.nv == 18 and si.joints_dof == 12
    assert si.get_phi_prior().dtype == np.float32 and len(si.get_bounding_ellipsoids()) == 13
    eig = si.get_physical_consistency(si.get_phi_prior())
    assert len(eig) == 5 and min(eig[2]) > 0                              # prior pseudo-inertias are PD
    s = Solver(np.zeros((36, 130)), np.zeros(36), 13, si.get_phi_prior(), 34.0, si.get_bounding_ellipsoids(),
               B_v=np.zeros((36, 12)), B_c=np.zeros((36, 12)))
    assert s._num_samples == 36 and s._nx == 130 and s.ndof == 12          # n = ROWS of the stack (quirk Q4)
    with pytest.raises(NotImplementedError):
        si.get_full_regressor_force(None, None, None, None, None, None)


def test_print_inertial_params_matches_reference_table_format(capsys):
    from src.sys_identification import SystemIdentification
    si = SystemIdentification.from_flat_model(H.flat_model("spot"))
    prior = si.get_phi_prior().copy()
    ident = prior.astype(np.float64) * 1.01
    si.print_inertial_params(prior, ident)
    out = capsys.readouterr().out
    assert '--------------- Inertial Parameters of "front_left_hip" ---------------' in out     # RUN_DEMO.md:11
    assert "|Parameter    |A priori     |Identified   |Change       |error %      |" in out        # RUN_DEMO.md:12
    assert "|mass (kg)    |     1.680000|     1.696800|     0.016800|          1.0|" in out
    assert "|c_x (m)      |    -0.005374|" in out and "Robot total mass:" in out


def test_butterworth_design_matches_scipy_and_fixture():
    """butter_lowpass restates scipy.signal.butter(order, wn, 'low'): checked live against scipy (the reference's own
    dependency, importable here) and against the coefficients stored in the scipy-generated fixture."""
    from system_identification_b200.filters import butter_lowpass
    g = np.load(os.path.join(H.GOLDEN_DIR, "filters_scipy.npz"))
    b, a = butter_lowpass(5, 0.15)
    assert np.abs(b - g["b"]).max() <= 1e-15 and np.abs(a - g["a"]).max() <= 1e-14
    signal = pytest.importorskip("scipy.signal")
    for order, wn in ((1, 0.9), (2, 0.3), (5, 0.15), (8, 0.05)):
        b, a = butter_lowpass(order, wn)
        bs, as_ = signal.butter(order, wn, btype="low", analog=False)
        assert np.abs(b - bs).max() <= 1e-14 * np.abs(bs).max() and np.abs(a - as_).max() <= 1e-13 * np.abs(as_).max()
    with pytest.raises(ValueError):
        butter_lowpass(5, 1.5)


def test_ingest_column_lists_match_the_reference_scripts():
    """Host logic of the ingest row without a GPU: the product's column lists are csv2dat.py's (restated independently in
    the oracle, which is pinned against the script's own output), and the device entry points refuse to run on a CPU."""
    from oracle import ingest as oi
    from system_identification_b200 import ingest
    assert ingest.LOW_Q_COLS == oi.LOW_Q_COLS and ingest.ODOM_Q_COLS == oi.ODOM_Q_COLS
    assert ingest.DQ_COLS == oi.DQ_COLS and ingest.TAU_COLS == oi.TAU_COLS
    assert ingest.ACCEL_COLS + ["body_ang_acceleration_" + a for a in "xyz"] + [f"low_motor_{i}_ddq" for i in range(12)] == oi.ddq_cols(True)
    import torch
    if not torch.cuda.is_available():
        for call in (lambda: ingest.load_dat(b"1.0\t2.0\n"), lambda: ingest.load_csv(b"a,b\n1,2\n"),
                     lambda: ingest.fd_rate(np.arange(4.0), np.zeros((2, 4))), lambda: ingest.round_dat(np.zeros(3))):
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                call()


@pytest.mark.parametrize("name", H.ROBOTS)
def test_host_consistency_mirror_equals_oracle(name):
    """SystemIdentification.get_physical_consistency (host, no GPU) == oracle/consistency.py, the restatement of reference
    src/sys_identification.py:324-389, on the prior and on perturbed (partly inconsistent) parameter vectors."""
    from oracle import consistency as oc
    from system_identification_b200.sys_identification import SystemIdentification
    flat = H.flat_model(name)
    si = SystemIdentification.from_flat_model(flat)
    rng = np.random.default_rng(7)
    prior = np.asarray(si.get_phi_prior(), dtype=np.float64)
    for phi in (prior, prior * (1 + 0.5 * rng.standard_normal(prior.size))):
        got = si.get_physical_consistency(phi)
        ref = oc.physical_consistency(phi, flat.ellipsoids)
        for a, b in zip(got, ref):
            assert np.array_equal(np.asarray(a), np.asarray(b))
    # a consistent prior has non-negative margins everywhere (the reference's acceptance rule, :325-326)
    ref = oc.physical_consistency(prior, flat.ellipsoids)
    assert min(np.min(np.real(v)) for v in ref[:3]) > -1e-6
