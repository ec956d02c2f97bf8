"""GPU parity of SURVEY 8f row f4 through the C-ABI: sysid_tsqr behind Solver.solve_llsq_svd (reference src/solver.py:32-39)
and sysid_physical_consistency (reference src/sys_identification.py:324-389)."""
import numpy as np
import pytest

import helpers as H
from oracle import llsq

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,c", [(5037, 130), (64, 154), (1, 7), (20011, 154), (300, 159)])
def test_tsqr_triangle_properties_and_numpy_qr(rows, c):
    from system_identification_b200 import ops
    rng = np.random.default_rng(rows + c)
    A = rng.normal(0, 1.0, (rows, c)) * rng.uniform(1e-3, 1e3, c)          # badly scaled columns
    b = rng.normal(0, 1.0, rows)
    Ra = ops.tsqr(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    assert np.array_equal(np.tril(Ra, -1), np.zeros_like(Ra))               # upper triangular
    Aa = np.c_[A, b]
    assert H.rel(Ra.T @ Ra, Aa.T @ Aa) <= 1e-13
    # same factor as LAPACK's QR up to the sign of each row
    Rn = np.linalg.qr(Aa, mode="r")
    k = min(rows, c + 1)
    assert np.abs(np.abs(Ra[:k]) - np.abs(Rn[:k])).max() <= 1e-11 * np.abs(Rn).max()
    # deterministic: bit-identical on a second run
    Rb = ops.tsqr(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    assert np.array_equal(Ra, Rb)
    if rows > c + 1:
        x = np.linalg.solve(Ra[:c, :c], Ra[:c, c])
        xo = np.linalg.lstsq(A, b, rcond=None)[0]
        assert H.rel(x, xo) <= 1e-9
        assert abs(abs(Ra[c, c]) - np.linalg.norm(A @ xo - b)) <= 1e-10 * np.linalg.norm(b)


def test_solve_llsq_svd_full_rank_vs_reference_restatement():
    from system_identification_b200.solver import Solver
    rng = np.random.default_rng(41)
    Y = rng.normal(0, 1.0, (3000, 130)) * rng.uniform(1e-2, 1e2, 130)
    tau = Y @ rng.normal(0, 1.0, 130) + rng.normal(0, 1e-3, 3000)
    x = Solver(Y, tau, 13, np.zeros(130), 1.0, []).solve_llsq_svd()
    assert H.rel(x, llsq.solve_llsq_svd(Y, tau)) <= 1e-9


@pytest.mark.parametrize("name", H.ROBOTS)
def test_solve_llsq_svd_on_the_projected_regressor(name):
    """The real stack is rank deficient (only base parameters are identifiable): the minimum-norm solutions agree where
    the problem determines them -- in the predictions and in the retained singular subspace."""
    from system_identification_b200.solver import Solver
    flat, data = H.small_log(name, 48)
    _, _, A, b = H.oracle_blocks(flat, data)
    Y = np.ascontiguousarray(A[:, :130])
    s = Solver(Y, b, 13, flat.phi_prior, flat.robot_mass, flat.ellipsoids)
    x = s.solve_llsq_svd()
    xo = llsq.solve_llsq_svd(Y, b)
    sv = np.linalg.svd(Y, compute_uv=False)
    got = s._singular_values.cpu().numpy()
    keep = sv > 1e-10 * sv[0]
    assert keep.sum() < 130                                                  # rank deficient indeed
    assert np.abs(got[keep] - sv[keep]).max() <= 1e-12 * sv[0]                # same singular values as the stack
    # the reference's answer is itself fragile here: singular values a few ulps above pinv's 1e-15 cutoff survive and
    # their 1/sigma amplifies rounding noise into entries of order 1e6 (in numpy's SVD of the stack just as in the SVD of
    # R), so the two solutions are compared where the data determine them
    assert H.rel(Y @ x, Y @ xo) <= 1e-5
    _, _, Vt = np.linalg.svd(Y, full_matrices=False)
    well = sv > 1e-6 * sv[0]
    assert np.abs(Vt[well] @ (x - xo)).max() <= 1e-6 * np.abs(Vt[well] @ xo).max()


@pytest.mark.parametrize("name", H.ROBOTS)
def test_physical_consistency_batch_vs_oracle(name):
    """consistency_kernel (sysid_physical_consistency) against oracle/consistency.py, the restatement of reference
    src/sys_identification.py:324-389 -- and the host mirror (SystemIdentification.get_physical_consistency) against the
    same oracle (bit for bit: both run numpy's float32 eigvals on the same float32 matrices)."""
    from oracle import consistency as oc
    from system_identification_b200.sys_identification import SystemIdentification
    flat = H.flat_model(name)
    si = SystemIdentification.from_flat_model(flat)
    rng = np.random.default_rng(43)
    prior = np.asarray(si.get_phi_prior(), dtype=np.float64)
    phis = np.stack([prior] + [prior * (1 + 0.3 * rng.standard_normal(prior.size)) for _ in range(40)])
    out = si.get_physical_consistency_batch(phis)
    assert out.shape == (41, 5, 13)
    for i in (0, 1, 17, 40):
        ref = np.array([np.real(np.asarray(v, dtype=np.complex128)) for v in oc.physical_consistency(phis[i], flat.ellipsoids)])   # (5, L)
        host = np.array([np.real(np.asarray(v, dtype=np.complex128)) for v in si.get_physical_consistency(phis[i])])
        assert np.array_equal(host, ref), (name, i)
        for k in range(5):
            scale = max(np.abs(phis[i]).max(), np.abs(ref[k]).max(), 1.0) if k < 4 else np.abs(ref[k]).max()
            # the reference's eigvals run in float32 (its matrices are np.float32): float32 agreement is all there is
            assert np.abs(out[i, k] - ref[k]).max() <= 2e-5 * scale, (name, i, k)
