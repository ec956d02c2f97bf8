"""CPU tests of the ingest oracle (oracle/ingest.py) against the reference's own scripts: the committed fixture
tests/golden/ingest_g1.npz holds what /root/reference/g1-data/{low_ddq_contact_tick,low_ddq,csv2dat}.py and np.loadtxt
produced for a synthetic logger CSV (tests/golden/make_ingest_golden.py); np.savetxt / np.loadtxt are also called live."""
import io
import os

import numpy as np
import pandas as pd
import pytest

import helpers as H
from oracle import ingest as oi

FILES = ("low_q", "odom_q", "dq", "ddq", "tau", "contact")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(H.GOLDEN_DIR, "ingest_g1.npz")))


def same(a, b):
    """Bit-for-bit equality of float arrays, NaN == NaN, -0.0 != +0.0."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.int64) | (np.isnan(a) * -1), b.view(np.int64) | (np.isnan(b) * -1))


def test_parse_dat_text_equals_loadtxt_on_the_reference_files(gold):
    for name in FILES:
        got = oi.parse_dat_text(gold["dat_text_" + name].tobytes())
        assert got.dtype == np.float32 and same(got, gold["loadtxt_" + name]), name


def test_parse_dat_text_equals_loadtxt_live():
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.normal(0, 50.0, (5, 301)), [np.r_[np.nan, np.inf, -np.inf, -0.0, 1e-7, 9e9, rng.normal(0, 1, 295)]]])
    buf = io.BytesIO()
    np.savetxt(buf, x, delimiter="\t", fmt="%.6f")
    for dtype in (np.float32, np.float64):
        ref = np.loadtxt(io.BytesIO(buf.getvalue()), delimiter="\t", dtype=dtype)
        assert same(oi.parse_dat_text(buf.getvalue(), dtype=dtype), ref)
    with pytest.raises(ValueError):
        oi.parse_dat_text(b"1.0\t2.0\n3.0\n")
    with pytest.raises(ValueError):
        np.loadtxt(io.BytesIO(b"1.0\t2.0\n3.0\n"), delimiter="\t")


def test_fd_rate_and_contact_equal_the_reference_scripts(gold):
    df = pd.read_csv(io.BytesIO(gold["csv_text"].tobytes()))
    dq = np.stack([df[f"low_motor_{i}_dq"].to_numpy() for i in range(12)])
    gyro = np.stack([df[f"low_imu_gyro_{a}"].to_numpy() for a in "xyz"])
    assert same(oi.fd_rate(df["low_tick"].to_numpy(), dq, 1000.0), gold["upd_ddq"])
    assert same(oi.fd_rate(df["low_tick"].to_numpy(), gyro, 1000.0), gold["upd_body_acc"])
    assert same(oi.fd_rate(df["timestamp"].to_numpy(), dq, 1.0), gold["plain_ddq"])           # low_ddq.py: no x1000
    # every branch is present in the fixture
    u = gold["upd_ddq"]
    assert np.isnan(u[:, 0]).all() and np.isnan(u[:, 7]).all() and (u[:, 15] == 0).all() and np.isnan(u[:, 23]).all()
    tau = np.stack([df["low_motor_4_tau_est"].to_numpy(), df["low_motor_10_tau_est"].to_numpy()])
    assert same(oi.contact_from_tau(tau), gold["upd_contact"])
    assert set(np.unique(gold["upd_contact"])) == {0.0, 1.0, 2.0}


def test_round_dat_equals_savetxt_loadtxt():
    rng = np.random.default_rng(12)
    x = np.r_[rng.normal(0, 30.0, 4000), (rng.integers(-4000000, 4000000, 3000) + 0.5) * 1e-6,       # near-ties
              [0.0000005, 0.0000015, 0.0000025, -0.0000005, 0.125 + 2.0 ** -21, -0.0, -1e-9, 8.6e9, 1e12, np.nan, np.inf, -np.inf]]
    buf = io.BytesIO()
    np.savetxt(buf, x[None, :], delimiter="\t", fmt="%.6f")
    for f32, dtype in ((True, np.float32), (False, np.float64)):
        ref = np.loadtxt(io.BytesIO(buf.getvalue()), delimiter="\t", dtype=dtype)
        assert same(oi.round_dat(x, float32=f32), ref)


def test_csv_to_log_equals_the_reference_pipeline(gold):
    """calculate_low_motor_ddq -> to_csv -> csv2dat -> loadtxt, against the oracle's file-free composition."""
    df = pd.read_csv(io.BytesIO(gold["csv_text"].tobytes()))
    log = oi.csv_to_log(df, fix_ddq_off_by_one=False)
    for name in FILES:
        assert same(log[name], gold["loadtxt_" + name]), name
    assert log["ddq"].shape[0] == 17                                 # the reference's off-by-one (csv2dat.py:36)
    fixed = oi.csv_to_log(df, fix_ddq_off_by_one=True)
    assert fixed["ddq"].shape[0] == 18 and same(fixed["ddq"][7:], log["ddq"][6:]) and same(fixed["ddq"][:6], log["ddq"][:6])
