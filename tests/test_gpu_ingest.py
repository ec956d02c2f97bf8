"""GPU parity of the log ingest, through the C-ABI: BIT-EXACT against the outputs of the reference's
own scripts (tests/golden/ingest_g1.npz, made by tests/golden/make_ingest_golden.py from
/root/reference/g1-data/{low_ddq_contact_tick,low_ddq,csv2dat}.py + np.loadtxt), against np.loadtxt / np.savetxt called
live, and against the oracle restatement on larger seeded inputs."""
import io
import os

import numpy as np
import pytest

import helpers as H
from oracle import ingest as oi

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FILES = ("low_q", "odom_q", "dq", "ddq", "tau", "contact")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(H.GOLDEN_DIR, "ingest_g1.npz")))


def same(a, b):
    """Bit-for-bit equality of float arrays, NaN == NaN, -0.0 != +0.0."""
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else a
    b = b.cpu().numpy() if isinstance(b, torch.Tensor) else b
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.int64) | (np.isnan(a) * -1), b.view(np.int64) | (np.isnan(b) * -1))


def test_load_dat_equals_loadtxt_on_the_reference_files(gold, tmp_path):
    from system_identification_b200 import ingest
    for name in FILES:
        text = gold["dat_text_" + name].tobytes()
        assert same(ingest.load_dat(text), gold["loadtxt_" + name]), name                 # bytes
        path = tmp_path / f"g1_robot_{name}.dat"
        path.write_bytes(text)
        assert same(ingest.load_dat(str(path)), gold["loadtxt_" + name]), name            # file
    # read_data composes the five loads (and the filters of row f2) like the reference's read_data
    (tmp_path / "g1_robot_q.dat").write_bytes(gold["dat_text_low_q"].tobytes())
    # the reference's ddq file has 17 rows (csv2dat.py:36): shapes pass through unchanged
    q, dq, ddq, tau, cnt = ingest.read_data(str(tmp_path) + os.sep, "g1", "none")
    assert same(q, gold["loadtxt_low_q"]) and same(ddq, gold["loadtxt_ddq"]) and same(cnt, gold["loadtxt_contact"])
    with open("/proc/self/maps") as f:
        assert "libsysid_b200.so" in f.read()


def test_load_dat_large_multiblock_and_special_values():
    """Fields straddling the 8 KB block and 32-byte slice boundaries, nan / inf / -0, both dtypes, no final newline."""
    from system_identification_b200 import ingest
    rng = np.random.default_rng(21)
    x = rng.normal(0, 300.0, (7, 5003))
    x[2, :6] = [np.nan, np.inf, -np.inf, -0.0, 1e-7, 9.0e9]
    x[5] = rng.integers(-3, 4, 5003)                                   # short fields: many per slice
    buf = io.BytesIO()
    np.savetxt(buf, x, delimiter="\t", fmt="%.6f")
    text = buf.getvalue()
    for dtype in (np.float32, np.float64):
        ref = np.loadtxt(io.BytesIO(text), delimiter="\t", dtype=dtype)
        assert same(ingest.load_dat(text, dtype=dtype), ref)
        assert same(ingest.load_dat(text[:-1], dtype=dtype), ref)      # last row without '\n'
        assert same(ingest.load_dat(text + b"\n\n", dtype=dtype), ref)   # trailing blank lines
    # other formats np.savetxt can write: short exponent forms ride the exact path, 19-digit mantissas the integer
    # division; what neither covers is converted on the host (below)
    for fmt, block in (("%.8e", x[:3, :50]), ("%.18e", np.abs(x[:2, 6:60]) + 1.0), ("%.10f", x[:2, :60]), ("%.17g", x[:2, :60])):
        buf = io.BytesIO()
        np.savetxt(buf, block, delimiter="\t", fmt=fmt)
        ref = np.loadtxt(io.BytesIO(buf.getvalue()), delimiter="\t")
        assert same(ingest.load_dat(buf.getvalue(), dtype=np.float64), ref), fmt
    # fields outside the device's exact-conversion domain (floating-point residue of a logger, 21-digit integers) are
    # flagged by the kernel and converted on the host by float() -- the value np.loadtxt gives -- only those fields
    hard = b"1.0\t1.2345678901234567e-30\t-2.7755575615628914e-17\n123456789012345678901.5\t1.5e-30\tnan\n"
    for dtype in (np.float32, np.float64):
        assert same(ingest.load_dat(hard, dtype=dtype), np.loadtxt(io.BytesIO(hard), delimiter="\t", dtype=dtype))
    with pytest.raises(ValueError, match="could not convert"):
        ingest.load_dat(b"1.0\t1.5e-30\n2.0\t1.5e-3x\n", dtype=np.float64)
    import pandas as pd
    csv = b"a,b\n1.5,-2.7755575615628914e-17\n,3e-40\n"
    got = ingest.load_csv(csv)
    ref = pd.read_csv(io.BytesIO(csv), float_precision="round_trip")
    assert same(got["a"], ref["a"].to_numpy()) and same(got["b"], ref["b"].to_numpy())


def test_load_dat_errors_like_loadtxt():
    from system_identification_b200 import ingest
    for bad in (b"1.0\t2.0\n3.0\n", b"1.0\t2.0\n3.0\t4.0\t5.0\n", b"1.0\tabc\n", b"1.0\t\t2.0\n", b"1.0\t2.0\t\n"):
        with pytest.raises(ValueError):
            np.loadtxt(io.BytesIO(bad), delimiter="\t")
        with pytest.raises(ValueError):
            ingest.load_dat(bad)
    with pytest.raises(ValueError):
        ingest.load_dat(b"\n\n")
    assert same(ingest.load_dat(b"3.5"), np.array([[3.5]]))


def test_fd_rate_contact_and_round_equal_the_reference_scripts(gold):
    import pandas as pd
    from system_identification_b200 import ingest
    df = pd.read_csv(io.BytesIO(gold["csv_text"].tobytes()))
    dq = np.stack([df[f"low_motor_{i}_dq"].to_numpy() for i in range(12)])
    gyro = np.stack([df[f"low_imu_gyro_{a}"].to_numpy() for a in "xyz"])
    assert same(ingest.fd_rate(df["low_tick"].to_numpy(), dq, 1000.0), gold["upd_ddq"])
    assert same(ingest.fd_rate(df["low_tick"].to_numpy(), gyro, 1000.0), gold["upd_body_acc"])
    assert same(ingest.fd_rate(df["timestamp"].to_numpy(), dq, 1.0), gold["plain_ddq"])
    tau = np.stack([df["low_motor_4_tau_est"].to_numpy(), df["low_motor_10_tau_est"].to_numpy()])
    assert same(ingest.contact_from_tau(tau), gold["upd_contact"])
    # the whole pipeline, file-free, against the .dat files the reference wrote and np.loadtxt read
    log = ingest.csv_to_log(df, fix_ddq_off_by_one=False)
    for name in FILES:
        assert same(log[name], gold["loadtxt_" + name]), name
    fixed = ingest.csv_to_log(df)
    assert fixed["ddq"].shape[0] == 18 and same(fixed["ddq"][7:], gold["loadtxt_ddq"][6:])
    with pytest.raises(ValueError, match="Missing columns"):
        ingest.csv_to_log({"low_tick": np.arange(4)})


def test_round_dat_equals_savetxt_loadtxt_on_ties_and_random_values():
    from system_identification_b200 import ingest
    rng = np.random.default_rng(22)
    half = (rng.integers(-4000000, 4000000, 60000) + 0.5) * 1e-6                 # binary neighbours of decimal ties
    x = np.r_[rng.normal(0, 30.0, 100000), half, np.nextafter(half, np.inf), np.nextafter(half, -np.inf),
              rng.integers(-2 ** 20, 2 ** 20, 20000) * 2.0 ** -21,                # exact binary ties of the sixth decimal? some are
              rng.normal(0, 1e-6, 20000), rng.normal(0, 4e9, 2000),
              [0.0000005, 0.0000015, 0.0000025, -0.0000005, 0.125 + 2.0 ** -21, -0.0, 0.0, -1e-9, 8.58e9, 8.6e9, 1e12, np.nan, np.inf, -np.inf]]
    buf = io.BytesIO()
    np.savetxt(buf, x[None, :], delimiter="\t", fmt="%.6f")
    for f32, dtype in ((True, np.float32), (False, np.float64)):
        ref = np.loadtxt(io.BytesIO(buf.getvalue()), delimiter="\t", dtype=dtype)
        got = ingest.round_dat(x, float32=f32).reshape(-1)
        assert same(got, ref)
    # and the text route agrees with the text-free route
    assert same(ingest.load_dat(buf.getvalue(), dtype=np.float64).reshape(-1), ingest.round_dat(x, float32=False).reshape(-1))


def test_fd_rate_vs_oracle_full_size():
    """BASELINE configs[2]-sized log (20 000 samples, 15 channels): vectorised device loop == the reference's row loop."""
    from system_identification_b200 import ingest
    rng = np.random.default_rng(23)
    N = 20000
    tick = np.cumsum(rng.integers(0, 4, N)).astype(np.float64)                     # zeros: repeated ticks
    x = np.round(rng.normal(0, 2.0, (15, N)), 3)                                   # coarse values: exact repeats happen
    assert same(ingest.fd_rate(tick, x, 1000.0), oi.fd_rate(tick, x, 1000.0))
    tau = rng.normal(0, 12.0, (2, N))
    assert same(ingest.contact_from_tau(tau), oi.contact_from_tau(tau))


def test_cache_round_trip(tmp_path):
    from system_identification_b200 import ingest
    flat, data = H.small_log("g1_12dof", 64)
    ingest.save_cache(str(tmp_path / "cache"), **dict(zip(("q", "dq", "ddq", "tau", "contact"), data)))
    back = ingest.load_cache(str(tmp_path / "cache"))
    for a, b in zip(back, data):
        assert same(a, b)


def test_g1_driver_on_dat_files_equals_loadtxt_route(tmp_path, capsys, monkeypatch):
    """BASELINE configs[2] end to end: a G1 log in the csv2dat .dat layout -> g1_identification.py (device parse + filter +
    fused identify) prints the reference's tables, and the parameters equal those identified from the same files read by
    np.loadtxt + scipy.signal.filtfilt (the reference's read_data) to solver precision."""
    import sys
    import scipy.signal as signal
    sys.path.insert(0, H.ROOT)
    import g1_identification
    from system_identification_b200 import ingest
    from system_identification_b200.identify import identify
    from system_identification_b200.sys_identification import SystemIdentification
    flat = H.flat_model("g1_12dof")
    si = SystemIdentification.from_flat_model(flat)
    dm = si.device_model

    def regress(q, dq, ddq, cnt):
        from system_identification_b200.ops import to_device
        dev = tuple(to_device(a) for a in (q, dq, ddq, np.zeros((12, q.shape[1])), cnt))
        Y = dm.regressor_batch(*dev[:3]).cpu().numpy()
        _, _, P = dm.projected_batch(*dev, want_P=True)
        return Y, P.cpu().numpy()
    data, _, _, _ = H.identifiable_log(flat, 300, 91, regress)          # the log of test_end_to_end_identify_vs_oracle
    d = str(tmp_path) + os.sep
    for key, arr in zip(("low_q", "dq", "ddq", "tau", "contact"), data):
        np.savetxt(d + f"g1_robot_{key}.dat", arr, delimiter="\t", fmt="%.6f")            # what csv2dat.py:50-55 writes
    # the reference's read_data on the same files (spot_identification.py:9-24)
    ref = [np.loadtxt(d + f"g1_robot_{k}.dat", delimiter="\t", dtype=np.float32) for k in ("low_q", "dq", "ddq", "tau", "contact")]
    b, a = signal.butter(5, 0.15, btype="low", analog=False)
    ref[1], ref[2], ref[3] = (signal.filtfilt(b, a, v, axis=1) for v in ref[1:4])
    phi_ref = identify(si, *[np.asarray(v, dtype=np.float64) for v in ref])
    # the device route
    q, dq, ddq, tau, cnt = ingest.read_data(d, "g1", "butterworth", q_name="low_q")
    assert same(q, ref[0]) and same(cnt, ref[4])
    assert np.abs(dq.cpu().numpy() - ref[1]).max() <= 1e-10 * np.abs(ref[1]).max()
    phi_dev = identify(si, q, dq, ddq, tau, cnt)
    assert H.rel(phi_dev, phi_ref) <= 1e-6
    # and the driver itself
    monkeypatch.setattr(sys, "argv", ["g1_identification.py", "--data", d])
    g1_identification.main()
    out = capsys.readouterr().out
    assert "Identified" in out and "Prior" in out


def test_load_csv_equals_exact_pandas_and_feeds_csv_to_log(gold):
    """The logger CSV parsed on the device: every column equals pandas' round-trip parse (the exact decimal -> double
    conversion; pandas' default parser differs from it by an ulp on some fields), empty fields are NaN, and csv_to_log on
    the device-parsed columns reproduces the reference pipeline's .dat arrays like the pandas route does."""
    import pandas as pd
    from system_identification_b200 import ingest
    text = gold["csv_text"].tobytes()
    cols = ingest.load_csv(text)
    ref = pd.read_csv(io.BytesIO(text), float_precision="round_trip")
    assert list(cols) == list(ref.columns)
    for name in ref.columns:
        assert same(cols[name], ref[name].to_numpy().astype(np.float64)), name
    log = ingest.csv_to_log(cols, fix_ddq_off_by_one=False)
    for name in FILES:
        assert same(log[name], gold["loadtxt_" + name]), name
    # empty fields are NaN (pandas), ragged rows and garbage raise
    small = ingest.load_csv(b"a,b,c\n1.5,,3\n4,5,6\n")
    assert same(small["b"], np.array([np.nan, 5.0])) and same(small["c"], np.array([3.0, 6.0]))
    with pytest.raises(ValueError):
        ingest.load_csv(b"a,b\n1,2\n3\n")
    with pytest.raises(ValueError, match="column 'b'"):
        ingest.load_csv(b"a,b\n1,x\n")
