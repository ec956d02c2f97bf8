"""Log ingest on the device: from the reference's text files to the channel-major fp64 arrays the fused
kernel consumes, without np.loadtxt and without the per-row pandas loops.

    q = load_dat(path + "g1_robot_low_q.dat")              # np.loadtxt(..., delimiter='\\t', dtype=np.float32)
    q, dq, ddq, tau, cnt = read_data(path, "spot", "butterworth")   # reference spot_identification.py:9-33, on the device
    log = csv_to_log(pd.read_csv("run.csv"))                # low_ddq_contact_tick.py + csv2dat.py + read_data, no files
    log = csv_to_log(load_csv("run.csv"))                   # ... and without pandas: the CSV text parsed on the device

The arithmetic is done by libsysid_b200.so (csrc/ingest_kernels.cuh): sysid_dat_scan / sysid_dat_parse, sysid_fd_rate,
sysid_contact_from_tau, sysid_round_dat.  No CPU fallback: without the library or a CUDA device these calls raise.
Results are bit-identical to the reference's scripts (tests/golden/ingest_g1.npz was produced by running them).
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _require_cuda, _stream

# column lists of the reference's csv2dat.py (g1-data/csv2dat.py:18-40)
G1_MOTORS = 12
LOW_Q_COLS = ["odom_position_x", "odom_position_y", "odom_position_z",
              "low_imu_quat_x", "low_imu_quat_y", "low_imu_quat_z", "low_imu_quat_w"] + [f"low_motor_{i}_q" for i in range(G1_MOTORS)]
ODOM_Q_COLS = ["odom_position_x", "odom_position_y", "odom_position_z",
               "odom_imu_quaternion_x", "odom_imu_quaternion_y", "odom_imu_quaternion_z", "odom_imu_quaternion_w"] + \
              [f"low_motor_{i}_q" for i in range(G1_MOTORS)]
DQ_COLS = ["odom_velocity_x", "odom_velocity_y", "odom_velocity_z",
           "low_imu_gyro_x", "low_imu_gyro_y", "low_imu_gyro_z"] + [f"low_motor_{i}_dq" for i in range(G1_MOTORS)]
TAU_COLS = [f"low_motor_{i}_tau_est" for i in range(G1_MOTORS)]
GYRO_COLS = ["low_imu_gyro_x", "low_imu_gyro_y", "low_imu_gyro_z"]
ACCEL_COLS = ["low_imu_accel_x", "low_imu_accel_y", "low_imu_accel_z"]
CONTACT_TAU_COLS = ["low_motor_4_tau_est", "low_motor_10_tau_est"]      # low_ddq_contact_tick.py:72-81


def _dev(device=None):
    _require_cuda()
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _as_dev2d(x, device=None):
    """(channels, N) fp64 CUDA tensor with unit inner stride from a numpy array or tensor (values widened exactly)."""
    if isinstance(x, torch.Tensor):
        t = x.to(device=_dev(device), dtype=torch.float64)
    else:
        t = torch.from_numpy(np.array(x, dtype=np.float64, order="C")).to(_dev(device))     # copy: exact widening
    if t.dim() == 1:
        t = t.unsqueeze(0)
    return t if t.stride(1) == 1 else t.contiguous()


def _patch_unconverted(out, raw, delim, rows, cols, transposed, f32, empty_nan, where="", names=None):
    """The device converts a field only when it can do so EXACTLY (Clinger's fast path and a 128-bit integer division:
    every '%.6f' / '%.17g' / '%.18e' field of ordinary magnitude).  Fields outside that domain -- e.g. floating-point
    residue such as -2.7755575615628914e-17 in a logger CSV -- are flagged by the kernel (written as NaN and counted) and
    converted here, on the host, by Python's correctly rounded float(), which is what np.loadtxt / pandas would have used:
    only those few fields, patched into the device array.  A field that is not a number raises ValueError, like np.loadtxt."""
    cand = torch.isnan(out).nonzero().cpu().numpy()                   # flagged fields (and literal nan fields: re-read, unchanged)
    cuts = np.flatnonzero((raw == delim) | (raw == 0x0A))
    vals = np.empty(len(cand), dtype=np.float64)
    for n_, (a, b) in enumerate(cand):
        row, col = (int(b), int(a)) if transposed else (int(a), int(b))
        k = row * cols + col
        lo = int(cuts[k - 1]) + 1 if k > 0 else 0
        hi = int(cuts[k]) if k < cuts.size else raw.size
        tok = bytes(raw[lo:hi]).strip()
        if tok == b"" and empty_nan:
            vals[n_] = np.nan
            continue
        try:
            vals[n_] = float(tok)
        except ValueError:
            cname = f" (column {names[col]!r})" if names is not None else ""
            raise ValueError(f"could not convert string {tok!r} to float at row {row}, column {col}{cname}{where}") from None
    if f32:
        vals = vals.astype(np.float32).astype(np.float64)
    if len(cand):
        ij = torch.from_numpy(cand).to(out.device)
        out[ij[:, 0], ij[:, 1]] = torch.from_numpy(vals).to(out.device)
    return out


# ------------------------------------------------------------------------------------------------ .dat text
def load_dat(source, delimiter="\t", dtype=np.float32, device=None):
    """np.loadtxt(source, delimiter=delimiter, dtype=dtype) for the text np.savetxt(fmt='%.6f') writes, parsed on the
    device.  source: a path, or the file's bytes.  Returns a CUDA float64 tensor (rows, cols); with dtype=np.float32
    (what the reference's read_data asks for) every value is rounded through float32 and widened exactly.
    Raises ValueError where np.loadtxt would (ragged rows, fields that are not numbers)."""
    lib = _lib.load()
    dev = _dev(device)
    if dtype not in (np.float32, np.float64, "float32", "float64", torch.float32, torch.float64):
        raise ValueError("dtype must be float32 or float64")
    f32 = dtype in (np.float32, "float32", torch.float32)
    if isinstance(source, (bytes, bytearray, memoryview)):
        raw = np.frombuffer(source, dtype=np.uint8)
    else:
        raw = np.fromfile(os.fspath(source), dtype=np.uint8)
    n = raw.size
    while n > 0 and raw[n - 1] in (0x0A, 0x0D, 0x20):          # trailing blank lines (np.loadtxt skips them)
        n -= 1
    if n == 0:
        raise ValueError("input contained no data")
    n += 1 if n < raw.size and raw[n] == 0x0A else 0                 # keep the newline that ends the last row
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                  # a bytes object is read-only; it is only read
        host = torch.from_numpy(raw[:n])
    text = torch.empty(n + 16, dtype=torch.uint8, device=dev)
    text[:n].copy_(host)
    ws = torch.empty(lib.sysid_dat_workspace_bytes(n), dtype=torch.uint8, device=dev)
    dims = (C.c_int64 * 2)()
    rc = lib.sysid_dat_scan(_ptr(text), n, ord(delimiter), _ptr(ws), ws.numel(), dims, _stream())
    if rc == -1 and dims[0] > 0:
        raise ValueError(lib.sysid_last_error().decode())
    _lib.check(rc)
    rows, cols = int(dims[0]), int(dims[1])
    out = torch.empty((rows, cols), dtype=torch.float64, device=dev)
    info = (C.c_int64 * 4)()
    rc = lib.sysid_dat_parse(_ptr(text), n, ord(delimiter), _ptr(ws), ws.numel(), rows, cols, _ptr(out), cols, 1 if f32 else 0,
                             info, _stream())
    if rc == -1 and info[0] > 0 and info[2] == 0:
        # fields outside the device's exact-conversion domain: converted on the host, only those (see _patch_unconverted)
        where = "" if isinstance(source, (bytes, bytearray, memoryview)) else f" in {os.fspath(source)}"
        return _patch_unconverted(out, raw[:n], ord(delimiter), rows, cols, False, f32, False, where)
    if rc == -1:
        raise ValueError(lib.sysid_last_error().decode())
    _lib.check(rc)
    return out


def load_csv(source, device=None):
    """pd.read_csv(source) for an all-numeric logger CSV (reference g1-data/csv2dat.py:15, low_ddq_contact_tick.py:21),
    parsed on the device: returns {column name: CUDA float64 (N,) tensor} (rows of one channel-major (columns, N) tensor, so
    csv_to_log takes it as it is).  Every field is converted exactly (strtod's value; pandas' default parser is off by an
    ulp on ~15 % of 17-digit fields); an empty field is NaN, as in pandas.  A 10^6-row log is ~2 GB of text: seconds of
    pandas, tens of milliseconds here plus the upload."""
    lib = _lib.load()
    dev = _dev(device)
    if isinstance(source, (bytes, bytearray, memoryview)):
        raw = np.frombuffer(source, dtype=np.uint8)
    else:
        raw = np.fromfile(os.fspath(source), dtype=np.uint8)
    nl = np.flatnonzero(raw[:1 << 20] == 0x0A)
    if nl.size == 0:
        raise ValueError("no header line")
    names = [c.strip().strip('"') for c in bytes(raw[:nl[0]]).decode("utf-8").rstrip("\r").split(",")]
    body = raw[nl[0] + 1:]
    n = body.size
    while n > 0 and body[n - 1] in (0x0A, 0x0D, 0x20):
        n -= 1
    if n == 0:
        raise ValueError("input contained no data")
    n += 1 if n < body.size and body[n] == 0x0A else 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        host = torch.from_numpy(body[:n])
    text = torch.empty(n + 16, dtype=torch.uint8, device=dev)
    text[:n].copy_(host)
    ws = torch.empty(lib.sysid_dat_workspace_bytes(n), dtype=torch.uint8, device=dev)
    dims = (C.c_int64 * 2)()
    rc = lib.sysid_dat_scan(_ptr(text), n, ord(","), _ptr(ws), ws.numel(), dims, _stream())
    if rc == -1 and dims[0] > 0:
        raise ValueError(lib.sysid_last_error().decode())
    _lib.check(rc)
    rows, cols = int(dims[0]), int(dims[1])
    if cols != len(names):
        raise ValueError(f"header names {len(names)} columns, the rows hold {cols}")
    out = torch.empty((cols, rows), dtype=torch.float64, device=dev)
    info = (C.c_int64 * 4)()
    rc = lib.sysid_dat_parse_ex(_ptr(text), n, ord(","), _ptr(ws), ws.numel(), rows, cols, _ptr(out), rows, 2 | 4, info, _stream())
    if rc == -1 and info[0] > 0 and info[2] == 0:
        _patch_unconverted(out, body[:n], ord(","), rows, cols, True, False, True, names=names)
        return {name: out[i] for i, name in enumerate(names)}
    if rc == -1:
        raise ValueError(lib.sysid_last_error().decode())
    _lib.check(rc)
    return {name: out[i] for i, name in enumerate(names)}


def read_data(path, robot_name, filter_type, q_name="q"):
    """The reference's read_data (spot_identification.py:9-33; demo/solo_identification.py:9-33) on the device: the five
    .dat files parsed by load_dat as float32, then the same filter.  q_name selects `<robot>_robot_<q_name>.dat`
    (csv2dat writes low_q and odom_q for the G1: g1-data/csv2dat.py:50-51).  Returns (q, dq, ddq, tau, contact) as CUDA
    float64 (channels, N) tensors."""
    from . import filters
    q, dq, ddq, tau, cnt = (load_dat(path + robot_name + f"_robot_{k}.dat") for k in (q_name, "dq", "ddq", "tau", "contact"))
    if filter_type == "butterworth":
        b, a = filters.butter_lowpass(5, 0.15)
        dq, ddq, tau = (filters.filtfilt(b, a, v, float32_input=True) for v in (dq, ddq, tau))
    elif filter_type == "savitzky":
        dq, ddq, tau = (filters.savgol_filter(v, 21, 5, float32_input=True) for v in (dq, ddq, tau))
    return q, dq, ddq, tau, cnt


# ------------------------------------------------------------------------------------------------ CSV post-processing
def fd_rate(tick, x, scale=1000.0):
    """The row loop of calculate_low_motor_ddq for all channels at once (g1-data/low_ddq_contact_tick.py:46-70):
    y[:, 0] = nan; y[:, i] = (x[:, i] - x[:, i-1]) * scale / (tick[i] - tick[i-1]) when the tick advanced, 0 when neither
    moved, nan otherwise.  scale=1000: millisecond ticks; scale=1: low_ddq.py (seconds)."""
    lib = _lib.load()
    x = _as_dev2d(x)
    t = _as_dev2d(tick, x.device).reshape(-1)
    ch, N = x.shape
    if t.numel() != N:
        raise ValueError("tick and x disagree on the number of samples")
    y = torch.empty((ch, N), dtype=torch.float64, device=x.device)
    _lib.check(lib.sysid_fd_rate(_ptr(t), _ptr(x), _ptr(y), ch, N, x.stride(0) if ch > 1 else N, N, float(scale), _stream()))
    return y


def contact_from_tau(tau, hi=10.0, lo=-5.0):
    """np.where(tau >= 10, 1, np.where(tau > -5, 2, 0)) (g1-data/low_ddq_contact_tick.py:72-81), any shape."""
    lib = _lib.load()
    one_d = (tau.dim() if isinstance(tau, torch.Tensor) else np.ndim(tau)) == 1
    t = _as_dev2d(tau).contiguous()
    out = torch.empty_like(t)
    _lib.check(lib.sysid_contact_from_tau(_ptr(t), _ptr(out), t.numel(), float(hi), float(lo), _stream()))
    return out.reshape(-1) if one_d else out


def round_dat(x, float32=True):
    """What np.savetxt(fmt='%.6f') followed by np.loadtxt(dtype=np.float32 if float32 else float) does to the values of x
    (g1-data/csv2dat.py:50-55 then read_data), without the text."""
    lib = _lib.load()
    x = _as_dev2d(x)
    ch, N = x.shape
    y = torch.empty((ch, N), dtype=torch.float64, device=x.device)
    _lib.check(lib.sysid_round_dat(_ptr(x), _ptr(y), ch, N, x.stride(0) if ch > 1 else N, N, 1 if float32 else 0, _stream()))
    return y


def _rows(columns, names):
    cols = [columns[c] for c in names]
    if all(isinstance(c, torch.Tensor) for c in cols):
        return torch.stack([c.to(dtype=torch.float64) for c in cols])          # load_csv output: already on the device
    return np.stack([np.asarray(c.cpu() if isinstance(c, torch.Tensor) else c, dtype=np.float64) for c in cols])


def csv_to_log(columns, tick_col="low_tick", scale=1000.0, relabel_contact=True, fix_ddq_off_by_one=True, float32=True):
    """A logger CSV (a pandas DataFrame, or any mapping column name -> 1-D array) -> the arrays read_data would load
    after the reference's two post-processing scripts, with no intermediate file:

      low_ddq_contact_tick.calculate_low_motor_ddq   joint / body angular accelerations by finite differences of the
                                                     tick, contact labels from the ankle torques (relabel_contact)
      csv2dat.main                                   column selection and '%.6f' text
      read_data                                      np.loadtxt(dtype=np.float32)

    fix_ddq_off_by_one: csv2dat.py:36 lists low_motor_{1..11}_ddq (motor 0 is dropped: a 17-row ddq no regressor accepts);
    True (default) emits all 12 motors, False reproduces the reference's 17 rows.
    Returns a dict of CUDA float64 (channels, N) tensors: low_q, odom_q, dq, ddq, tau, contact."""
    missing = [c for c in set(LOW_Q_COLS + ODOM_Q_COLS + DQ_COLS + TAU_COLS + ACCEL_COLS + [tick_col]) if c not in columns]
    if missing:
        raise ValueError(f"Missing columns in CSV: {sorted(missing)}")
    tick = columns[tick_col] if isinstance(columns[tick_col], torch.Tensor) else np.asarray(columns[tick_col], dtype=np.float64)
    motors = range(G1_MOTORS) if fix_ddq_off_by_one else range(1, G1_MOTORS)
    rates = fd_rate(tick, _rows(columns, GYRO_COLS + [f"low_motor_{i}_dq" for i in motors]), scale)
    ddq = torch.cat([_as_dev2d(_rows(columns, ACCEL_COLS)), rates], dim=0)
    if relabel_contact:
        contact = contact_from_tau(_rows(columns, CONTACT_TAU_COLS))
    else:
        contact = _as_dev2d(_rows(columns, ["odom_foot_contact_1", "odom_foot_contact_2"]))
    out = {"low_q": _as_dev2d(_rows(columns, LOW_Q_COLS)), "odom_q": _as_dev2d(_rows(columns, ODOM_Q_COLS)),
           "dq": _as_dev2d(_rows(columns, DQ_COLS)), "ddq": ddq, "tau": _as_dev2d(_rows(columns, TAU_COLS)), "contact": contact}
    return {k: round_dat(v, float32=float32) for k, v in out.items()}


# ------------------------------------------------------------------------------------------------ columnar binary cache
def save_cache(path, **arrays):
    """Columnar binary cache of a parsed log: one little-endian fp64 .npy per array under `path` (a directory), so that
    the next run memory-maps the values instead of parsing 10^8 characters of text."""
    os.makedirs(path, exist_ok=True)
    for k, v in arrays.items():
        a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        np.save(os.path.join(path, k + ".npy"), np.ascontiguousarray(a, dtype="<f8"))


def load_cache(path, names=("q", "dq", "ddq", "tau", "contact"), device=None):
    """Memory-map the cache written by save_cache and upload it.  Returns a tuple of CUDA float64 tensors."""
    dev = _dev(device)
    out = []
    for k in names:
        a = np.load(os.path.join(path, k + ".npy"), mmap_mode="r")
        out.append(torch.from_numpy(np.array(a, dtype=np.float64, order="C")).to(dev))
    return tuple(out)
