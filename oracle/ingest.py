"""TEST INFRASTRUCTURE -- CPU restatement of the reference's log ingest.  Only tests/, smoke() and
bench.py's CPU legs may import this; the product path never does.

Unlike stages 1-3, this row's reference code RUNS in the build container (pandas + numpy are installed), so the
restatement is pinned against the reference itself: tests/golden/make_ingest_golden.py executes
/root/reference/g1-data/low_ddq_contact_tick.py and csv2dat.py on a synthetic logger CSV and commits their outputs
(tests/golden/ingest_g1.npz); tests/test_oracle_ingest.py checks every function below against that fixture and against
np.loadtxt / np.savetxt called directly.

Plain Python loops on purpose (small cases only): each function follows the reference line by line.
"""
from __future__ import annotations

import numpy as np

MOTORS = 12
# g1-data/csv2dat.py:18-40
LOW_Q_COLS = ['odom_position_x', 'odom_position_y', 'odom_position_z',
              'low_imu_quat_x', 'low_imu_quat_y', 'low_imu_quat_z', 'low_imu_quat_w'] + [f'low_motor_{i}_q' for i in range(12)]
ODOM_Q_COLS = ['odom_position_x', 'odom_position_y', 'odom_position_z',
               'odom_imu_quaternion_x', 'odom_imu_quaternion_y', 'odom_imu_quaternion_z', 'odom_imu_quaternion_w'] + \
              [f'low_motor_{i}_q' for i in range(12)]
DQ_COLS = ['odom_velocity_x', 'odom_velocity_y', 'odom_velocity_z',
           'low_imu_gyro_x', 'low_imu_gyro_y', 'low_imu_gyro_z'] + [f'low_motor_{i}_dq' for i in range(12)]
TAU_COLS = [f'low_motor_{i}_tau_est' for i in range(12)]
CONTACT_COLS = ['odom_foot_contact_1', 'odom_foot_contact_2']


def ddq_cols(fix_off_by_one):
    """g1-data/csv2dat.py:33-36: the reference lists range(1, 12) -- motor 0 is missing (17 rows)."""
    first = 0 if fix_off_by_one else 1
    return ['low_imu_accel_x', 'low_imu_accel_y', 'low_imu_accel_z',
            'body_ang_acceleration_x', 'body_ang_acceleration_y', 'body_ang_acceleration_z'] + \
           [f'low_motor_{i}_ddq' for i in range(first, 12)]


def parse_dat_text(text, delimiter="\t", dtype=np.float32):
    """np.loadtxt(file, delimiter=delimiter, dtype=dtype) (spot_identification.py:10-14): every field through Python's
    float() (correctly rounded, like the strtod numpy's reader calls), then cast to dtype.  Blank lines are skipped;
    ragged rows raise ValueError."""
    if isinstance(text, (bytes, bytearray)):
        text = bytes(text).decode("ascii")
    rows = []
    for line in text.split("\n"):
        if line.strip() == "":
            continue
        rows.append([float(f) for f in line.split(delimiter)])
    if not rows:
        raise ValueError("input contained no data")
    if any(len(r) != len(rows[0]) for r in rows):
        raise ValueError("the number of columns changed between rows")
    return np.array(rows, dtype=np.float64).astype(dtype)


def fd_rate(tick, x, scale=1000.0):
    """g1-data/low_ddq_contact_tick.py:46-70 (scale 1000), low_ddq.py:19-33 (no scale == scale 1): for every channel
    y[0] = nan, then per row the reference's three branches in the reference's order."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    tick = np.asarray(tick)
    y = np.full(x.shape, np.nan)
    for ch in range(x.shape[0]):
        for row in range(1, x.shape[1]):
            delta_time = tick[row] - tick[row - 1]
            delta = x[ch, row] - x[ch, row - 1]
            if delta_time > 0:
                y[ch, row] = delta * scale / delta_time
            elif delta == 0:
                y[ch, row] = 0.0
            else:
                y[ch, row] = np.nan
    return y


def contact_from_tau(tau, hi=10.0, lo=-5.0):
    """g1-data/low_ddq_contact_tick.py:72-81."""
    tau = np.asarray(tau, dtype=np.float64)
    return np.where(tau >= hi, 1, np.where(tau > lo, 2, 0)).astype(np.float64)


def round_dat(x, float32=True):
    """np.savetxt(fmt='%.6f') (g1-data/csv2dat.py:50-55) then np.loadtxt(dtype=float32) (read_data), value by value."""
    x = np.asarray(x, dtype=np.float64)
    out = np.array([float("%.6f" % v) for v in x.reshape(-1)], dtype=np.float64).reshape(x.shape)
    return out.astype(np.float32).astype(np.float64) if float32 else out


def csv_to_log(columns, tick_col="low_tick", scale=1000.0, relabel_contact=True, fix_ddq_off_by_one=True, float32=True):
    """low_ddq_contact_tick.calculate_low_motor_ddq, csv2dat.main and read_data's loadtxt composed, without files.
    columns: mapping name -> 1-D array (a DataFrame works).  Returns dict of (channels, N) float64 arrays."""
    col = {k: np.asarray(columns[k], dtype=np.float64) for k in columns if k != tick_col}
    tick = np.asarray(columns[tick_col])
    for i in range(MOTORS):
        col[f'low_motor_{i}_ddq'] = fd_rate(tick, col[f'low_motor_{i}_dq'], scale)[0]
    for ax in "xyz":
        col[f'body_ang_acceleration_{ax}'] = fd_rate(tick, col[f'low_imu_gyro_{ax}'], scale)[0]
    if relabel_contact:
        col['odom_foot_contact_1'] = contact_from_tau(col['low_motor_4_tau_est'])
        col['odom_foot_contact_2'] = contact_from_tau(col['low_motor_10_tau_est'])
    sets = {"low_q": LOW_Q_COLS, "odom_q": ODOM_Q_COLS, "dq": DQ_COLS, "ddq": ddq_cols(fix_ddq_off_by_one),
            "tau": TAU_COLS, "contact": CONTACT_COLS}
    return {k: round_dat(np.stack([col[c] for c in names]), float32) for k, names in sets.items()}
