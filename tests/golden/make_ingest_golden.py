"""Generates tests/golden/ingest_g1.npz by RUNNING THE REFERENCE'S OWN SCRIPTS (they need only pandas + numpy, which this
container has; /root/reference does not exist on the GPU box, hence the committed fixture):

    /root/reference/g1-data/low_ddq_contact_tick.py   calculate_low_motor_ddq(csv, motor_count=35)  -> *_updated_tick.csv
    /root/reference/g1-data/low_ddq.py                calculate_low_motor_ddq(csv)  ('timestamp' column, no x1000)
    /root/reference/g1-data/csv2dat.py                main()                                        -> six .dat files
    np.loadtxt(..., delimiter='\\t', dtype=np.float32)                                               (read_data)

on a small synthetic logger CSV that exercises every branch (repeated ticks with and without a change of the signal,
a tick that runs backwards, ankle torques on both sides of both contact thresholds, values on '%.6f' rounding ties).

    python tests/golden/make_ingest_golden.py
"""
import importlib.util
import warnings
import io
import os
import sys
import tempfile

import numpy as np
import pandas as pd

REF = "/root/reference/g1-data"
HERE = os.path.dirname(os.path.abspath(__file__))
N = 48
MOTORS_IN_CSV = 35          # the reference script's __main__ passes motor_count=35 (g1-data/low_ddq_contact_tick.py:113)


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_csv(rng):
    tick = np.cumsum(rng.integers(1, 4, size=N)).astype(np.int64) + 1000
    tick[7] = tick[6]            # tick did not advance, signal changed      -> nan
    tick[15] = tick[14]          # tick did not advance, signal did not      -> 0.0 (dq forced equal below)
    tick[23] = tick[22] - 1      # tick ran backwards                        -> nan
    cols = {"low_tick": tick, "timestamp": tick.astype(np.float64) * 1e-3 + 1.7e9}
    t = np.arange(N) * 0.002
    for k, ax in enumerate("xyz"):
        cols[f"odom_position_{ax}"] = 0.1 * np.sin(3 * t + k) + rng.normal(0, 1e-3, N)
        cols[f"odom_velocity_{ax}"] = 0.3 * np.cos(3 * t + k) + rng.normal(0, 1e-3, N)
        cols[f"low_imu_gyro_{ax}"] = 0.5 * np.sin(5 * t + k) + rng.normal(0, 1e-2, N)
        cols[f"low_imu_accel_{ax}"] = (9.81 if ax == "z" else 0.0) + rng.normal(0, 5e-2, N)
    for pre in ("low_imu_quat", "odom_imu_quaternion"):
        qv = rng.normal(0, 0.1, (N, 4)) + np.array([0, 0, 0, 1.0])
        qv /= np.linalg.norm(qv, axis=1, keepdims=True)
        for k, ax in enumerate("xyzw"):
            cols[f"{pre}_{ax}"] = qv[:, k]
    for i in range(MOTORS_IN_CSV):
        cols[f"low_motor_{i}_q"] = 0.4 * np.sin(2 * t + i) + rng.normal(0, 1e-3, N)
        dq = 0.8 * np.cos(2 * t + i) + rng.normal(0, 1e-2, N)
        dq[15] = dq[14]
        cols[f"low_motor_{i}_dq"] = dq
        cols[f"low_motor_{i}_tau_est"] = rng.normal(0, 12.0, N)
    # both contact thresholds, exactly on and just off them
    cols["low_motor_4_tau_est"][:6] = [10.0, 9.999999, -5.0, -4.999999, 25.0, -30.0]
    cols["low_motor_10_tau_est"][:6] = [-5.0, 10.0, 0.0, np.nextafter(10.0, 0), np.nextafter(-5.0, 0), 11.0]
    # '%.6f' ties and near-ties (binary values just above / below / on a half unit of the sixth decimal)
    cols["low_motor_0_q"][:8] = [0.0000005, 0.0000015, 0.0000025, -0.0000005, 0.5000005, 1.0000015, -0.0000004, 2.5e-7]
    cols["low_motor_1_q"][:4] = [0.125 + 2.0 ** -21, 0.1234565, 1234.5678905, -0.9999995]
    cols["odom_foot_contact_1"] = rng.integers(0, 2, N)
    cols["odom_foot_contact_2"] = rng.integers(0, 2, N)
    return pd.DataFrame(cols)


def main():
    warnings.simplefilter("ignore")          # the reference's column-by-column inserts trip a pandas PerformanceWarning
    rng = np.random.default_rng(3003)
    df = synthetic_csv(rng)
    tick_mod = _load("low_ddq_contact_tick")
    plain_mod = _load("low_ddq")
    csv2dat = _load("csv2dat")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        csv_path = os.path.join(tmp, "log.csv")
        df.to_csv(csv_path, index=False)
        out["csv_text"] = np.frombuffer(open(csv_path, "rb").read(), dtype=np.uint8)
        stdout = sys.stdout
        sys.stdout = io.StringIO()
        try:
            tick_mod.calculate_low_motor_ddq(csv_path, motor_count=MOTORS_IN_CSV)
            plain_mod.calculate_low_motor_ddq(csv_path, motor_count=MOTORS_IN_CSV)
            cwd = os.getcwd()
            os.chdir(tmp)
            argv = sys.argv
            sys.argv = ["csv2dat.py", os.path.join(tmp, "log_updated_tick.csv")]
            try:
                csv2dat.main()
            finally:
                sys.argv = argv
                os.chdir(cwd)
        finally:
            sys.stdout = stdout
        # the reference's intermediate results (float64 columns of the updated CSVs).  to_csv writes shortest round-trip
        # digits; float_precision='round_trip' reads them back exactly (pandas' default fast parser is off by an ulp on
        # ~15 % of the fields -- csv2dat.py itself reads with the default parser, which the .dat files below include)
        upd = pd.read_csv(os.path.join(tmp, "log_updated_tick.csv"), float_precision="round_trip")
        plain = pd.read_csv(os.path.join(tmp, "log_updated.csv"), float_precision="round_trip")
        out["upd_ddq"] = np.stack([upd[f"low_motor_{i}_ddq"].to_numpy() for i in range(12)])
        out["upd_body_acc"] = np.stack([upd[f"body_ang_acceleration_{a}"].to_numpy() for a in "xyz"])
        out["upd_contact"] = np.stack([upd["odom_foot_contact_1"].to_numpy(), upd["odom_foot_contact_2"].to_numpy()]).astype(np.float64)
        out["plain_ddq"] = np.stack([plain[f"low_motor_{i}_ddq"].to_numpy() for i in range(12)])
        for name in ("low_q", "odom_q", "dq", "ddq", "tau", "contact"):
            path = os.path.join(tmp, f"g1_robot_{name}.dat")
            out["dat_text_" + name] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
            out["loadtxt_" + name] = np.loadtxt(path, delimiter="\t", dtype=np.float32)      # read_data
    np.savez_compressed(os.path.join(HERE, "ingest_g1.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
