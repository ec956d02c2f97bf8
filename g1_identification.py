"""Unitree G1 (12-DoF legs) identification driver.

The reference ships drivers for Solo (demo/solo_identification.py) and Spot (spot_identification.py) only; its G1 data
tooling (g1-data/csv2dat.py) writes `g1_robot_{low_q,odom_q,dq,ddq,tau,contact}.dat` but nothing reads them (SURVEY 8,
"reference gaps").  This script mirrors spot_identification.py's main() step for step for that robot:

    read_data            ingest.read_data: the five .dat files parsed ON THE DEVICE (np.loadtxt semantics, float32),
                         or ingest.csv_to_log straight from the logger CSV, then the Butterworth / Savitzky-Golay filter
    regressor + solve    identify(): fused regressor + projector + Gram kernel, LMI fit on the device
    printers             print_inertial_params, print_tau_prediction_rmse (reference formulas)

    python g1_identification.py --data DIR/ [--q low_q|odom_q] [--filter butterworth|savitzky|none]
    python g1_identification.py --csv run.csv
"""
import argparse
import os

from src.sys_identification import SystemIdentification
from system_identification_b200 import filters, ingest
from system_identification_b200.identify import identify
from system_identification_b200.model import FlatModel

HERE = os.path.dirname(os.path.realpath(__file__))


def read_csv_log(csv_path, q_name, filter_type):
    """Logger CSV -> (q, dq, ddq, tau, contact) on the device: the reference's low_ddq_contact_tick.py + csv2dat.py +
    read_data composed without intermediate files (all 12 motors in ddq: csv2dat.py:36 drops motor 0)."""
    log = ingest.csv_to_log(ingest.load_csv(csv_path))         # the CSV text is parsed on the device (no pandas)
    q, dq, ddq, tau, cnt = log[q_name], log["dq"], log["ddq"], log["tau"], log["contact"]
    # the first row of the finite differences is NaN (low_ddq_contact_tick.py:38-43 leaves it so): drop that sample
    q, dq, ddq, tau, cnt = (a[:, 1:].contiguous() for a in (q, dq, ddq, tau, cnt))
    if filter_type == "butterworth":
        b, a = filters.butter_lowpass(5, 0.15)
        dq, ddq, tau = (filters.filtfilt(b, a, v, float32_input=True) for v in (dq, ddq, tau))
    elif filter_type == "savitzky":
        dq, ddq, tau = (filters.savgol_filter(v, 21, 5, float32_input=True) for v in (dq, ddq, tau))
    return q, dq, ddq, tau, cnt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", help="directory holding g1_robot_*.dat (trailing separator optional)")
    ap.add_argument("--csv", help="logger CSV (alternative to --data)")
    ap.add_argument("--q", default="low_q", choices=("low_q", "odom_q"), help="IMU or odometry quaternion variant")
    ap.add_argument("--filter", default="butterworth", choices=("butterworth", "savitzky", "none"))
    ap.add_argument("--urdf", help="g1_12dof.urdf (with --config); default: the committed flattened descriptor")
    ap.add_argument("--config", help="robot YAML (link_names, end_effectors_frame_names, mass)")
    args = ap.parse_args()
    if bool(args.data) == bool(args.csv):
        ap.error("give exactly one of --data and --csv")

    # Read the data
    robot_name = "g1"
    if args.csv:
        q, dq, ddq, tau, cnt = read_csv_log(args.csv, args.q, args.filter)
    else:
        q, dq, ddq, tau, cnt = ingest.read_data(os.path.join(args.data, ""), robot_name, args.filter, q_name=args.q)

    # Instantiate the identification problem
    if args.urdf:
        sys_idnt = SystemIdentification(args.urdf, args.config, floating_base=True)
    else:
        flat = FlatModel.load(os.path.join(HERE, "system_identification_b200", "robots", "g1_12dof.json"))
        sys_idnt = SystemIdentification.from_flat_model(flat)
    phi_prior = sys_idnt.get_phi_prior()

    # Regressor, normal equations and the LMI-constrained fit, all on the device
    phi_identified = identify(sys_idnt, q, dq, ddq, tau, cnt)
    sys_idnt.print_inertial_params(phi_prior, phi_identified)
    sys_idnt.print_tau_prediction_rmse(q, dq, ddq, tau, cnt, phi_prior, "Prior")
    sys_idnt.print_tau_prediction_rmse(q, dq, ddq, tau, cnt, phi_identified, "Identified")


if __name__ == "__main__":
    main()
