"""Shared test helpers.  Tests are the only place (with smoke() and bench.py's baseline legs) that may import oracle/."""
import os

import numpy as np

from oracle import dynamics as dy
from oracle import urdf_tree as ut
from system_identification_b200 import synth
from system_identification_b200.model import FlatModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROBOTS_DIR = os.path.join(ROOT, "system_identification_b200", "robots")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REFERENCE_FILES = "/root/reference/files"
ROBOTS = ("solo12", "spot", "g1_12dof")
URDFS = {"solo12": ("solo_description/solo12.urdf", "solo_description/solo12_config.yaml"),
         "spot": ("spot_description/spot.urdf", "spot_description/spot_config.yaml"),
         "g1_12dof": ("g1_description/g1_12dof.urdf", None)}


def flat_model(name) -> FlatModel:
    return FlatModel.load(os.path.join(ROBOTS_DIR, name + ".json"))


def oracle_tree(flat):
    return ut.tree_from_flat(flat)


def small_log(name, N, seed=None, tau_seed=3):
    m = flat_model(name)
    q, dq, ddq, cnt = synth.make_trajectory(m, N, synth.SEEDS[name] if seed is None else seed)
    tau = synth.synth_tau(m, N, tau_seed)
    return m, (q, dq, ddq, tau, cnt)


def oracle_blocks(flat, data):
    """Per-sample oracle outputs: Y (N,nv,p), P (N,nv,nv), stacked A (N*nv,c), b (N*nv)."""
    t = oracle_tree(flat)
    q, dq, ddq, tau, cnt = data
    N = q.shape[1]
    Y = np.array([dy.joint_torque_regressor(t, q[:, i], dq[:, i], ddq[:, i]) for i in range(N)])
    P = np.array([dy.null_space_projector(t, q[:, i], cnt[:, i], flat.ee_names) for i in range(N)])
    A, b = dy.stacked_system(t, q, dq, ddq, tau, cnt, flat.ee_names)
    return Y, P, A, b


def split_stats(stats, c):
    stats = np.asarray(stats)
    return stats[:c * c].reshape(c, c), stats[c * c:c * c + c], float(stats[c * c + c]), float(stats[c * c + c + 1])


def identifiable_log(flat, N, seed, regress_fn, noise_scale=1.0):
    """Log whose torques come from a perturbed ground truth: regress_fn(q,dq,ddq,cnt) -> (Y (N,nv,p), P (N,nv,nv))."""
    q, dq, ddq, cnt = synth.make_trajectory(flat, N, seed)
    Y, P = regress_fn(q, dq, ddq, cnt)
    rng = np.random.default_rng(seed + 5)
    phi_true = flat.body_params[1:].reshape(-1) * (1 + 0.15 * rng.standard_normal(10 * flat.nbodies))
    sc = 1.0 if flat.name.startswith("solo") else 10.0
    bv = rng.uniform(0, 0.02, flat.joints_dof) * sc
    bc = rng.uniform(0, 0.05, flat.joints_dof) * sc
    tau = synth.torques_from_truth(flat, Y, P, dq, phi_true, bv, bc, 0.05 * sc * noise_scale, seed)
    return (q, dq, ddq, tau, cnt), phi_true, bv, bc


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))
