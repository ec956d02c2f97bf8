"""Synthetic workspace for the reference's driver scripts (test / timing infrastructure).

The reference ships no data (data/ is git-ignored) and its scripts find everything relative to their own location
(reference demo/solo_identification.py:58-65, spot_identification.py:59-66):

    demo/solo_identification.py    <ws>/data/solo/solo_robot_{q,dq,ddq,tau,contact}.dat   <ws>/files/solo_description/...
    demo/spot_identification.py    <ws>/data/spot_robot_*.dat                             <ws>/files/spot_description/...
    spot_identification.py (root)  path = PARENT of the directory holding the script: <ws>/<repo>/spot_identification.py
                                   reads <ws>/data/spot_robot_*.dat and <ws>/files/spot_description/...

make_workspace() lays that out under a scratch directory: the script is COPIED byte for byte from the staged reference
checkout (baseline/_ref, written by __graft_entry__.build() from /root/reference; git-ignored, travels to the GPU box),
files/ is a symlink to the staged files, and the five .dat files hold a seeded synthetic log in the reference's format
(tab-separated '%.6f', channels x N -- what g1-data/csv2dat.py:50-55 writes and np.loadtxt reads).
"""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
SCRIPTS = {
    # name: (script relative to the reference root, where it goes in the workspace, data dir, file prefix, robot)
    "demo_solo": ("demo/solo_identification.py", "demo/solo_identification.py", "data/solo", "solo", "solo12"),
    "demo_spot": ("demo/spot_identification.py", "demo/spot_identification.py", "data", "spot", "spot"),
    "root_spot": ("spot_identification.py", "repo/spot_identification.py", "data", "spot", "spot"),
}


def staged_available():
    return all(os.path.exists(os.path.join(STAGED, s[0])) for s in SCRIPTS.values()) and os.path.isdir(os.path.join(STAGED, "files"))


def write_dat(path, a):
    np.savetxt(path, np.asarray(a), fmt="%.6f", delimiter="\t")


def synthetic_log(robot, N, seed=None, tau_from_truth=True):
    """Five (channels x N) arrays for `robot`; torques from a perturbed ground truth when a CUDA device is there."""
    sys.path.insert(0, ROOT)
    from system_identification_b200 import synth
    from system_identification_b200.model import FlatModel
    flat = FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", robot + ".json"))
    q, dq, ddq, cnt = synth.make_trajectory(flat, N, synth.SEEDS[robot] if seed is None else seed)
    tau = synth.synth_tau(flat, N, 3)
    if tau_from_truth:
        import torch
        if torch.cuda.is_available():
            from system_identification_b200.ops import DeviceModel, to_device
            dm = DeviceModel(flat)
            dev = [to_device(a) for a in (q, dq, ddq, tau, cnt)]
            noise = 0.05 if robot.startswith("solo") else 0.5
            tau = synth.identifiable_tau_device(flat, dm, dev, seed=29, perturb=0.1, bv_max=0.02 if robot.startswith("solo") else 0.2,
                                                bc_max=0.05 if robot.startswith("solo") else 0.5, noise=noise).cpu().numpy()
    return flat, (q, dq, ddq, tau, cnt)


def make_workspace(ws, which, N, seed=None, tau_from_truth=True):
    """Returns (path of the copied script, flat model, the five arrays written)."""
    src_rel, dst_rel, data_rel, prefix, robot = SCRIPTS[which]
    os.makedirs(ws, exist_ok=True)
    script = os.path.join(ws, dst_rel)
    os.makedirs(os.path.dirname(script), exist_ok=True)
    shutil.copyfile(os.path.join(STAGED, src_rel), script)
    files = os.path.join(ws, "files")
    if not os.path.exists(files):
        os.symlink(os.path.join(STAGED, "files"), files)
    flat, data = synthetic_log(robot, N, seed, tau_from_truth)
    d = os.path.join(ws, data_rel)
    os.makedirs(d, exist_ok=True)
    for name, a in zip(("q", "dq", "ddq", "tau", "contact"), data):
        write_dat(os.path.join(d, f"{prefix}_robot_{name}.dat"), a)
    return script, flat, data


if __name__ == "__main__":
    # python tools/demo_workspace.py demo_solo 20000 /tmp/ws   -> prints the script to run with `python -m system_identification_b200.run`
    which, N, ws = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    print(make_workspace(ws, which, N)[0])
