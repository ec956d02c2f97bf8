#!/usr/bin/env python
"""Pins the oracle (and through it the CUDA path) against the REAL dependencies of the reference: pinocchio and, where
available, cvxpy with a conic solver.  Run it on any box where `import pinocchio` works:

    python tests/golden/make_pinocchio_golden.py --reference /path/to/system_identification [--robots solo12 spot g1_12dof]

It writes tests/golden/pinocchio_<robot>.npz; tests/test_pinocchio_golden.py picks those files up automatically (the
oracle is compared on CPU, the CUDA path under `-m gpu`) and is skipped while they are absent.  Neither pinocchio nor
cvxpy is installable in the build image (no network, not in the wheelhouse), which is why no such file is committed yet:
until one is, DESIGN.md section 2 says "parity unpinned against pinocchio/MOSEK".

What is recorded, per robot, on the seeded 48-sample log of the committed oracle fixture tests/golden/<robot>_N48.npz:

  stage 1   Y = pin.computeJointTorqueRegressor(model, data, q, dq, ddq)        reference src/sys_identification.py:406
            after the reference's own _update_fk sequence (forwardKinematics, framesForwardKinematics,
            computeJointJacobians; :113-117); J_c rows = getFrameJacobian(..., LOCAL_WORLD_ALIGNED)[0:3] for the feet in
            contact (:119-129); P = I - pinv(J_c) J_c (:131-135)
  model     joint names / parents / placements / idx_q / idx_v, foot frame parents and offsets, gravity: checks the
            URDF flattening and the joint order
  quirk Q1  pin.Inertia.toDynamicParameters() of a known inertia: the column order [m, mc, Ixx, Ixy, Iyy, Ixz, Iyz, Izz]
  stage 3   when cvxpy and the reference's own src/ import: Solver(...).solve_fully_consistent() of the stack
            (reference src/solver.py:123-210), with MOSEK when licensed, otherwise the same cvxpy problem handed to
            CLARABEL (or SCS) by swapping the `solver=` argument only -- the problem construction stays the reference's
  priors    when the reference's SystemIdentification imports (needs trimesh + urdf_parser_py): get_phi_prior(),
            get_bounding_ellipsoids()
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

ROBOTS = {
    "solo12": ("files/solo_description/solo12.urdf", "files/solo_description/solo12_config.yaml"),
    "spot": ("files/spot_description/spot.urdf", "files/spot_description/spot_config.yaml"),
    "g1_12dof": ("files/g1_description/g1_12dof.urdf", None),       # the reference's g1_config.yaml is a copy of Spot's
}
N_GOLD = 48


def seeded_log(name):
    """The same inputs as tests/golden/<robot>_N48.npz (the committed oracle fixture: seeded trajectory, ground-truth torques)."""
    from system_identification_b200.model import FlatModel
    flat = FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", name + ".json"))
    g = np.load(os.path.join(HERE, f"{name}_N48.npz"))
    return flat, tuple(np.array(g[k]) for k in ("q", "dq", "ddq", "tau", "cnt"))


def pinocchio_part(pin, urdf, flat, data):
    q, dq, ddq, tau, cnt = data
    model = pin.buildModelFromUrdf(urdf, pin.JointModelFreeFlyer())          # src/sys_identification.py:16
    rdata = model.createData()
    model.gravity.linear = np.array([0, 0, -9.81])                           # :22
    ee_ids = [model.getFrameId(n) for n in flat.ee_names]                    # :50-54
    out = {
        "joint_names": np.array(list(model.names)),
        "parents": np.array(list(model.parents), dtype=np.int32),
        "idx_q": np.array([model.joints[i].idx_q for i in range(model.njoints)], dtype=np.int32),
        "idx_v": np.array([model.joints[i].idx_v for i in range(model.njoints)], dtype=np.int32),
        "joint_shortnames": np.array([model.joints[i].shortname() for i in range(model.njoints)]),
        "place_R": np.array([np.array(model.jointPlacements[i].rotation) for i in range(model.njoints)]),
        "place_p": np.array([np.array(model.jointPlacements[i].translation) for i in range(model.njoints)]),
        "ee_parent": np.array([model.frames[f].parentJoint if hasattr(model.frames[f], "parentJoint") else model.frames[f].parent
                               for f in ee_ids], dtype=np.int32),
        "ee_offset": np.array([np.array(model.frames[f].placement.translation) for f in ee_ids]),
        "body_dyn_params": np.array([np.array(model.inertias[i].toDynamicParameters()) for i in range(model.njoints)]),
        "nq": model.nq, "nv": model.nv,
    }
    # quirk Q1: the order of toDynamicParameters
    I3 = np.array([[1.0, 0.2, 0.3], [0.2, 2.0, 0.5], [0.3, 0.5, 3.0]])
    out["q1_dyn_params_of_known_inertia"] = np.array(pin.Inertia(1.5, np.array([0.1, -0.2, 0.3]), I3).toDynamicParameters())
    Y, P, Jc = [], [], []
    nv = model.nv
    for i in range(q.shape[1]):
        qi, vi, ai = q[:, i].copy(), dq[:, i].copy(), ddq[:, i].copy()
        pin.forwardKinematics(model, rdata, qi, vi, ai)                      # _update_fk, :113-117
        pin.framesForwardKinematics(model, rdata, qi)
        pin.computeJointJacobians(model, rdata, qi)
        Y.append(np.array(pin.computeJointTorqueRegressor(model, rdata, qi, vi, ai)))      # :406
        c = cnt[:, i]
        m = int(np.sum(c))                                                   # _compute_J_c, :119-129
        J = np.zeros((3 * m, nv))
        j = 0
        for k in range(len(ee_ids)):
            if c[k]:
                J[j:j + 3, :] = pin.getFrameJacobian(model, rdata, ee_ids[k], pin.LOCAL_WORLD_ALIGNED)[0:3, :]
                j += 3
        full = np.zeros((12, nv)); full[:J.shape[0]] = J[:12]
        Jc.append(full)
        P.append(np.eye(nv) - np.linalg.pinv(J) @ J)                         # :131-135
    out["Y"] = np.array(Y); out["P"] = np.array(P); out["J_c_padded"] = np.array(Jc)
    out["pinocchio_version"] = str(getattr(pin, "__version__", "unknown"))
    return out


def reference_part(reference_root, urdf, yaml_path, data):
    """Runs the reference's own classes (needs pinocchio, trimesh, urdf_parser_py, cvxpy).  Returns {} when they do not import."""
    q, dq, ddq, tau, cnt = data
    sys.path.insert(0, reference_root)
    try:
        import cvxpy as cp
        from src.solver import Solver                         # the REFERENCE's module: reference_root is first on sys.path
        from src.sys_identification import SystemIdentification
    except Exception as e:                                    # noqa: BLE001
        print("  reference classes not importable:", e)
        return {}
    finally:
        sys.path.pop(0)
    si = SystemIdentification(urdf, yaml_path, floating_base=True)
    out = {"phi_prior": np.array(si.get_phi_prior()),
           "ellipsoid_semi_axes": np.array([e["semi_axes"] for e in si.get_bounding_ellipsoids()]),
           "ellipsoid_centers": np.array([e["center"] for e in si.get_bounding_ellipsoids()]),
           "robot_mass": float(si.get_robot_mass())}
    Ys, Ts, Bv, Bc = [], [], [], []
    for i in range(q.shape[1]):                               # the demo's loops, demo/solo_identification.py:36-55
        y, t = si.get_proj_regressor_torque(q[:, i], dq[:, i], ddq[:, i], tau[:, i], cnt[:, i])
        bv, bc = si.get_proj_friction_regressors(q[:, i], dq[:, i], ddq[:, i], cnt[:, i])
        Ys.append(y); Ts.append(t); Bv.append(bv); Bc.append(bc)
    Ys, Ts, Bv, Bc = np.vstack(Ys), np.hstack(Ts), np.vstack(Bv), np.vstack(Bc)
    out.update(Y_proj=Ys, tau_proj=Ts, B_v=Bv, B_c=Bc)
    installed = cp.installed_solvers()
    solver_name = "MOSEK" if "MOSEK" in installed else ("CLARABEL" if "CLARABEL" in installed else ("SCS" if "SCS" in installed else None))
    if solver_name is None:
        print("  no conic solver installed; stage 3 not recorded")
        return out
    orig_solve = cp.Problem.solve
    if solver_name != "MOSEK":
        def swapped(self, *a, **k):                           # same problem, stated fallback solver (north_star)
            k.pop("mosek_params", None); k["solver"] = getattr(cp, solver_name)
            if solver_name == "SCS":
                k.setdefault("eps", 1e-9); k.setdefault("max_iters", 200000)
            return orig_solve(self, *a, **k)
        cp.Problem.solve = swapped
    try:
        sol = Solver(Ys, Ts, si.get_num_links(), si.get_phi_prior(), si.get_robot_mass(), si.get_bounding_ellipsoids(), B_v=Bv, B_c=Bc)
        phi = sol.solve_fully_consistent()                    # src/solver.py:123-210
        out.update(phi_identified=np.array(phi), b_v=np.array(sol._b_v.value), b_c=np.array(sol._b_c.value), sdp_solver=solver_name,
                   cvxpy_version=str(cp.__version__))
    except Exception as e:                                    # noqa: BLE001
        print("  stage 3 failed:", e)
    finally:
        cp.Problem.solve = orig_solve
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference", help="checkout of xiaohu97/system_identification (URDFs, YAMLs, src/)")
    ap.add_argument("--robots", nargs="*", default=list(ROBOTS))
    args = ap.parse_args()
    try:
        import pinocchio as pin
    except Exception as e:                                    # noqa: BLE001
        print("pinocchio is not importable here (%s): nothing written.  Run this script on a box that has it." % e)
        return 2
    for name in args.robots:
        urdf_rel, yaml_rel = ROBOTS[name]
        urdf = os.path.join(args.reference, urdf_rel)
        yaml_path = os.path.join(args.reference, yaml_rel) if yaml_rel else os.path.join(ROOT, "system_identification_b200", "robots", "g1_12dof_config.yaml")
        flat, data = seeded_log(name)
        print(name, "...")
        out = {"q": data[0], "dq": data[1], "ddq": data[2], "tau": data[3], "cnt": data[4]}
        out.update(pinocchio_part(pin, urdf, flat, data))
        if os.path.exists(yaml_path):
            out.update(reference_part(args.reference, urdf, yaml_path, data))
        path = os.path.join(HERE, f"pinocchio_{name}.npz")
        np.savez_compressed(path, **out)
        print("  wrote", path, "keys:", sorted(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
