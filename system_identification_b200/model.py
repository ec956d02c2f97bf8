"""Flat robot model: what the host uploads once through the C-ABI (`sysid_model_create`).

Replaces the pinocchio Model/Data pair the reference builds in
SystemIdentification.__init__ (reference src/sys_identification.py:11-73): joint 0 = universe,
joint 1 = free-flyer root, then the moving URDF joints depth-first with children ordered by joint
name; fixed joints are folded into their moving ancestor.  Everything the device needs is a few
hundred bytes: parents, joint types/axes, joint placements and the end-effector points.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

JT_FF, JT_RX, JT_RY, JT_RZ, JT_RU = 0, 1, 2, 3, 4


@dataclass
class FlatModel:
    name: str
    joint_names: list                 # len njoints, [0] == "universe"
    parent: np.ndarray                # int32 [njoints]
    jtype: np.ndarray                 # int32 [njoints]; -1 for the universe
    axis: np.ndarray                  # f64 [njoints,3]
    place_R: np.ndarray               # f64 [njoints,3,3]
    place_p: np.ndarray               # f64 [njoints,3]
    lower: np.ndarray                 # f64 [njoints] joint limits (synthetic trajectories only)
    upper: np.ndarray
    gravity: np.ndarray               # f64 [3]
    ee_names: list = field(default_factory=list)
    ee_joint: np.ndarray = None       # int32 [n_ee] parent joint of each end-effector frame
    ee_offset: np.ndarray = None      # f64 [n_ee,3] frame origin in that joint's frame
    link_names: list = field(default_factory=list)
    robot_mass: float = 0.0
    phi_prior: np.ndarray = None      # float32 [10*L], REFERENCE order (m,h,Ixx,Ixy,Ixz,Iyy,Iyz,Izz)
    ellipsoids: list = field(default_factory=list)   # [{'semi_axes':(3,), 'center':(3,)}]
    body_params: np.ndarray = None    # f64 [njoints,10] merged URDF inertias, PINOCCHIO order (synthetic ground truth)
    floating_base: bool = True

    # ---- sizes -------------------------------------------------------------------------------
    @property
    def njoints(self):
        return int(self.parent.shape[0])

    @property
    def nbodies(self):
        return self.njoints - 1

    @property
    def nq(self):
        return int(sum(7 if t == JT_FF else 1 for t in self.jtype[1:]))

    @property
    def nv(self):
        return int(sum(6 if t == JT_FF else 1 for t in self.jtype[1:]))

    @property
    def base_dof(self):
        return 6 if self.floating_base else 0

    @property
    def joints_dof(self):
        return self.nv - self.base_dof

    @property
    def n_ee(self):
        return len(self.ee_names)

    @property
    def num_params(self):
        return 10 * self.nbodies

    def ncols(self, friction=True):
        return self.num_params + (2 * self.joints_dof if friction else 0)

    def depth(self):
        d = np.zeros(self.njoints, dtype=np.int32)
        for i in range(1, self.njoints):
            d[i] = d[self.parent[i]] + 1
        return d

    # ---- (de)serialisation ------------------------------------------------------------------
    def to_json(self):
        def arr(a):
            return None if a is None else np.asarray(a).tolist()
        return json.dumps({
            "name": self.name, "joint_names": self.joint_names, "parent": arr(self.parent), "jtype": arr(self.jtype),
            "axis": arr(self.axis), "place_R": arr(self.place_R), "place_p": arr(self.place_p),
            "lower": arr(self.lower), "upper": arr(self.upper), "gravity": arr(self.gravity),
            "ee_names": self.ee_names, "ee_joint": arr(self.ee_joint), "ee_offset": arr(self.ee_offset),
            "link_names": self.link_names, "robot_mass": self.robot_mass,
            "phi_prior_f32": None if self.phi_prior is None else [float(np.float32(v)) for v in self.phi_prior],
            "ellipsoids": [{"semi_axes": arr(e["semi_axes"]), "center": arr(e["center"])} for e in self.ellipsoids],
            "body_params": arr(self.body_params), "floating_base": self.floating_base,
        }, indent=1)

    @staticmethod
    def from_json(text):
        d = json.loads(text)
        return FlatModel(
            name=d["name"], joint_names=d["joint_names"], parent=np.array(d["parent"], dtype=np.int32),
            jtype=np.array(d["jtype"], dtype=np.int32), axis=np.array(d["axis"], dtype=np.float64),
            place_R=np.array(d["place_R"], dtype=np.float64), place_p=np.array(d["place_p"], dtype=np.float64),
            lower=np.array(d["lower"], dtype=np.float64), upper=np.array(d["upper"], dtype=np.float64),
            gravity=np.array(d["gravity"], dtype=np.float64), ee_names=d["ee_names"],
            ee_joint=np.array(d["ee_joint"], dtype=np.int32), ee_offset=np.array(d["ee_offset"], dtype=np.float64).reshape(-1, 3),
            link_names=d["link_names"], robot_mass=d["robot_mass"],
            phi_prior=None if d["phi_prior_f32"] is None else np.array(d["phi_prior_f32"], dtype=np.float32),
            ellipsoids=[{"semi_axes": np.array(e["semi_axes"]), "center": np.array(e["center"])} for e in d["ellipsoids"]],
            body_params=None if d["body_params"] is None else np.array(d["body_params"], dtype=np.float64),
            floating_base=d["floating_base"])

    @staticmethod
    def load(path):
        with open(path, "r") as f:
            return FlatModel.from_json(f.read())

    def save(self, path):
        with open(path, "w") as f:
            f.write(self.to_json())
