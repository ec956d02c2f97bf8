"""Host-side URDF front end: URDF + YAML -> FlatModel, priors and bounding ellipsoids.

This is the once-per-run host step of the path.  It replaces
three third-party packages the reference imports and the build image does not have:
  * pinocchio.buildModelFromUrdf(path, JointModelFreeFlyer())   reference src/sys_identification.py:16
  * urdf_parser_py.urdf.URDF.from_xml_file                     reference src/sys_identification.py:236,271
  * trimesh.load_mesh(...).bounding_box                        reference src/sys_identification.py:258-261
"""
from __future__ import annotations

import math
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

from .model import FlatModel, JT_FF, JT_RX, JT_RY, JT_RZ, JT_RU

_AXIS_TYPES = {(1.0, 0.0, 0.0): JT_RX, (0.0, 1.0, 0.0): JT_RY, (0.0, 0.0, 1.0): JT_RZ}


def rpy_matrix(roll, pitch, yaw):
    """Fixed-axis roll/pitch/yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    sr, cr = math.sin(roll), math.cos(roll)
    sp, cp = math.sin(pitch), math.cos(pitch)
    sy, cy = math.sin(yaw), math.cos(yaw)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr]], dtype=np.float64)


def hat(v):
    x, y, z = v
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def _vec(text, default=(0.0, 0.0, 0.0)):
    return np.array([float(t) for t in text.split()] if text else default, dtype=np.float64)


class _Pose:
    __slots__ = ("R", "p")

    def __init__(self, R=None, p=None):
        self.R = np.eye(3) if R is None else R
        self.p = np.zeros(3) if p is None else p

    def __mul__(self, o):
        return _Pose(self.R @ o.R, self.p + self.R @ o.p)

    @staticmethod
    def from_xml(elem):
        if elem is None:
            return _Pose()
        return _Pose(rpy_matrix(*_vec(elem.get("rpy"))), _vec(elem.get("xyz")))


class UrdfRobot:
    """Minimal URDF document: links (inertial + visuals) and joints, in file order."""

    def __init__(self, path):
        self.path = path
        doc = ET.parse(path).getroot()
        self.name = doc.get("name", "robot")
        self.links = {}        # name -> dict, insertion order == file order
        self.joints = {}       # name -> dict
        for e in doc.findall("link"):
            self.links[e.get("name")] = self._read_link(e)
        for e in doc.findall("joint"):
            lim = e.find("limit")
            ax = e.find("axis")
            self.joints[e.get("name")] = {
                "name": e.get("name"), "type": e.get("type"),
                "parent": e.find("parent").get("link"), "child": e.find("child").get("link"),
                "pose": _Pose.from_xml(e.find("origin")),
                "axis": _vec(ax.get("xyz") if ax is not None else None, (1.0, 0.0, 0.0)),
                "lower": float(lim.get("lower", 0.0)) if lim is not None else 0.0,
                "upper": float(lim.get("upper", 0.0)) if lim is not None else 0.0,
            }

    @staticmethod
    def _read_link(e):
        out = {"name": e.get("name"), "inertial": None, "visuals": []}
        ine = e.find("inertial")
        if ine is not None:
            o = ine.find("origin")
            t = ine.find("inertia")
            comp = {k: float(t.get(k, 0.0)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")}
            out["inertial"] = {
                "mass": float(ine.find("mass").get("value")),
                "xyz": _vec(o.get("xyz") if o is not None else None),
                "rpy": _vec(o.get("rpy") if o is not None else None),
                "I": np.array([[comp["ixx"], comp["ixy"], comp["ixz"]],
                               [comp["ixy"], comp["iyy"], comp["iyz"]],
                               [comp["ixz"], comp["iyz"], comp["izz"]]]),
            }
        for v in e.findall("visual"):
            o = v.find("origin")
            shape = list(v.find("geometry"))[0]
            out["visuals"].append({
                "xyz": _vec(o.get("xyz")) if o is not None else None,
                "rpy": _vec(o.get("rpy")) if o is not None else np.zeros(3),
                "shape": shape.tag, "attrib": dict(shape.attrib)})
        return out

    # ------------------------------------------------------------------ kinematic tree
    def flatten(self, floating_base=True, gravity=(0.0, 0.0, -9.81)) -> FlatModel:
        # floating_base=False (reference src/sys_identification.py:15-18: pin.buildModelFromUrdf(path), the root link welded to the
        # universe) is EMULATED on the free-flyer kernels: the same tree, the base pinned at the identity with zero twist, and three
        # virtual point contacts on the root body that are always "in stance" -- their Jacobian rows [I | -[r_k]x | 0] span the six
        # base coordinates, so the contact null space is exactly {base twist = 0} x null(J_c restricted to the joints)
        # (load_robot adds the contacts; SystemIdentification converts inputs and slices outputs)
        kids = {n: [] for n in self.links}
        has_parent = set()
        for j in self.joints.values():
            kids[j["parent"]].append(j)
            has_parent.add(j["child"])
        root = [n for n in self.links if n not in has_parent]
        if len(root) != 1:
            raise ValueError(f"URDF needs exactly one root link, got {root}")

        rec = {"names": ["universe", "root_joint"], "parent": [0, 0], "jtype": [-1, JT_FF],
               "axis": [np.zeros(3), np.zeros(3)], "pose": [_Pose(), _Pose()], "lo": [0.0, 0.0], "hi": [0.0, 0.0],
               "params": [np.zeros(10), np.zeros(10)]}
        frames = {}
        # depth-first, pre-order joint numbering; children visited in joint-name order (urdfdom's std::map)
        def enter(link, jid, pose):
            frames[link] = (jid, pose)
            rec["params"][jid] = rec["params"][jid] + self._body_params(self.links[link], pose)
            for j in sorted(kids[link], key=lambda d: d["name"]):
                jp = pose * j["pose"]
                if j["type"] == "fixed":
                    enter(j["child"], jid, jp)
                elif j["type"] == "revolute":
                    rec["names"].append(j["name"]); rec["parent"].append(jid)
                    rec["jtype"].append(_AXIS_TYPES.get(tuple(j["axis"].tolist()), JT_RU))
                    rec["axis"].append(j["axis"] / np.linalg.norm(j["axis"])); rec["pose"].append(jp)
                    rec["lo"].append(j["lower"]); rec["hi"].append(j["upper"]); rec["params"].append(np.zeros(10))
                    enter(j["child"], len(rec["names"]) - 1, _Pose())
                else:
                    raise ValueError(f"joint {j['name']}: type {j['type']!r} is not supported on this path")

        enter(root[0], 1, _Pose())
        m = FlatModel(
            name=self.name, joint_names=rec["names"], parent=np.array(rec["parent"], dtype=np.int32),
            jtype=np.array(rec["jtype"], dtype=np.int32), axis=np.array(rec["axis"]),
            place_R=np.array([p.R for p in rec["pose"]]), place_p=np.array([p.p for p in rec["pose"]]),
            lower=np.array(rec["lo"]), upper=np.array(rec["hi"]), gravity=np.array(gravity, dtype=np.float64),
            body_params=np.array(rec["params"]))
        m._frames = frames
        m.floating_base = bool(floating_base)
        return m

    @staticmethod
    def _body_params(link, pose):
        """[m, m c, Ixx, Ixy, Iyy, Ixz, Iyz, Izz] about the joint origin (pinocchio's
        Inertia::toDynamicParameters order); additive over the links folded into one joint."""
        ine = link["inertial"]
        if ine is None:
            return np.zeros(10)
        m = ine["mass"]
        c = pose.p + pose.R @ ine["xyz"]
        Rt = pose.R @ rpy_matrix(*ine["rpy"])
        Io = Rt @ ine["I"] @ Rt.T - m * hat(c) @ hat(c)
        return np.array([m, *(m * c), Io[0, 0], Io[0, 1], Io[1, 1], Io[0, 2], Io[1, 2], Io[2, 2]])

    # ------------------------------------------------------------------ priors (reference :266-322)
    def phi_prior(self, link_names):
        """float32 vector, reference order [m, h, Ixx, Ixy, Ixz, Iyy, Iyz, Izz] per link; the bare
        <inertial> of each listed link, taken in URDF file order (as the reference iterates)."""
        chosen = [l for l in self.links.values() if l["name"] in link_names]
        out = np.zeros(10 * len(link_names), dtype=np.float32)
        for i in range(len(link_names)):
            ine = chosen[i]["inertial"]
            m, c = ine["mass"], ine["xyz"]
            R = rpy_matrix(*ine["rpy"])
            Ibar = R @ ine["I"] @ R.T + m * hat(c) @ hat(c).T
            out[10 * i] = m
            out[10 * i + 1:10 * i + 4] = m * c
            out[10 * i + 4:10 * i + 7] = Ibar[0]
            out[10 * i + 7:10 * i + 9] = Ibar[1, 1:]
            out[10 * i + 9] = Ibar[2, 2]
        return out

    # ------------------------------------------------------------------ ellipsoids (reference :235-264)
    def bounding_ellipsoids(self, link_names, files_root, mesh_fallbacks=None):
        out = []
        for link in self.links.values():
            if link["name"] not in link_names:
                continue
            for vis in link["visuals"]:
                a = vis["attrib"]
                origin = vis["xyz"] if vis["xyz"] is not None else np.zeros(3)
                if vis["shape"] == "box":
                    semi, center = _vec(a["size"]) / 2, origin
                elif vis["shape"] == "cylinder":
                    semi, center = np.array([float(a["radius"]), float(a["radius"]), float(a["length"]) / 2]), origin
                elif vis["shape"] == "sphere":
                    semi, center = np.full(3, float(a["radius"])), origin
                elif vis["shape"] == "mesh":
                    path = self._mesh_path(a["filename"], files_root)
                    if not os.path.exists(path) and mesh_fallbacks and a["filename"] in mesh_fallbacks:
                        path = self._mesh_path(mesh_fallbacks[a["filename"]], files_root)
                    lo, hi = mesh_bounds(path)
                    semi, center = (hi - lo) / 2, (hi + lo) / 2 + origin
                else:
                    raise ValueError(f"Unsupported geometry type for link {link['name']}")
                out.append({"semi_axes": semi, "center": center})
        return out

    def _mesh_path(self, filename, files_root):
        # reference rule: <repo>/files/ + filename[10:] (strips 'package://'); G1 URDFs use bare
        # relative paths, which that rule mangles -> resolve those against the URDF directory.
        if filename.startswith("package://"):
            return os.path.join(files_root, filename[len("package://"):])
        return os.path.join(os.path.dirname(os.path.abspath(self.path)), filename)


def mesh_bounds(path):
    """Axis-aligned bounds of a mesh's vertices: binary STL, ASCII STL or Wavefront OBJ."""
    low = path.lower()
    if low.endswith(".obj"):
        pts = np.loadtxt((ln[2:] for ln in open(path, "r", errors="ignore") if ln.startswith("v ")), usecols=(0, 1, 2), ndmin=2)
        return pts.min(axis=0), pts.max(axis=0)
    if low.endswith(".stl"):
        raw = open(path, "rb").read()
        if len(raw) >= 84:
            (count,) = struct.unpack_from("<I", raw, 80)
            if 84 + 50 * count == len(raw):
                tri = np.ndarray((count, 12), dtype="<f4", buffer=raw, offset=84, strides=(50, 4))
                pts = tri[:, 3:12].reshape(-1, 3).astype(np.float64)
                return pts.min(axis=0), pts.max(axis=0)
        pts = np.array([[float(t) for t in ln.split()[1:4]] for ln in raw.decode("ascii", "ignore").splitlines()
                        if ln.strip().startswith("vertex")])
        return pts.min(axis=0), pts.max(axis=0)
    raise ValueError(f"unsupported mesh file {path}")


def merged_priors(robot: UrdfRobot, m: FlatModel, files_root, mesh_fallbacks=None):
    """Deviation from reference style, for robots whose listed links carry many fixed children
    (G1-12dof: torso, head and arms are welded to the pelvis; quirk Q9 would otherwise pit a 3.8 kg
    pelvis prior against a 32 kg total-mass constraint).  Prior of body i = the merged inertia the
    kinematic model itself uses; ellipsoid of body i = AABB of every visual welded to that joint."""
    phi = np.zeros(10 * m.nbodies, dtype=np.float32)
    for i in range(1, m.njoints):
        pm = m.body_params[i]
        phi[10 * (i - 1):10 * i] = [pm[0], pm[1], pm[2], pm[3], pm[4], pm[5], pm[7], pm[6], pm[8], pm[9]]
    boxes = {i: [np.full(3, np.inf), np.full(3, -np.inf)] for i in range(1, m.njoints)}
    for name, (jid, pose) in m._frames.items():
        for vis in robot.links[name]["visuals"]:
            a = vis["attrib"]
            vp = pose * _Pose(rpy_matrix(*vis["rpy"]), vis["xyz"] if vis["xyz"] is not None else np.zeros(3))
            if vis["shape"] == "mesh":
                path = robot._mesh_path(a["filename"], files_root)
                if not os.path.exists(path) and mesh_fallbacks and a["filename"] in mesh_fallbacks:
                    path = robot._mesh_path(mesh_fallbacks[a["filename"]], files_root)
                if not os.path.exists(path):
                    continue
                lo, hi = mesh_bounds(path)
            elif vis["shape"] == "box":
                hi = _vec(a["size"]) / 2; lo = -hi
            elif vis["shape"] == "cylinder":
                hi = np.array([float(a["radius"]), float(a["radius"]), float(a["length"]) / 2]); lo = -hi
            elif vis["shape"] == "sphere":
                hi = np.full(3, float(a["radius"])); lo = -hi
            else:
                raise ValueError(f"Unsupported geometry type for link {name}")
            corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
            w = corners @ vp.R.T + vp.p
            boxes[jid][0] = np.minimum(boxes[jid][0], w.min(0)); boxes[jid][1] = np.maximum(boxes[jid][1], w.max(0))
    ell = [{"semi_axes": (boxes[i][1] - boxes[i][0]) / 2, "center": (boxes[i][1] + boxes[i][0]) / 2} for i in range(1, m.njoints)]
    return phi, ell


FIXED_BASE_ANCHORS = ((0.25, 0.0, 0.0), (0.0, 0.25, 0.0), (0.0, 0.0, 0.25))      # three non-collinear points on the root body


def load_robot(urdf_file, config, floating_base=True, files_root=None, mesh_fallbacks=None, merged=False) -> FlatModel:
    """URDF + parsed YAML 'robot' section -> fully populated FlatModel."""
    robot = UrdfRobot(urdf_file)
    m = robot.flatten(floating_base)
    m.name = config.get("name") or m.name
    m.robot_mass = config.get("mass")
    m.link_names = list(config.get("link_names", []))
    m.ee_names = list(config.get("end_effectors_frame_names", []) or [])
    ee_joint, ee_off = [], []
    if not floating_base:
        for k, off in enumerate(FIXED_BASE_ANCHORS):
            ee_joint.append(1); ee_off.append(np.array(off))
        if len(m.ee_names) > 1:
            raise ValueError("fixed-base emulation leaves room for one contact frame (three of the four contact slots pin the base)")
    for n in m.ee_names:
        if n not in m._frames:
            raise ValueError(f"end-effector frame {n!r} is not a link of {urdf_file}")
        jid, pose = m._frames[n]
        ee_joint.append(jid); ee_off.append(pose.p)
    if not floating_base:
        m.ee_names = [f"__fixed_base_anchor_{k}" for k in range(len(FIXED_BASE_ANCHORS))] + m.ee_names
    m.ee_joint = np.array(ee_joint, dtype=np.int32)
    m.ee_offset = np.array(ee_off, dtype=np.float64).reshape(-1, 3)
    if files_root is None:
        files_root = os.path.dirname(os.path.dirname(os.path.abspath(urdf_file)))
    if merged:
        m.phi_prior, m.ellipsoids = merged_priors(robot, m, files_root, mesh_fallbacks)
    else:
        m.ellipsoids = robot.bounding_ellipsoids(m.link_names, files_root, mesh_fallbacks)
        m.phi_prior = robot.phi_prior(m.link_names) if m.link_names else None
    return m
