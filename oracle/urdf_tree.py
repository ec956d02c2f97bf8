"""ORACLE (test infrastructure, not product code) -- URDF -> flat kinematic tree.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  PARITY UNPINNED: pinocchio / urdf_parser_py / trimesh are not installable in
the build image, so this is a restatement of their documented behaviour, validated independently
(RNEA identity, finite-difference Jacobians, the demo/RUN_DEMO.md prior table).

What it restates (reference = xiaohu97/system_identification):
  * src/sys_identification.py:16   pin.buildModelFromUrdf(path, JointModelFreeFlyer())
  * src/sys_identification.py:22   gravity (0, 0, -9.81)
  * src/sys_identification.py:51-54 getFrameId(end-effector names)
  * src/sys_identification.py:235-264 _compute_bounding_ellipsoids (urdf_parser_py + trimesh AABB)
  * src/sys_identification.py:266-322 _compute_inertial_params / get_phi_prior (float32 vector)

pinocchio conventions restated here (upstream):
  joint 0 = universe, joint 1 = "root_joint" free-flyer (nq 7 [x y z qx qy qz qw], nv 6, local
  frame), then the non-fixed URDF joints in depth-first order with each link's children sorted by
  JOINT NAME (urdfdom keeps joints in a std::map); fixed joints add no joint: the child link's
  inertia and frame are merged into the nearest moving ancestor with the accumulated offset.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

JT_FF, JT_RX, JT_RY, JT_RZ, JT_RU = 0, 1, 2, 3, 4


def rpy_to_matrix(r, p, y):
    """R = Rz(yaw) Ry(pitch) Rx(roll)  (URDF fixed-axis rpy; same as pin.utils.rpyToMatrix)."""
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    return Rz @ Ry @ Rx


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def _floats(s, n, default):
    if s is None:
        return np.array(default, dtype=np.float64)
    vals = [float(t) for t in s.split()]
    assert len(vals) == n, f"expected {n} floats in {s!r}"
    return np.array(vals, dtype=np.float64)


def _origin(elem):
    """(<origin xyz rpy>) -> (R, p); missing element or attribute = identity."""
    if elem is None:
        return np.eye(3), np.zeros(3)
    xyz = _floats(elem.get("xyz"), 3, [0, 0, 0])
    rpy = _floats(elem.get("rpy"), 3, [0, 0, 0])
    return rpy_to_matrix(*rpy), xyz


@dataclass
class UrdfLink:
    name: str
    mass: float | None = None
    com: np.ndarray | None = None       # inertial origin xyz
    rpy: np.ndarray | None = None       # inertial origin rpy
    I_c: np.ndarray | None = None       # 3x3 about the CoM, inertial-frame axes
    visuals: list = field(default_factory=list)  # dicts {kind, origin_xyz, ...}


@dataclass
class UrdfJoint:
    name: str
    jtype: str
    parent: str
    child: str
    R: np.ndarray
    p: np.ndarray
    axis: np.ndarray
    lower: float
    upper: float


def parse_urdf(path):
    root = ET.parse(path).getroot()
    links, joints = [], []
    for le in root.findall("link"):
        ln = UrdfLink(name=le.get("name"))
        ie = le.find("inertial")
        if ie is not None:
            oe = ie.find("origin")
            ln.com = _floats(oe.get("xyz") if oe is not None else None, 3, [0, 0, 0])
            ln.rpy = _floats(oe.get("rpy") if oe is not None else None, 3, [0, 0, 0])
            ln.mass = float(ie.find("mass").get("value"))
            it = ie.find("inertia")
            g = lambda k: float(it.get(k, "0"))
            ln.I_c = np.array([[g("ixx"), g("ixy"), g("ixz")],
                               [g("ixy"), g("iyy"), g("iyz")],
                               [g("ixz"), g("iyz"), g("izz")]], dtype=np.float64)
        for ve in le.findall("visual"):
            oe = ve.find("origin")
            xyz = _floats(oe.get("xyz") if oe is not None else None, 3, [0, 0, 0])
            ge = ve.find("geometry")
            vis = {"origin_xyz": xyz, "has_origin": oe is not None}
            child = list(ge)[0]
            vis["kind"] = child.tag
            if child.tag == "box":
                vis["size"] = _floats(child.get("size"), 3, None)
            elif child.tag == "cylinder":
                vis["radius"] = float(child.get("radius"))
                vis["length"] = float(child.get("length"))
            elif child.tag == "sphere":
                vis["radius"] = float(child.get("radius"))
            elif child.tag == "mesh":
                vis["filename"] = child.get("filename")
            ln.visuals.append(vis)
        links.append(ln)
    for je in root.findall("joint"):
        R, p = _origin(je.find("origin"))
        ae = je.find("axis")
        axis = _floats(ae.get("xyz") if ae is not None else None, 3, [1, 0, 0])
        lim = je.find("limit")
        lo = float(lim.get("lower", "0")) if lim is not None else 0.0
        hi = float(lim.get("upper", "0")) if lim is not None else 0.0
        joints.append(UrdfJoint(je.get("name"), je.get("type"), je.find("parent").get("link"),
                                je.find("child").get("link"), R, p, axis, lo, hi))
    return links, joints


@dataclass
class Tree:
    """Flat model: index 0 = universe. Arrays are per joint (length njoints)."""
    names: list
    parent: np.ndarray        # int
    jtype: np.ndarray         # JT_*
    axis: np.ndarray          # (njoints,3)
    place_R: np.ndarray       # (njoints,3,3)  joint placement in the parent joint frame
    place_p: np.ndarray       # (njoints,3)
    idx_q: np.ndarray
    idx_v: np.ndarray
    nq: int
    nv: int
    gravity: np.ndarray       # (3,) linear gravity, reference sets (0,0,-9.81)
    dyn_params: np.ndarray    # (njoints,10) merged body params, PINOCCHIO order [m, mc, Ixx,Ixy,Iyy,Ixz,Iyz,Izz]
    frames: dict              # link name -> (joint id, R, p) of the link frame in its joint frame
    lower: np.ndarray         # per-joint position limits (revolute), for synthetic trajectories
    upper: np.ndarray

    @property
    def njoints(self):
        return len(self.names)

    @property
    def nbodies(self):
        return len(self.names) - 1


def _link_dyn_params(link: UrdfLink, R, p):
    """10 dynamic parameters (pinocchio order) of one URDF link expressed in the joint frame,
    the link frame being placed at (R, p) in that joint frame.  Parameters about the joint origin
    are additive, which is how pinocchio's appendBodyToJoint merges fixed children."""
    if link.mass is None:
        return np.zeros(10)
    m = link.mass
    Ri = rpy_to_matrix(*link.rpy)
    com = p + R @ link.com
    Rt = R @ Ri
    I_c = Rt @ link.I_c @ Rt.T
    I_o = I_c + m * skew(com) @ skew(com).T
    return np.array([m, m * com[0], m * com[1], m * com[2],
                     I_o[0, 0], I_o[0, 1], I_o[1, 1], I_o[0, 2], I_o[1, 2], I_o[2, 2]])


def build_tree(urdf_path, floating_base=True, gravity=(0.0, 0.0, -9.81)) -> Tree:
    links, joints = parse_urdf(urdf_path)
    by_name = {l.name: l for l in links}
    children = {l.name: [] for l in links}
    child_names = set()
    for j in joints:
        if j.parent in children and j.child in by_name:
            children[j.parent].append(j)
            child_names.add(j.child)
    roots = [l.name for l in links if l.name not in child_names]
    assert len(roots) == 1, f"URDF must have one root link, found {roots}"

    names = ["universe"]
    parent, jtype, axis, pR, pp, lower, upper = [0], [-1], [np.zeros(3)], [np.eye(3)], [np.zeros(3)], [0.0], [0.0]
    dyn = [np.zeros(10)]
    frames = {}

    def add_joint(name, par, jt, ax, R, p, lo=0.0, hi=0.0):
        names.append(name); parent.append(par); jtype.append(jt); axis.append(np.asarray(ax, float))
        pR.append(R); pp.append(p); dyn.append(np.zeros(10)); lower.append(lo); upper.append(hi)
        return len(names) - 1

    def visit(link_name, jid, R, p):
        frames[link_name] = (jid, R.copy(), p.copy())
        dyn[jid] = dyn[jid] + _link_dyn_params(by_name[link_name], R, p)
        for j in sorted(children[link_name], key=lambda jj: jj.name):
            Rj, pj = R @ j.R, p + R @ j.p
            if j.jtype == "fixed":
                visit(j.child, jid, Rj, pj)
            elif j.jtype in ("revolute", "continuous"):
                if j.jtype == "continuous":
                    raise ValueError("continuous joints (pinocchio RUB*, nq=2) are not on the path")
                if np.array_equal(j.axis, [1, 0, 0]):
                    jt = JT_RX
                elif np.array_equal(j.axis, [0, 1, 0]):
                    jt = JT_RY
                elif np.array_equal(j.axis, [0, 0, 1]):
                    jt = JT_RZ
                else:
                    jt = JT_RU
                ax = j.axis / np.linalg.norm(j.axis)
                nid = add_joint(j.name, jid, jt, ax, Rj, pj, j.lower, j.upper)
                visit(j.child, nid, np.eye(3), np.zeros(3))
            else:
                raise ValueError(f"unsupported joint type {j.jtype!r} ({j.name})")

    if floating_base:
        rid = add_joint("root_joint", 0, JT_FF, np.zeros(3), np.eye(3), np.zeros(3))
        visit(roots[0], rid, np.eye(3), np.zeros(3))
    else:
        visit(roots[0], 0, np.eye(3), np.zeros(3))

    n = len(names)
    idx_q, idx_v = np.zeros(n, int), np.zeros(n, int)
    nq = nv = 0
    for i in range(1, n):
        idx_q[i], idx_v[i] = nq, nv
        if jtype[i] == JT_FF:
            nq += 7; nv += 6
        else:
            nq += 1; nv += 1
    return Tree(names=names, parent=np.array(parent), jtype=np.array(jtype), axis=np.array(axis),
                place_R=np.array(pR), place_p=np.array(pp), idx_q=idx_q, idx_v=idx_v, nq=nq, nv=nv,
                gravity=np.array(gravity, dtype=np.float64), dyn_params=np.array(dyn), frames=frames,
                lower=np.array(lower), upper=np.array(upper))


# ----------------------------------------------------------------------------------------------
# mesh axis-aligned bounding boxes (what trimesh.load_mesh(...).bounding_box gives)
# ----------------------------------------------------------------------------------------------
def mesh_aabb(path):
    """(min(3), max(3)) over all vertices of a binary/ASCII STL or a Wavefront OBJ."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".obj":
        vs = []
        with open(path, "r", errors="ignore") as f:
            for line in f:
                if line.startswith("v "):
                    t = line.split()
                    vs.append((float(t[1]), float(t[2]), float(t[3])))
        v = np.array(vs, dtype=np.float64)
        return v.min(0), v.max(0)
    if ext == ".stl":
        with open(path, "rb") as f:
            data = f.read()
        ntri = struct.unpack_from("<I", data, 80)[0] if len(data) >= 84 else -1
        if len(data) == 84 + 50 * ntri:      # binary STL
            rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                                count=ntri, offset=84)
            v = rec["v"].reshape(-1, 3).astype(np.float64)
            return v.min(0), v.max(0)
        vs = []
        for line in data.decode("ascii", errors="ignore").splitlines():
            t = line.split()
            if len(t) == 4 and t[0] == "vertex":
                vs.append((float(t[1]), float(t[2]), float(t[3])))
        v = np.array(vs, dtype=np.float64)
        return v.min(0), v.max(0)
    raise ValueError(f"unsupported mesh format {path}")


def resolve_mesh_path(filename, files_root, urdf_path):
    """Reference rule (src/sys_identification.py:255-257): <repo>/files/ + filename[10:], i.e. strip
    'package://'.  G1 URDFs use bare 'meshes/x.STL' for which that rule yields garbage (SURVEY.md
    'Reference gaps'): fall back to a path relative to the URDF's own directory."""
    if filename.startswith("package://"):
        return os.path.join(files_root, filename[10:])
    return os.path.join(os.path.dirname(os.path.abspath(urdf_path)), filename)


def bounding_ellipsoids(urdf_path, link_names, files_root, mesh_fallbacks=None):
    """src/sys_identification.py:235-264.  Iterates links in URDF order (not link_names order) and
    appends one ellipsoid per <visual> of each listed link, exactly as the reference does."""
    links, _ = parse_urdf(urdf_path)
    out = []
    for link in links:
        if link.name not in link_names:
            continue
        for vis in link.visuals:
            k = vis["kind"]
            if k == "box":
                semi_axes = vis["size"] / 2
                center = vis["origin_xyz"] if vis["has_origin"] else np.zeros(3)
            elif k == "cylinder":
                semi_axes = np.array([vis["radius"], vis["radius"], vis["length"] / 2])
                center = vis["origin_xyz"] if vis["has_origin"] else np.zeros(3)
            elif k == "sphere":
                semi_axes = np.array([vis["radius"]] * 3)
                center = vis["origin_xyz"] if vis["has_origin"] else np.zeros(3)
            elif k == "mesh":
                mp = resolve_mesh_path(vis["filename"], files_root, urdf_path)
                if not os.path.exists(mp) and mesh_fallbacks and vis["filename"] in mesh_fallbacks:
                    mp = resolve_mesh_path(mesh_fallbacks[vis["filename"]], files_root, urdf_path)
                lo, hi = mesh_aabb(mp)
                semi_axes = (hi - lo) / 2
                center = (hi + lo) / 2 + vis["origin_xyz"]
            else:
                raise ValueError(f"Unsupported geometry type for link {link.name}")
            out.append({"semi_axes": np.asarray(semi_axes, float), "center": np.asarray(center, float)})
    return out


def phi_prior(urdf_path, link_names):
    """src/sys_identification.py:266-322: float32 vector, REFERENCE order per link
    [m, hx, hy, hz, Ixx, Ixy, Ixz, Iyy, Iyz, Izz], inertia about the link (joint) origin, using the
    bare <inertial> of each listed URDF link in URDF order (fixed children ignored, quirk Q9)."""
    links, _ = parse_urdf(urdf_path)
    sel = [l for l in links if l.name in link_names]
    phi = np.zeros(10 * len(link_names), dtype=np.float32)
    for i in range(len(link_names)):
        l = sel[i]
        m, com = l.mass, l.com
        h = m * com
        R = rpy_to_matrix(*l.rpy)
        I_bar = R @ l.I_c @ R.T + m * skew(com) @ skew(com).T
        j = 10 * i
        phi[j] = m
        phi[j + 1:j + 4] = h
        phi[j + 4:j + 7] = I_bar[0, :]
        phi[j + 7:j + 9] = I_bar[1, 1:]
        phi[j + 9] = I_bar[2, 2]
    return phi


def tree_from_flat(flat) -> Tree:
    """Oracle Tree from a product FlatModel / JSON descriptor (duck-typed: the oracle does not import
    product code).  Lets the oracle run where the URDF files are absent (the GPU box)."""
    n = len(flat.parent)
    idx_q, idx_v = np.zeros(n, int), np.zeros(n, int)
    nq = nv = 0
    jt = np.asarray(flat.jtype)
    for i in range(1, n):
        idx_q[i], idx_v[i] = nq, nv
        if jt[i] == JT_FF:
            nq += 7; nv += 6
        else:
            nq += 1; nv += 1
    frames = {name: (int(j), np.eye(3), np.asarray(o, dtype=np.float64))
              for name, j, o in zip(flat.ee_names, flat.ee_joint, flat.ee_offset)}
    dyn = np.asarray(flat.body_params, dtype=np.float64) if flat.body_params is not None else np.zeros((n, 10))
    return Tree(names=list(flat.joint_names), parent=np.asarray(flat.parent, dtype=int), jtype=jt.astype(int),
                axis=np.asarray(flat.axis, dtype=np.float64), place_R=np.asarray(flat.place_R, dtype=np.float64),
                place_p=np.asarray(flat.place_p, dtype=np.float64), idx_q=idx_q, idx_v=idx_v, nq=nq, nv=nv,
                gravity=np.asarray(flat.gravity, dtype=np.float64), dyn_params=dyn, frames=frames,
                lower=np.asarray(flat.lower, dtype=np.float64), upper=np.asarray(flat.upper, dtype=np.float64))
