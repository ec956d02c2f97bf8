"""URDF front ends (oracle and product) against each other, against the committed descriptors, against the
reference's only known-answer table (demo/RUN_DEMO.md 'A priori' column), and on a synthetic URDF."""
import os

import numpy as np
import pytest
import yaml

import helpers as H
from oracle import urdf_tree as ut
from system_identification_b200.urdf import UrdfRobot, load_robot, mesh_bounds
from system_identification_b200.sys_identification import MESH_FALLBACKS

HAVE_REF = os.path.isdir(H.REFERENCE_FILES)
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def test_run_demo_prior_known_answer():
    """reference demo/RUN_DEMO.md:12-21, Spot front_left_hip, 'A priori' column (6 decimals)."""
    flat = H.flat_model("spot")
    i = flat.link_names.index("front_left_hip")
    p = flat.phi_prior[10 * i:10 * i + 10].astype(float)
    assert flat.phi_prior.dtype == np.float32                       # quirk Q3
    expect = dict(m=1.68, c=(-0.005374, 0.012842, 0.000099), I=(0.002391, 0.000126, -0.000009, 0.002057, 0.000222, 0.002396))
    assert abs(p[0] - expect["m"]) < 5e-7
    assert np.abs(p[1:4] / p[0] - expect["c"]).max() < 5e-7
    assert np.abs(p[4:10] - expect["I"]).max() < 5e-7


@pytest.mark.parametrize("name", H.ROBOTS)
def test_descriptor_trees(name):
    flat = H.flat_model(name)
    assert flat.njoints == 14 and flat.nq == 19 and flat.nv == 18 and flat.jtype[1] == 0
    expect_parent = {"solo12": [0, 0, 1, 2, 3, 1, 5, 6, 1, 8, 9, 1, 11, 12], "spot": [0, 0, 1, 2, 3, 1, 5, 6, 1, 8, 9, 1, 11, 12],
                     "g1_12dof": [0, 0, 1, 2, 3, 4, 5, 6, 1, 8, 9, 10, 11, 12]}[name]
    assert list(flat.parent) == expect_parent
    masses = {"solo12": 2.501304, "spot": 34.000002, "g1_12dof": 32.106857}
    assert abs(flat.body_params[:, 0].sum() - masses[name]) < 1e-5
    assert len(flat.ellipsoids) == 13 and len(flat.phi_prior) == 130
    back = type(flat).from_json(flat.to_json())
    assert np.array_equal(back.place_R, flat.place_R) and np.array_equal(back.phi_prior, flat.phi_prior)


@pytest.mark.skipif(not HAVE_REF, reason="reference files only exist in the build container")
@pytest.mark.parametrize("name", H.ROBOTS)
def test_product_and_oracle_front_ends_agree_on_reference_urdfs(name):
    urdf, cfg = H.URDFS[name]
    urdf = os.path.join(H.REFERENCE_FILES, urdf)
    t = ut.build_tree(urdf)
    flat = UrdfRobot(urdf).flatten()
    assert list(flat.parent) == list(t.parent) and flat.joint_names == t.names
    assert np.abs(flat.place_R - t.place_R).max() < 1e-15 and np.abs(flat.place_p - t.place_p).max() < 1e-15
    assert np.abs(flat.body_params - t.dyn_params).max() < 1e-12
    shipped = H.flat_model(name)
    assert np.array_equal(shipped.parent, flat.parent) and np.abs(shipped.place_R - flat.place_R).max() < 1e-15
    if cfg is not None:
        conf = yaml.safe_load(open(os.path.join(H.REFERENCE_FILES, cfg)))["robot"]
        full = load_robot(urdf, conf, files_root=H.REFERENCE_FILES, mesh_fallbacks=MESH_FALLBACKS)
        assert np.array_equal(full.phi_prior, ut.phi_prior(urdf, conf["link_names"]))
        ell = ut.bounding_ellipsoids(urdf, conf["link_names"], H.REFERENCE_FILES, mesh_fallbacks=MESH_FALLBACKS)
        for a, b in zip(full.ellipsoids, ell):
            assert np.abs(a["semi_axes"] - b["semi_axes"]).max() < 1e-12 and np.abs(a["center"] - b["center"]).max() < 1e-12
        assert np.array_equal(full.phi_prior, shipped.phi_prior)


def test_synthetic_urdf_fixed_joints_unaligned_axis_and_primitives(tmp_path):
    """Box/cylinder/sphere/mesh visuals, a fixed child folded into its parent, an unaligned axis, ASCII STL."""
    urdf = os.path.join(DATA, "toy_biped.urdf")
    conf = yaml.safe_load(open(os.path.join(DATA, "toy_biped.yaml")))["robot"]
    flat = load_robot(urdf, conf, files_root=DATA)
    t = ut.build_tree(urdf)
    assert flat.joint_names == ["universe", "root_joint", "l_hip", "l_knee", "r_hip", "r_knee"]
    assert list(flat.parent) == [0, 0, 1, 2, 1, 4] == list(t.parent)
    assert list(flat.jtype[1:]) == [0, 2, 4, 2, 1]                  # FF, RY, RU (axis 0 0.6 0.8), RY, RX
    assert np.abs(flat.body_params - t.dyn_params).max() < 1e-14
    # the fixed 'l_foot' (0.2 kg at z=-0.3) is folded into the knee body
    assert abs(flat.body_params[3, 0] - (0.5 + 0.2)) < 1e-15
    assert list(flat.ee_joint) == [3, 5] and np.allclose(flat.ee_offset[0], [0, 0, -0.3])
    ell = {n: e for n, e in zip(conf["link_names"], flat.ellipsoids)}
    assert np.allclose(ell["torso"]["semi_axes"], [0.1, 0.15, 0.2])                       # box size / 2
    assert np.allclose(ell["l_thigh"]["semi_axes"], [0.03, 0.03, 0.15]) and np.allclose(ell["l_thigh"]["center"], [0, 0, -0.15])
    assert np.allclose(ell["l_shank"]["semi_axes"], [0.04, 0.04, 0.04])                   # sphere
    lo, hi = mesh_bounds(os.path.join(DATA, "meshes", "wedge.stl"))
    assert np.allclose(lo, [0, 0, 0]) and np.allclose(hi, [0.1, 0.05, 0.2])
    assert np.allclose(ell["r_thigh"]["semi_axes"], [0.05, 0.025, 0.1])                   # mesh AABB / 2
    assert np.allclose(ell["r_thigh"]["center"], np.array([0.05, 0.025, 0.1]) + [0, 0, -0.2])
    prior = ut.phi_prior(urdf, conf["link_names"])
    assert np.array_equal(prior, flat.phi_prior) and prior.dtype == np.float32


def test_unsupported_geometry_and_joint_raise(tmp_path):
    bad = tmp_path / "bad.urdf"
    bad.write_text('<robot name="b"><link name="a"/><link name="c"/><joint name="j" type="prismatic">'
                   '<parent link="a"/><child link="c"/></joint></robot>')
    with pytest.raises(ValueError):
        UrdfRobot(str(bad)).flatten()
