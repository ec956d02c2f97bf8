"""Flatten the reference's robot descriptions into the small JSON descriptors shipped with the
package (system_identification_b200/robots/*.json), so tests, smoke() and bench.py can run where
/root/reference does not exist (the GPU box).  Run in the build container:

    python tools/make_robot_descriptors.py [/root/reference/files]

Numbers only (tree topology, joint placements, inertial priors, mesh AABBs) -- no reference code.
"""
import os
import sys

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from system_identification_b200.urdf import load_robot  # noqa: E402

FILES = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/files"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "system_identification_b200", "robots")
# body.obj is listed in the reference's .MISSING_LARGE_BLOBS; the collision hull of the same body is present
SPOT_FALLBACK = {"package://spot_description/meshes/base/visual/body.obj":
                 "package://spot_description/meshes/base/collision/body_collision.obj"}
G1_FALLBACK = {"meshes/torso_link_23dof_rev_1_0.STL": "meshes/torso_link.STL"}

jobs = [
    ("solo12", "solo_description/solo12.urdf", os.path.join(FILES, "solo_description/solo12_config.yaml"), None, False),
    ("spot", "spot_description/spot.urdf", os.path.join(FILES, "spot_description/spot_config.yaml"), SPOT_FALLBACK, False),
    ("g1_12dof", "g1_description/g1_12dof.urdf", os.path.join(OUT, "g1_12dof_config.yaml"), G1_FALLBACK, True),
    ("g1_29dof", "g1_description/g1_29dof.urdf", os.path.join(OUT, "g1_29dof_config.yaml"), G1_FALLBACK, True),
    ("g1_29dof_lock_waist", "g1_description/g1_29dof_lock_waist.urdf", os.path.join(OUT, "g1_29dof_config.yaml"), G1_FALLBACK, True),
]
for name, urdf, cfg, fb, merged in jobs:
    with open(cfg) as f:
        conf = yaml.safe_load(f)["robot"]
    m = load_robot(os.path.join(FILES, urdf), conf, files_root=FILES, mesh_fallbacks=fb, merged=merged)
    m.save(os.path.join(OUT, name + ".json"))
    print(name, "njoints", m.njoints, "nq", m.nq, "nv", m.nv, "mass(sum prior)", float(m.phi_prior[0::10].sum()), "ellipsoids", len(m.ellipsoids))
