"""Generate the committed golden fixtures (tests/golden/*.npz) from the ORACLE on seeded synthetic logs.

    python tests/golden/make_golden.py

The reference ships no test vectors of its own; the only in-repo known answer is the
'A priori' column of demo/RUN_DEMO.md, which tests/test_oracle_urdf.py checks directly.  These fixtures pin the
oracle's outputs so that (a) a regression in the oracle is caught on CPU and (b) the CUDA path is compared on the
GPU box against numbers that were produced in the build container.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import dynamics as dy, sdp  # noqa: E402
import helpers as H  # noqa: E402

N = 48
for name in H.ROBOTS:
    flat, data0 = H.small_log(name, N)
    t = H.oracle_tree(flat)
    q, dq, ddq, _, cnt = data0

    def regress(q, dq, ddq, cnt):
        Y = np.array([dy.joint_torque_regressor(t, q[:, i], dq[:, i], ddq[:, i]) for i in range(q.shape[1])])
        P = np.array([dy.null_space_projector(t, q[:, i], cnt[:, i], flat.ee_names) for i in range(q.shape[1])])
        return Y, P

    # a longer identifiable log for the SDP fixture (statistics only are stored)
    data_id, phi_true, bv, bc = H.identifiable_log(flat, 400, 77, regress)
    A_id, b_id = dy.stacked_system(t, *data_id, flat.ee_names)
    G, r, s, n = dy.gram_from_stack(A_id, b_id)
    prob = sdp.build_problem(G, r, s, n, flat.nbodies, flat.phi_prior, flat.robot_mass, flat.ellipsoids, flat.joints_dof)
    x0 = np.concatenate([flat.phi_prior.astype(float), np.ones(2 * flat.joints_dof)])
    xb, ib = sdp.solve_barrier(prob, x0)
    xa, ia = sdp.solve_alm(prob)
    # the small log: per-sample blocks
    tau = data_id[3][:, :N]
    data = (q, dq, ddq, tau, cnt)
    Y, P, A, b = H.oracle_blocks(flat, data)
    Gs, rs, ss, ns = dy.gram_from_stack(A, b)
    tot, per_joint = dy.tau_prediction_rmse(t, *data, flat.phi_prior.astype(float), flat.ee_names)
    out = os.path.join(H.GOLDEN_DIR, f"{name}_N{N}.npz")
    np.savez_compressed(out, q=q, dq=dq, ddq=ddq, tau=tau, cnt=cnt, Y=Y[:6], P=P, A_first=A[:flat.nv * 4], b=b,
                        G=Gs, r=rs, s=ss, n=ns, rmse_total=tot, rmse_joint=per_joint,
                        sdp_G=G, sdp_r=r, sdp_s=s, sdp_n=n, sdp_x_barrier=xb, sdp_x_alm=xa,
                        sdp_obj=ia["objective"], sdp_barrier_gap=ib["gap"], sdp_alm_kkt=ia["kkt"])
    print(name, out, "barrier-vs-alm", H.rel(xb, xa), os.path.getsize(out) // 1024, "KiB")
