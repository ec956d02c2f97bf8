"""ORACLE (test infrastructure, not product code) -- stage 1/2 of the hot path in plain numpy fp64.

PARITY UNPINNED against pinocchio itself (not installable here); pinned instead by (i) the identity
Y(q,v,a) pi == RNEA(q,v,a; pi) with an RNEA written independently below, (ii) finite differences of
the foot position for J_c, (iii) projector identities.  See tests/test_oracle_dynamics.py.

Restates:
  * pin.computeJointTorqueRegressor            called at src/sys_identification.py:395,406 (upstream alg.)
  * _update_fk / _compute_J_c / getFrameJacobian(LOCAL_WORLD_ALIGNED)[0:3]   src/sys_identification.py:113-129
  * _compute_null_space_proj  P = I - pinv(J_c) J_c                          src/sys_identification.py:131-135
  * get_proj_regressor_torque / get_proj_friction_regressors                 src/sys_identification.py:401-418
  * demo stacking loops                                                      demo/solo_identification.py:36-55,79-84
  * print_tau_prediction_rmse formulas                                       src/sys_identification.py:421-437
"""
from __future__ import annotations

import numpy as np

from .urdf_tree import JT_FF, JT_RX, JT_RY, JT_RZ, JT_RU, Tree, skew


# ---------------------------------------------------------------- joint transforms
def quat_to_matrix(x, y, z, w):
    """Eigen::Quaternion::toRotationMatrix -- NO normalisation (pinocchio free-flyer uses it as is)."""
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]], dtype=np.float64)


def joint_transform(tree: Tree, i, q):
    jt = tree.jtype[i]
    iq = tree.idx_q[i]
    if jt == JT_FF:
        return quat_to_matrix(q[iq + 3], q[iq + 4], q[iq + 5], q[iq + 6]), np.array(q[iq:iq + 3], dtype=np.float64)
    c, s = np.cos(q[iq]), np.sin(q[iq])
    if jt == JT_RX:
        R = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)
    elif jt == JT_RY:
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)
    elif jt == JT_RZ:
        R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)
    else:  # Rodrigues about a unit axis
        u = tree.axis[i]
        K = skew(u)
        R = np.eye(3) + s * K + (1 - c) * (K @ K)
    return R, np.zeros(3)


def motion_subspace(tree: Tree, i):
    """S_i as a (6, nv_i) matrix, motions ordered [linear; angular]."""
    jt = tree.jtype[i]
    if jt == JT_FF:
        return np.eye(6)
    S = np.zeros((6, 1))
    if jt == JT_RU:
        S[3:6, 0] = tree.axis[i]
    else:
        S[3 + (jt - JT_RX), 0] = 1.0
    return S


def joint_nv(tree, i):
    return 6 if tree.jtype[i] == JT_FF else 1


def act_inv_motion(R, p, m):
    """SE3(R,p).actInv(motion [v; w]) = [R^T (v - p x w); R^T w]."""
    v, w = m[:3], m[3:]
    return np.concatenate([R.T @ (v - np.cross(p, w)), R.T @ w])


def act_force(R, p, f):
    """SE3(R,p).act(force [f; n]) = [R f; R n + p x (R f)]  (column-wise for a 6xk block)."""
    f = np.asarray(f)
    lin = R @ f[:3]
    ang = R @ f[3:] + skew(p) @ lin
    return np.concatenate([lin, ang], axis=0)


def motion_cross(m1, m2):
    """(v,w) x (v2,w2) = (w x v2 + v x w2 ; w x w2)."""
    v, w = m1[:3], m1[3:]
    v2, w2 = m2[:3], m2[3:]
    return np.concatenate([np.cross(w, v2) + np.cross(v, w2), np.cross(w, w2)])


# ---------------------------------------------------------------- regressor
def _Br(u):
    """3x6 matrix with Br(u) @ [Ixx,Ixy,Iyy,Ixz,Iyz,Izz] = I u  (pinocchio parameter order)."""
    return np.array([[u[0], u[1], 0, u[2], 0, 0],
                     [0, u[0], u[1], 0, u[2], 0],
                     [0, 0, 0, u[0], u[1], u[2]]], dtype=np.float64)


def body_regressor(v, a):
    """pinocchio::bodyRegressor(v, a_gf): 6x10, force = B @ [m, mc, Ixx,Ixy,Iyy,Ixz,Iyz,Izz]."""
    vl, w = v[:3], v[3:]
    al, al_ang = a[:3], a[3:]
    acc = al + np.cross(w, vl)
    B = np.zeros((6, 10))
    B[0:3, 0] = acc
    B[0:3, 1:4] = skew(al_ang) + skew(w) @ skew(w)
    B[3:6, 1:4] = -skew(acc)
    B[3:6, 4:10] = _Br(al_ang) + skew(w) @ _Br(w)
    return B


def forward_pass(tree: Tree, q, v, a):
    """liMi, v, a_gf in local joint frames (first loop of computeJointTorqueRegressor)."""
    n = tree.njoints
    liR = np.zeros((n, 3, 3)); lip = np.zeros((n, 3))
    vel = np.zeros((n, 6)); acc = np.zeros((n, 6))
    acc[0, :3] = -tree.gravity
    for i in range(1, n):
        lam = tree.parent[i]
        Rj, pj = joint_transform(tree, i, q)
        liR[i] = tree.place_R[i] @ Rj
        lip[i] = tree.place_p[i] + tree.place_R[i] @ pj
        S = motion_subspace(tree, i)
        iv, nvi = tree.idx_v[i], joint_nv(tree, i)
        vJ = S @ v[iv:iv + nvi]
        vel[i] = vJ
        if lam > 0:
            vel[i] = vel[i] + act_inv_motion(liR[i], lip[i], vel[lam])
        acc[i] = motion_cross(vel[i], vJ) + S @ a[iv:iv + nvi] + act_inv_motion(liR[i], lip[i], acc[lam])
    return liR, lip, vel, acc


def joint_torque_regressor(tree: Tree, q, v, a):
    """Y (nv x 10*nbodies), pinocchio column order per body [m, mcx,mcy,mcz, Ixx,Ixy,Iyy,Ixz,Iyz,Izz]."""
    q = np.asarray(q, dtype=np.float64); v = np.asarray(v, dtype=np.float64); a = np.asarray(a, dtype=np.float64)
    n = tree.njoints
    liR, lip, vel, acc = forward_pass(tree, q, v, a)
    Y = np.zeros((tree.nv, 10 * (n - 1)))
    for i in range(n - 1, 0, -1):
        B = body_regressor(vel[i], acc[i])
        j = i
        while j > 0:
            S = motion_subspace(tree, j)
            iv, nvj = tree.idx_v[j], joint_nv(tree, j)
            Y[iv:iv + nvj, 10 * (i - 1):10 * i] = S.T @ B
            if tree.parent[j] > 0:
                B = act_force(liR[j], lip[j], B)
            j = tree.parent[j]
    return Y


def rnea(tree: Tree, q, v, a, dyn_params=None):
    """Independent check: recursive Newton-Euler with an explicit 6x6 spatial inertia per body
    (parameters in pinocchio order), f = I a + v x* (I v)."""
    q = np.asarray(q, dtype=np.float64); v = np.asarray(v, dtype=np.float64); a = np.asarray(a, dtype=np.float64)
    pi = tree.dyn_params if dyn_params is None else np.asarray(dyn_params).reshape(tree.njoints - 1, 10)
    if dyn_params is not None:
        pi = np.vstack([np.zeros(10), pi])
    n = tree.njoints
    liR, lip, vel, acc = forward_pass(tree, q, v, a)
    f = np.zeros((n, 6))
    for i in range(1, n):
        m, h = pi[i, 0], pi[i, 1:4]
        Ixx, Ixy, Iyy, Ixz, Iyz, Izz = pi[i, 4:10]
        Ib = np.array([[Ixx, Ixy, Ixz], [Ixy, Iyy, Iyz], [Ixz, Iyz, Izz]])
        I6 = np.zeros((6, 6))
        I6[:3, :3] = m * np.eye(3); I6[:3, 3:] = -skew(h); I6[3:, :3] = skew(h); I6[3:, 3:] = Ib
        mom = I6 @ vel[i]
        Ia = I6 @ acc[i]
        vl, w = vel[i, :3], vel[i, 3:]
        f[i] = Ia + np.concatenate([np.cross(w, mom[:3]), np.cross(w, mom[3:]) + np.cross(vl, mom[:3])])
    tau = np.zeros(tree.nv)
    for i in range(n - 1, 0, -1):
        S = motion_subspace(tree, i)
        iv, nvi = tree.idx_v[i], joint_nv(tree, i)
        tau[iv:iv + nvi] = S.T @ f[i]
        lam = tree.parent[i]
        if lam > 0:
            f[lam] += act_force(liR[i], lip[i], f[i])
    return tau


# ---------------------------------------------------------------- contact Jacobian / projector
def world_placements(tree: Tree, q):
    q = np.asarray(q, dtype=np.float64)
    n = tree.njoints
    oR = np.zeros((n, 3, 3)); op = np.zeros((n, 3))
    oR[0] = np.eye(3)
    for i in range(1, n):
        lam = tree.parent[i]
        Rj, pj = joint_transform(tree, i, q)
        lR = tree.place_R[i] @ Rj
        lp = tree.place_p[i] + tree.place_R[i] @ pj
        oR[i] = oR[lam] @ lR
        op[i] = op[lam] + oR[lam] @ lp
    return oR, op


def frame_position(tree: Tree, q, frame_name):
    oR, op = world_placements(tree, q)
    j, _, pf = tree.frames[frame_name]
    return op[j] + oR[j] @ pf


def frame_jacobian_lwa_linear(tree: Tree, q, frame_name):
    """3 x nv: translational rows of getFrameJacobian(model, data, frame, LOCAL_WORLD_ALIGNED)."""
    oR, op = world_placements(tree, q)
    jf, _, pf = tree.frames[frame_name]
    p_f = op[jf] + oR[jf] @ pf
    J = np.zeros((3, tree.nv))
    c = jf
    while c > 0:
        iv = tree.idx_v[c]
        jt = tree.jtype[c]
        if jt == JT_FF:
            J[:, iv:iv + 3] = oR[c]
            J[:, iv + 3:iv + 6] = -skew(p_f - op[c]) @ oR[c]
        else:
            ax = tree.axis[c] if jt == JT_RU else np.eye(3)[jt - JT_RX]
            J[:, iv] = np.cross(oR[c] @ ax, p_f - op[c])
        c = tree.parent[c]
    return J


def contact_jacobian(tree: Tree, q, cnt, ee_frames):
    """src/sys_identification.py:119-129, including quirk Q5: rows allocated from int(sum(cnt)),
    feet selected by truthiness (state 2 counts as stance)."""
    cnt = np.asarray(cnt)
    m = int(np.sum(cnt))
    n_true = int(sum(1 for k in range(len(ee_frames)) if cnt[k]))
    rows = max(m, n_true)   # the reference would raise if m < n_true (e.g. negative labels); never for {0,1,2}
    J_c = np.zeros((3 * rows, tree.nv))
    j = 0
    for k, name in enumerate(ee_frames):
        if cnt[k]:
            J_c[j:j + 3, :] = frame_jacobian_lwa_linear(tree, q, name)
            j += 3
    return J_c


def null_space_projector(tree: Tree, q, cnt, ee_frames):
    """P = I - pinv(J_c) J_c   (src/sys_identification.py:131-135; numpy pinv cutoff 1e-15 sigma_max)."""
    J_c = contact_jacobian(tree, q, cnt, ee_frames)
    if J_c.shape[0] == 0:
        return np.eye(tree.nv)
    return np.eye(tree.nv) - np.linalg.pinv(J_c) @ J_c


def selection_matrix(tree: Tree, floating_base=True):
    base = 6 if floating_base else 0
    d = tree.nv - base
    S = np.zeros((d, tree.nv))
    S[:, base:] = np.eye(d)
    return S


def proj_regressor_torque(tree, q, dq, ddq, tau, cnt, ee_frames, floating_base=True):
    """src/sys_identification.py:401-410."""
    Y = joint_torque_regressor(tree, q, dq, ddq)
    P = null_space_projector(tree, q, cnt, ee_frames)
    S = selection_matrix(tree, floating_base)
    return P @ Y, P @ S.T @ np.asarray(tau, dtype=np.float64)


def proj_friction_regressors(tree, q, dq, ddq, cnt, ee_frames, floating_base=True):
    """src/sys_identification.py:412-418."""
    P = null_space_projector(tree, q, cnt, ee_frames)
    S = selection_matrix(tree, floating_base)
    base = 6 if floating_base else 0
    dqj = np.asarray(dq, dtype=np.float64)[base:]
    return P @ S.T @ np.diag(dqj), P @ S.T @ np.diag(np.sign(dqj))


# ---------------------------------------------------------------- stage 2: stacking and Gram
def stacked_system(tree, q, dq, ddq, tau, cnt, ee_frames, floating_base=True, friction=True):
    """demo/solo_identification.py:36-55,79-84: A = [Y_proj | B_v | B_c] (N*nv x c), b (N*nv)."""
    N = q.shape[1]
    rows_A, rows_b = [], []
    for i in range(N):
        Yp, tp = proj_regressor_torque(tree, q[:, i], dq[:, i], ddq[:, i], tau[:, i], cnt[:, i], ee_frames, floating_base)
        if friction:
            Bv, Bc = proj_friction_regressors(tree, q[:, i], dq[:, i], ddq[:, i], cnt[:, i], ee_frames, floating_base)
            rows_A.append(np.hstack([Yp, Bv, Bc]))
        else:
            rows_A.append(Yp)
        rows_b.append(tp)
    return np.vstack(rows_A), np.hstack(rows_b)


def gram_from_stack(A, b):
    """Sufficient statistics of the least-squares term: G = A^T A, r = A^T b, s = b^T b, n = rows."""
    return A.T @ A, A.T @ b, float(b @ b), A.shape[0]


def tau_prediction_rmse(tree, q, dq, ddq, tau, cnt, phi, ee_frames, floating_base=True):
    """src/sys_identification.py:421-437 (quirk Q7: 'total' is mean squared norm, no root;
    phi multiplies the PINOCCHIO-ordered regressor as is, quirk Q1)."""
    pred, meas = [], []
    for i in range(q.shape[1]):
        y, t = proj_regressor_torque(tree, q[:, i], dq[:, i], ddq[:, i], tau[:, i], cnt[:, i], ee_frames, floating_base)
        pred.append((y @ phi)[6:]); meas.append(t[6:])
    err = np.vstack(pred) - np.vstack(meas)
    return float(np.mean(np.square(np.linalg.norm(err, axis=1)))), np.sqrt(np.mean(np.square(err), axis=0))
