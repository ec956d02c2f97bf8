"""Synthetic logs of the shapes the reference's read_data returns
(reference demo/solo_identification.py:9-33): five channel-major arrays
q (nq, N), dq (nv, N), ddq (nv, N), tau (d, N), contact (n_ee, N).

The reference ships no data (data/ and *.dat are git-ignored), so every benchmark and test input
is generated here; seeds per BASELINE config are fixed in SEEDS.
"""
from __future__ import annotations

import numpy as np

from .model import FlatModel

SEEDS = {"solo12": 1001, "spot": 1002, "g1_12dof": 1003, "g1_1m": 1004, "solo12_bootstrap": 1005}


def _rpy_R(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.empty(r.shape + (3, 3))
    R[..., 0, 0] = cy * cp; R[..., 0, 1] = cy * sp * sr - sy * cr; R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp; R[..., 1, 1] = sy * sp * sr + cy * cr; R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp; R[..., 2, 1] = cp * sr; R[..., 2, 2] = cp * cr
    return R


def _quat_from_R(R):
    """(x, y, z, w) unit quaternion, w >= 0 branch (smooth for the +-0.3 rad motions used here)."""
    w = 0.5 * np.sqrt(np.maximum(0.0, 1.0 + R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]))
    x = (R[..., 2, 1] - R[..., 1, 2]) / (4 * w)
    y = (R[..., 0, 2] - R[..., 2, 0]) / (4 * w)
    z = (R[..., 1, 0] - R[..., 0, 1]) / (4 * w)
    return np.stack([x, y, z, w], axis=0)


class _Sines:
    """sum_k a_k sin(2 pi f_k t + p_k) per channel, with analytic derivatives."""

    def __init__(self, rng, channels, amp, nterms=5, fmin=0.1, fmax=3.0):
        self.f = rng.uniform(fmin, fmax, size=(channels, nterms))
        self.ph = rng.uniform(0, 2 * np.pi, size=(channels, nterms))
        w = rng.uniform(0.5, 1.0, size=(channels, nterms))
        self.a = np.asarray(amp, dtype=np.float64).reshape(-1, 1) * w / w.sum(axis=1, keepdims=True)

    def __call__(self, t, order=0):
        om = 2 * np.pi * self.f[:, :, None]
        arg = om * t[None, None, :] + self.ph[:, :, None]
        if order == 0:
            v = np.sin(arg)
        elif order == 1:
            v = om * np.cos(arg)
        else:
            v = -om * om * np.sin(arg)
        return (self.a[:, :, None] * v).sum(axis=1)


def contact_schedule(model: FlatModel, N, rate_hz, rng):
    """Quadrupeds: trot (diagonal pairs alternate every 0.3 s) with ~20 % four-foot stance and ~5 % flight,
    values in {0,1}.  Bipeds (G1): alternating single support with double-support phases and the
    occasional state 2 ('contact lost', which the reference counts as stance -- quirk Q5)."""
    n_ee = model.n_ee
    cnt = np.zeros((n_ee, N), dtype=np.float64)
    seg = max(1, int(round(0.3 * rate_hz)))
    nseg = (N + seg - 1) // seg
    u = rng.uniform(size=nseg)
    for s in range(nseg):
        sl = slice(s * seg, min(N, (s + 1) * seg))
        if n_ee == 4:
            if u[s] < 0.20:
                cnt[:, sl] = 1
            elif u[s] < 0.25:
                pass
            elif s % 2 == 0:
                cnt[[0, 3], sl] = 1
            else:
                cnt[[1, 2], sl] = 1
        else:
            if u[s] < 0.30:
                cnt[:, sl] = 1
            elif u[s] < 0.35:
                cnt[:, sl] = 1
                cnt[s % n_ee, sl] = 2
            else:
                cnt[s % n_ee, sl] = 1
    return cnt


def make_trajectory(model: FlatModel, N, seed, rate_hz=500.0, chunk=65536):
    """Smooth excitation: joints = 5 sinusoids about mid-range, amplitude 0.3 x limit range; base
    position +-0.1 m, base rpy +-0.3 rad -> unit quaternion; base twist / acceleration in the LOCAL
    frame (free-flyer convention).  q and contact are rounded through float32 (quirk Q8)."""
    rng = np.random.default_rng(seed)
    d = model.joints_dof
    lo, hi = model.lower[2:], model.upper[2:]
    rng_ok = hi > lo
    mid = np.where(rng_ok, 0.5 * (lo + hi), 0.0)
    amp = np.where(rng_ok, 0.3 * (hi - lo), 0.5)
    amp = np.minimum(amp, 1.0)
    joints = _Sines(rng, d, amp)
    pos = _Sines(rng, 3, [0.1, 0.1, 0.1], fmax=1.5)
    rpy = _Sines(rng, 3, [0.3, 0.3, 0.3], fmax=1.5)
    cnt = contact_schedule(model, N, rate_hz, rng)

    q = np.empty((model.nq, N)); dq = np.empty((model.nv, N)); ddq = np.empty((model.nv, N))
    h = 1e-4

    def local_twist(t):
        e, de = rpy(t), rpy(t, 1)
        R = _rpy_R(e[0], e[1], e[2])                                # (n,3,3)
        v_world = pos(t, 1)
        v_loc = np.einsum("nji,jn->in", R, v_world)
        # world angular velocity from rpy rates (R = Rz Ry Rx)
        cy, sy, cp, sp = np.cos(e[2]), np.sin(e[2]), np.cos(e[1]), np.sin(e[1])
        wx = cy * cp * de[0] - sy * de[1]
        wy = sy * cp * de[0] + cy * de[1]
        wz = -sp * de[0] + de[2]
        w_loc = np.einsum("nji,jn->in", R, np.stack([wx, wy, wz]))
        return R, np.concatenate([v_loc, w_loc], axis=0)

    for s in range(0, N, chunk):
        t = np.arange(s, min(N, s + chunk)) / rate_hz
        sl = slice(s, s + t.size)
        R, tw = local_twist(t)
        _, twp = local_twist(t + h)
        _, twm = local_twist(t - h)
        q[0:3, sl] = pos(t) + np.array([[0.0], [0.0], [0.35]])
        q[3:7, sl] = _quat_from_R(R)
        q[7:, sl] = mid[:, None] + joints(t)
        dq[0:6, sl] = tw
        dq[6:, sl] = joints(t, 1)
        ddq[0:6, sl] = (twp - twm) / (2 * h)
        ddq[6:, sl] = joints(t, 2)
    q = q.astype(np.float32).astype(np.float64)
    cnt = cnt.astype(np.float32).astype(np.float64)
    return q, dq, ddq, cnt


def synth_tau(model: FlatModel, N, seed, scale=1.0):
    """Placeholder joint torques (smooth + noise) for pure-throughput runs where any finite values do."""
    rng = np.random.default_rng(seed + 7919)
    t = np.arange(N) / 500.0
    s = _Sines(rng, model.joints_dof, np.full(model.joints_dof, 2.0 * scale))
    return s(t) + 0.05 * scale * rng.standard_normal((model.joints_dof, N))


def torques_from_truth(model: FlatModel, Y, P, dq, phi_true_pin, b_v, b_c, noise, seed):
    """tau such that P S^T tau = P (Y phi_true + friction) + noise: given per-sample regressors
    Y (N, nv, p) and projectors P (N, nv, nv) from ANY implementation, solve the joint rows in the
    least-squares sense per sample.  Used by tests/bench to build identifiable problems."""
    rng = np.random.default_rng(seed + 104729)
    N = Y.shape[0]
    base = model.base_dof
    F = Y @ phi_true_pin                                   # (N, nv) generalised force the motion needs
    dqj = dq[base:, :].T
    tau = np.empty((model.joints_dof, N))
    for i in range(N):
        PS = P[i][:, base:]                                # P S^T
        rhs = P[i] @ F[i]
        tau[:, i] = np.linalg.lstsq(PS, rhs, rcond=None)[0]
    tau += (b_v[:, None] * dqj.T + b_c[:, None] * np.sign(dqj.T))
    tau += noise * rng.standard_normal(tau.shape)
    return tau


def identifiable_tau_device(flat: FlatModel, dm, dev, seed, perturb=0.05, bv_max=0.2, bc_max=0.5, noise=0.5, chunk=32768):
    """Device version of torques_from_truth for full-size logs (20 k .. 1 M samples): tau = the least-squares joint
    torques reproducing P (Y phi_true) + friction + noise, built chunk by chunk from the library's own per-sample
    operators and torch library calls.  Data generation only -- never inside a timed region.
    dm: ops.DeviceModel; dev: the five CUDA arrays (tau is only used for its shape).  Returns a CUDA (d, N) tensor."""
    import torch
    q, dq, ddq, tau, cnt = dev
    N = q.shape[1]
    rng = np.random.default_rng(seed)
    phi_true = torch.from_numpy(flat.body_params[1:].reshape(-1) * (1 + perturb * rng.standard_normal(10 * flat.nbodies))).cuda()
    bv = torch.from_numpy(rng.uniform(0, bv_max, flat.joints_dof)).cuda()
    bc = torch.from_numpy(rng.uniform(0, bc_max, flat.joints_dof)).cuda()
    out = torch.empty((flat.joints_dof, N), dtype=torch.float64, device=q.device)
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        sl = [a[:, lo:hi].contiguous() for a in (q, dq, ddq, tau, cnt)]
        Y = dm.regressor_batch(*sl[:3])
        _, _, P = dm.projected_batch(*sl, want_P=True)
        F = torch.einsum("nrc,c->nr", Y, phi_true)
        rhs = torch.einsum("nrk,nk->nr", P, F)
        PS = P[:, :, 6:]
        # minimum-norm least squares; P S^T has rank 18 - rank(J_c) - (0..6) only, so the cutoff must sit well above the
        # rounding noise of its zero singular values (1e-15 would let them through and produce torques of 1e8 N m)
        sol = torch.einsum("ndk,nk->nd", torch.linalg.pinv(PS, rtol=1e-9), rhs)
        dqj = sl[1][6:, :].T
        t = sol + bv * dqj + bc * torch.sign(dqj) + noise * torch.randn(sol.shape, generator=gen, device="cuda", dtype=torch.float64)
        out[:, lo:hi] = t.T
        del Y, P, F, rhs, PS, sol
    return out
