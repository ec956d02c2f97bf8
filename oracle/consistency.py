"""ORACLE (test infrastructure, not product code) -- the physical-consistency report of the reference,
SystemIdentification.get_physical_consistency (reference src/sys_identification.py:324-389), restated in numpy.

Pinned against the reference's own arithmetic by construction only where pinocchio is concerned: the single pinocchio
call in that function is pin.skew (:351-352), restated here; everything else is numpy and follows the reference line by
line, INCLUDING its float32 containers: the 6x6 spatial inertia, the 4x4 pseudo-inertia, Q and the CoM matrix are
np.float32 arrays (:350,356,363,371), so every entry is rounded to float32 and np.linalg.eigvals runs LAPACK's
single-precision sgeev on them; the bare 3x3 I_bar stays float64 (:345-347).

Per link, in the reference's parameter order phi_i = [m, h_x, h_y, h_z, I_xx, I_xy, I_xz, I_yy, I_yz, I_zz] (:336):
    min eig I_bar, min eig I (6x6), min eig J (4x4), min eig C (4x4), trace(J Q)
"""
from __future__ import annotations

import numpy as np


def _skew(v):
    """pin.skew: [v]x."""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def link_matrices(phi_i, semi_axes, center):
    """The five matrices of one link exactly as the reference builds them (dtype included)."""
    m, hx, hy, hz, Ixx, Ixy, Ixz, Iyy, Iyz, Izz = [phi_i[k] for k in range(10)]
    h = np.array([hx, hy, hz])
    semi_axes = np.asarray(semi_axes); center = np.asarray(center)
    I_bar = np.array([[Ixx, Ixy, Ixz], [Ixy, Iyy, Iyz], [Ixz, Iyz, Izz]])              # :345-347 (float64)
    I6 = np.zeros((6, 6), dtype=np.float32)                                            # :350-354
    I6[0:3, 0:3] = I_bar
    I6[0:3, 3:] = _skew(h)
    I6[3:, 0:3] = _skew(h).T
    I6[3:, 3:] = m * np.eye(3)
    J = np.zeros((4, 4), dtype=np.float32)                                             # :357-361
    J[:3, :3] = (1 / 2) * np.trace(I_bar) * np.eye(3) - I_bar
    J[:3, 3] = h
    J[3, :3] = h
    J[3, 3] = m
    Qf = np.zeros((4, 4), dtype=np.float32)                                            # :364-369
    Q = np.linalg.inv(np.diag(semi_axes) ** 2)
    Qf[:3, :3] = Q
    Qf[:3, 3] = Q @ center
    Qf[3, :3] = Q @ center
    Qf[3, 3] = 1 - (center @ Q @ center)
    Cm = np.zeros((4, 4), dtype=np.float32)                                            # :372-376
    Cm[0, 0] = m
    Cm[0, 1:] = h - m * center
    Cm[1:, 0] = h - m * center
    Cm[1:, 1:] = m * np.diag(semi_axes) ** 2
    return I_bar, I6, J, Cm, Qf


def physical_consistency(phi, bounding_ellipsoids):
    """Five lists (one entry per link), the return value of the reference function (:389)."""
    out = ([], [], [], [], [])
    for idx, ell in enumerate(bounding_ellipsoids):
        I_bar, I6, J, Cm, Qf = link_matrices(np.asarray(phi)[10 * idx:10 * idx + 10], ell["semi_axes"], ell["center"])
        out[0].append(np.min(np.linalg.eigvals(I_bar)))                                # :379-383
        out[1].append(np.min(np.linalg.eigvals(I6)))
        out[2].append(np.min(np.linalg.eigvals(J)))
        out[3].append(np.min(np.linalg.eigvals(Cm)))
        out[4].append(np.trace(J @ Qf))
    return out
