"""Host-side mirror of the reference's SystemIdentification class
(reference src/sys_identification.py:10-490): same constructor, getters, per-sample producers and
printers, so demo/solo_identification.py, demo/spot_identification.py and spot_identification.py
run unmodified against `src.sys_identification` (a re-export of this module).

What changes underneath: no pinocchio/urdf_parser_py/trimesh.  The URDF is flattened once on the
host (urdf.py) and uploaded through the C-ABI; every regressor / projector / RMSE evaluation runs in
the sm_100a kernels of libsysid_b200.so.  There is no CPU fallback for those.

New, batched entry points (not in the reference): `identify`, `gram`, `tau_prediction_rmse`.
"""
from __future__ import annotations

import os

import numpy as np
import yaml

from .model import FlatModel
from .urdf import load_robot

# body.obj / torso_link_23dof_rev_1_0.STL are absent from the reference checkout (.MISSING_LARGE_BLOBS)
MESH_FALLBACKS = {
    "package://spot_description/meshes/base/visual/body.obj": "package://spot_description/meshes/base/collision/body_collision.obj",
    "meshes/torso_link_23dof_rev_1_0.STL": "meshes/torso_link.STL",
}


class SystemIdentification(object):
    def __init__(self, urdf_file, config_file, floating_base):
        self._urdf_path = urdf_file
        self._floating_base = floating_base
        with open(config_file, "r") as file:
            config = yaml.safe_load(file)
        robot_config = config.get("robot", {})
        # reference rule for mesh paths: <repo>/files/<package path> (src/sys_identification.py:255-257);
        # here <repo>/files is the grand-parent directory of the URDF, which is the same place for the demos
        files_root = os.path.dirname(os.path.dirname(os.path.abspath(urdf_file)))
        flat = load_robot(urdf_file, robot_config, floating_base=floating_base, files_root=files_root,
                          mesh_fallbacks=MESH_FALLBACKS)
        self._init_from_flat(flat)

    @classmethod
    def from_flat_model(cls, flat: FlatModel):
        """Build from a flattened descriptor (system_identification_b200/robots/*.json) instead of URDF+YAML."""
        self = cls.__new__(cls)
        self._urdf_path = None
        self._floating_base = flat.floating_base
        self._init_from_flat(flat)
        return self

    def _init_from_flat(self, flat: FlatModel):
        self._flat = flat
        self._device_model = None
        self.nq = flat.nq
        self.nv = flat.nv
        self._base_dof = 6
        self.joints_dof = self.nv - self._base_dof
        self._S = np.zeros((self.joints_dof, self.nv))
        self._S[:, self._base_dof:] = np.eye(self.joints_dof)
        self._robot_name = flat.name
        self._robot_mass = flat.robot_mass
        self._link_names = list(flat.link_names)
        self._end_eff_frame_names = list(flat.ee_names)
        self._nb_ee = len(self._end_eff_frame_names)
        self._num_inertial_params = 10
        self._num_links = len(self._link_names)
        if self._num_links != flat.nbodies:
            raise ValueError(f"len(link_names) = {self._num_links} must equal the number of moving bodies {flat.nbodies}")
        self._phi_prior = np.zeros((self._num_inertial_params * self._num_links), dtype=np.float32)
        self.B_v = np.eye(self.joints_dof)
        self.B_c = np.eye(self.joints_dof)
        self._bounding_ellipsoids = list(flat.ellipsoids)
        self._last_key = None
        self._last_tau_key = None
        self._last_out = None
        self._block = None

    # ------------------------------------------------------------------ device plumbing
    @property
    def flat_model(self):
        return self._flat

    @property
    def device_model(self):
        if self._device_model is None:
            from .ops import DeviceModel
            self._device_model = DeviceModel(self._flat)
        return self._device_model

    # The reference demos call the two per-sample producers inside Python loops over q[:, i] (demo/solo_identification.py:36-55):
    # N launches of one sample each would be all latency.  When the five vectors are column views of 2-D arrays (which is what
    # q[:, i] is), the whole block of COMPAT_CHUNK columns around i is computed in ONE sysid_projected_batch launch on the
    # parent arrays and the following calls are served from it -- after checking that the bytes handed in are the bytes the
    # block was computed from.  Anything else (plain vectors, lists) takes the single-sample launch.
    COMPAT_CHUNK = 4096

    @staticmethod
    def _column_of(v):
        """(parent 2-D array, column index) when v is a column view parent[:, i]; None otherwise."""
        if not isinstance(v, np.ndarray) or v.ndim != 1:
            return None
        b = v.base
        if not isinstance(b, np.ndarray) or b.ndim != 2 or b.shape[0] != v.shape[0] or b.shape[1] == 0:
            return None
        if v.shape[0] > 1 and v.strides[0] != b.strides[0]:
            return None
        off = v.__array_interface__["data"][0] - b.__array_interface__["data"][0]
        if b.strides[1] <= 0 or off < 0 or off % b.strides[1] != 0:
            return None
        i = off // b.strides[1]
        return (b, int(i)) if i < b.shape[1] else None

    def _block_for(self, vecs):
        """vecs = (q, dq, ddq, tau or None, cnt) column views -> (block dict, local index) or None."""
        loc = [None if v is None else self._column_of(v) for v in vecs]
        if any(l is None for l, v in zip(loc, vecs) if v is not None):
            return None
        idx = {l[1] for l in loc if l is not None}
        N = {l[0].shape[1] for l in loc if l is not None}
        if len(idx) != 1 or len(N) != 1:
            return None
        i, N = idx.pop(), N.pop()
        ids = tuple(None if l is None else (id(l[0]), l[0].__array_interface__["data"][0], l[0].shape, l[0].dtype.str) for l in loc)
        blk = self._block
        lo = (i // self.COMPAT_CHUNK) * self.COMPAT_CHUNK
        # the friction blocks do not depend on tau: a block computed with torques serves a call without them
        same = blk is not None and blk["lo"] == lo and all(a == b_ for k, (a, b_) in enumerate(zip(blk["ids"], ids)) if not (k == 3 and b_ is None))
        if not same:
            from .ops import to_device
            hi = min(N, lo + self.COMPAT_CHUNK)
            host = [np.zeros((self.joints_dof, hi - lo)) if l is None else np.ascontiguousarray(l[0][:, lo:hi], dtype=np.float64) for l in loc]
            A, b = self.device_model.projected_batch(*(to_device(a) for a in host), friction=True)
            blk = {"ids": ids, "lo": lo, "host": host, "A": A.cpu().numpy(), "b": b.cpu().numpy()}
            self._block = blk
        k = i - lo
        # the served bytes must be the bytes that were asked for (a parent array modified in place invalidates the block)
        for v, h in zip(vecs, blk["host"]):
            if v is not None and not np.array_equal(np.asarray(v, dtype=np.float64), h[:, k], equal_nan=True):
                self._block = None
                return None
        return blk, k

    def _one_sample(self, q, dq, ddq, tau, cnt):
        """Per-sample compat path: (A (nv, c), b (nv)) of one sample; served from a block launch when the arguments are
        column views of the log (see above), else one launch of the projected-batch kernel with N = 1."""
        from .ops import to_device
        q = np.asarray(q); dq = np.asarray(dq); ddq = np.asarray(ddq); cnt = np.asarray(cnt)
        tau_in = None if tau is None else np.asarray(tau)
        hit = self._block_for((q, dq, ddq, tau_in, cnt))
        if hit is not None:
            blk, k = hit
            return blk["A"][k], blk["b"][k]
        tau = np.zeros(self.joints_dof) if tau_in is None else tau_in
        key = (q.tobytes(), dq.tobytes(), ddq.tobytes(), cnt.tobytes())
        if self._last_key == key and self._last_tau_key == tau.tobytes():
            return self._last_out
        dm = self.device_model
        packed = np.concatenate([q.astype(np.float64), dq.astype(np.float64), ddq.astype(np.float64),
                                 tau.astype(np.float64), cnt.astype(np.float64)])
        dev = to_device(packed.reshape(-1, 1))
        o = 0
        parts = []
        for n in (self.nq, self.nv, self.nv, self.joints_dof, self._nb_ee):
            parts.append(dev[o:o + n]); o += n
        A, b = dm.projected_batch(*parts, friction=True)
        out = (A[0].cpu().numpy(), b[0].cpu().numpy())
        self._last_key, self._last_tau_key, self._last_out = key, tau.tobytes(), out
        return out

    # ------------------------------------------------------------------ reference getters
    def get_robot_mass(self):
        return self._robot_mass

    def get_num_links(self):
        return self._num_links

    def get_bounding_ellipsoids(self):
        return self._bounding_ellipsoids

    def get_phi_prior(self):
        # float32 vector in the reference's per-link order [m, h_x, h_y, h_z, I_xx, I_xy, I_xz, I_yy, I_yz, I_zz]
        self._phi_prior[:] = self._flat.phi_prior
        return self._phi_prior

    def get_physical_consistency(self, phi):
        """Minimum eigenvalues of I_bar, the 6x6 spatial inertia, the 4x4 pseudo inertia and the CoM matrix, and
        tr(J Q), per link (reference src/sys_identification.py:324-389; float32 matrices as there)."""
        out = ([], [], [], [], [])
        for idx in range(self._num_links):
            p = np.asarray(phi[10 * idx:10 * idx + 10])
            m, h = p[0], np.array(p[1:4])
            I_bar = np.array([[p[4], p[5], p[6]], [p[5], p[7], p[8]], [p[6], p[8], p[9]]])
            ell = self._bounding_ellipsoids[idx]
            s, c = np.asarray(ell["semi_axes"]), np.asarray(ell["center"])
            hx = np.array([[0, -h[2], h[1]], [h[2], 0, -h[0]], [-h[1], h[0], 0]])
            I6 = np.zeros((6, 6), dtype=np.float32)
            I6[0:3, 0:3] = I_bar; I6[0:3, 3:] = hx; I6[3:, 0:3] = hx.T; I6[3:, 3:] = m * np.eye(3)
            J = np.zeros((4, 4), dtype=np.float32)
            J[:3, :3] = 0.5 * np.trace(I_bar) * np.eye(3) - I_bar; J[:3, 3] = h; J[3, :3] = h; J[3, 3] = m
            Qd = np.linalg.inv(np.diag(s) ** 2)
            Qf = np.zeros((4, 4), dtype=np.float32)
            Qf[:3, :3] = Qd; Qf[:3, 3] = Qd @ c; Qf[3, :3] = Qd @ c; Qf[3, 3] = 1 - c @ Qd @ c
            Cm = np.zeros((4, 4), dtype=np.float32)
            Cm[0, 0] = m; Cm[0, 1:] = h - m * c; Cm[1:, 0] = h - m * c; Cm[1:, 1:] = m * np.diag(s) ** 2
            out[0].append(np.min(np.linalg.eigvals(I_bar)))
            out[1].append(np.min(np.linalg.eigvals(I6)))
            out[2].append(np.min(np.linalg.eigvals(J)))
            out[3].append(np.min(np.linalg.eigvals(Cm)))
            out[4].append(np.trace(J @ Qf))
        return out

    def get_physical_consistency_batch(self, phis):
        """get_physical_consistency for many parameter vectors at once (e.g. bootstrap resamples) on the device:
        phis (batch, 10 L) -> numpy (batch, 5, L) = [min eig I_bar, min eig I, min eig J, min eig C, tr(J Q)] per link."""
        import torch
        from .ops import physical_consistency
        t = phis if isinstance(phis, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(phis, dtype=np.float64)))
        t = t.to(device="cuda", dtype=torch.float64)
        return physical_consistency(t.contiguous(), self._num_links, self._bounding_ellipsoids).cpu().numpy()

    def get_full_regressor_force(self, q, dq, ddq, tau, ee_force, cnt):
        raise NotImplementedError(
            "get_full_regressor_force (reference src/sys_identification.py:391-399) has no caller in the reference and "
            "needs measured foot forces that no .dat layout carries; it is outside the accelerated path (SURVEY.md #11)")

    # ------------------------------------------------------------------ per-sample producers (reference :401-418)
    def get_proj_regressor_torque(self, q, dq, ddq, tau, cnt):
        A, b = self._one_sample(q, dq, ddq, tau, cnt)
        return A[:, :self._num_inertial_params * self._num_links].copy(), b.copy()

    def get_proj_friction_regressors(self, q, dq, ddq, cnt):
        # the friction blocks do not depend on tau: reuse the launch of the matching regressor call if there was one
        key = (np.asarray(q).tobytes(), np.asarray(dq).tobytes(), np.asarray(ddq).tobytes(), np.asarray(cnt).tobytes())
        if self._block is None and self._last_out is not None and key == self._last_key:
            A = self._last_out[0]
        else:
            A = self._one_sample(q, dq, ddq, None, cnt)[0]
        p, d = self._num_inertial_params * self._num_links, self.joints_dof
        return A[:, p:p + d].copy(), A[:, p + d:p + 2 * d].copy()

    # ------------------------------------------------------------------ batched entry points (new)
    def _upload(self, q, dq, ddq, tau, cnt):
        import torch
        from .ops import to_device
        return tuple(to_device(a if isinstance(a, torch.Tensor) else np.asarray(a)) for a in (q, dq, ddq, tau, cnt))

    def gram(self, q, dq, ddq, tau, cnt, friction=True, weights=None):
        """Fused regressor+projector+Gram over all columns of the five (channels x N) arrays -> device stats tensor."""
        dev = self._upload(q, dq, ddq, tau, cnt)
        return self.device_model.gram_accumulate(*dev, friction=friction, weights=weights)

    def tau_prediction_rmse(self, q, dq, ddq, torque, cnt, phi):
        """(total, per-joint) with the reference's formulas: total = mean_i ||e_i||^2 (no root), per joint = RMSE."""
        import torch
        dev = self._upload(q, dq, ddq, torque, cnt)
        out = self.device_model.predict_rmse(*dev, torch.as_tensor(np.asarray(phi, dtype=np.float64))).cpu().numpy()
        return float(out[0]), out[1:].copy()

    def identify(self, q, dq, ddq, tau, cnt, lambda_reg=1e-1, tol=1e-10, max_iters=1000, reg_type="constant_pullback",
                 friction=True, return_info=False):
        """End-to-end identification of one log: what the demo scripts' main() computes between read_data and the
        printers (reference demo/solo_identification.py:67-88), in two kernel launches + one solve."""
        from .identify import identify as _identify
        return _identify(self, q, dq, ddq, tau, cnt, lambda_reg=lambda_reg, tol=tol, max_iters=max_iters,
                         reg_type=reg_type, friction=friction, return_info=return_info)

    # ------------------------------------------------------------------ printers (reference :421-490)
    def print_tau_prediction_rmse(self, q, dq, ddq, torque, cnt, phi, param_name):
        rmse_total, joint_tau_rmse = self.tau_prediction_rmse(q, dq, ddq, torque, cnt, phi)
        print("\n--------------------Torque Prediction Errors--------------------")
        print(f'RMSE for joint torques prediction using {param_name} parameters: total= {rmse_total}\nper_joints={joint_tau_rmse}')

    _ROWS = (("mass (kg)", None), ("c_x (m)", 0), ("c_y (m)", 1), ("c_z (m)", 2),
             ("I_xx (kg.m^2)", 0), ("I_xy (kg.m^2)", 1), ("I_xz (kg.m^2)", 2),
             ("I_yy (kg.m^2)", 3), ("I_yz (kg.m^2)", 4), ("I_zz (kg.m^2)", 5))

    def print_inertial_params(self, prior, identified):
        self._cell_width = 13
        w = self._cell_width
        totals = [0, 0]
        header = "|" + "|".join(f"{t:<{w}}" for t in ("Parameter", "A priori", "Identified", "Change", "error %")) + "|"
        for i in range(self._num_links):
            title = f'Inertial Parameters of "{self._link_names[i]}"'
            left = (69 - len(title)) // 2
            print(f'\n{"-" * left} {title} {"-" * (69 - len(title) - left)}')
            print(header)
            k = 10 * i
            masses = (prior[k], identified[k])
            coms = (prior[k + 1:k + 4] / masses[0], identified[k + 1:k + 4] / masses[1])
            inert = (prior[k + 4:k + 10], identified[k + 4:k + 10])
            for r, (label, sel) in enumerate(self._ROWS):
                src = masses if r == 0 else (coms if r < 4 else inert)
                a, b = (src[0], src[1]) if sel is None else (src[0][sel], src[1][sel])
                self._print_table(label, a, b)
            totals[0] += masses[0]
            totals[1] += masses[1]
        print(f'\nRobot total mass: {totals[0]} ---- Identified total mass: {totals[1]}')

    def _print_table(self, description, prior, ident):
        precision = 6
        w = self._cell_width
        change = ident - prior
        error = np.divide(change, np.abs(prior), where=prior != 0) * 100
        error = np.where(np.abs(prior) <= 1e-8, np.nan, error)
        print(f'|{description:<{w}}|{prior:>{w}.{precision}f}|{ident:>{w}.{precision}f}|{change:>{w}.{precision}f}|{error:>{w}.{1}f}|')
