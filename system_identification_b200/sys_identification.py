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
        self._fixed = not flat.floating_base
        if self._fixed:
            # reference src/sys_identification.py:15-18,34-37: nq = nv = number of joints, S = I.  The kernels run the free-flyer
            # model with the base pinned (urdf.py::flatten); inputs are widened and outputs sliced below.
            self.joints_dof = flat.nv - 6
            self.nq = self.nv = self.joints_dof
            self._base_dof = 0
            self._S = np.eye(self.joints_dof)
        else:
            self.nq = flat.nq
            self.nv = flat.nv
            self._base_dof = 6
            self.joints_dof = self.nv - self._base_dof
            self._S = np.zeros((self.joints_dof, self.nv))
            self._S[:, self._base_dof:] = np.eye(self.joints_dof)
        self._robot_name = flat.name
        self._robot_mass = flat.robot_mass
        self._link_names = list(flat.link_names)
        self._n_anchor = 3 if self._fixed else 0               # virtual contacts that pin the base (always in stance)
        self._end_eff_frame_names = list(flat.ee_names)[self._n_anchor:]
        self._nb_ee = len(self._end_eff_frame_names)
        self._num_inertial_params = 10
        self._num_links = len(self._link_names)
        if self._num_links != flat.nbodies - (1 if self._fixed else 0):
            raise ValueError(f"len(link_names) = {self._num_links} must equal the number of moving bodies {flat.nbodies - (1 if self._fixed else 0)}")
        self._phi_prior = np.zeros((self._num_inertial_params * self._num_links), dtype=np.float32)
        self.B_v = np.eye(self.joints_dof)
        self.B_c = np.eye(self.joints_dof)
        self._bounding_ellipsoids = list(flat.ellipsoids)
        self._last_key = None
        self._last_tau_key = None
        self._last_out = None
        self._block = None
        self._prev_sig = {True: None, False: None}

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
    # N launches of one sample each would be all latency.  The arguments of consecutive calls are 1-D views whose data
    # pointers advance by exactly one element (np.loadtxt arrays: +4 bytes; scipy filtfilt output, a reversed-and-sliced
    # view: -8 bytes).  When a call's pointers sit one element after the previous call's, the columns that FOLLOW are read
    # speculatively through strided views (bounded by the memory of the arrays that own the data), the whole block of up to
    # COMPAT_CHUNK samples is computed in ONE sysid_projected_batch launch, and later calls are served from it -- after
    # checking that the bytes handed in are the bytes the block was computed from.  Everything else (plain vectors, lists,
    # random access) takes the single-sample launch.
    COMPAT_CHUNK = 4096

    @staticmethod
    def _owner_bounds(v):
        """[lo, hi) byte range of the array that owns v's memory; None when v owns its data (then it is no column view)."""
        b = v.base
        if not isinstance(b, np.ndarray):
            return None
        while isinstance(b.base, np.ndarray):
            b = b.base
        if b.ndim == 0 or b.size == 0:
            return None
        lo = b.__array_interface__["data"][0]
        ext = [(n - 1) * st for n, st in zip(b.shape, b.strides)]
        return lo + sum(e for e in ext if e < 0), lo + sum(e for e in ext if e > 0) + b.itemsize

    def _sig(self, vecs):
        """(pointer, row stride, itemsize, rows, dtype) per argument, or None when an argument cannot be a column view."""
        out = []
        for v in vecs:
            if v is None:
                out.append(None)
                continue
            if not isinstance(v, np.ndarray) or v.ndim != 1 or v.dtype not in (np.float32, np.float64) or v.shape[0] == 0:
                return None
            out.append((v.__array_interface__["data"][0], v.strides[0], v.itemsize, v.shape[0], v.dtype.str))
        return out

    def _open_block(self, vecs, sig, prev):
        """Called when `sig` is one element after `prev` in every argument: speculative block starting at this call's column."""
        from numpy.lib.stride_tricks import as_strided
        from .ops import to_device
        steps, n = [], self.COMPAT_CHUNK
        for v, a, b in zip(vecs, sig, prev):
            if a is None:
                steps.append(None)
                continue
            if b is None or a[1:] != b[1:] or abs(a[0] - b[0]) != a[2]:
                return None
            e = a[0] - b[0]
            own = self._owner_bounds(v)
            if own is None:
                return None
            ptr, S, item, rows = a[0], a[1], a[2], a[3]
            # columns j = 0 .. n-1 of rows 0 .. rows-1 live at ptr + r S + j e: keep them inside the owner and inside one row
            lo_r, hi_r = min(0, (rows - 1) * S), max(0, (rows - 1) * S)
            if e > 0:
                n = min(n, (own[1] - item - ptr - hi_r) // e + 1)
            else:
                n = min(n, (ptr + lo_r - own[0]) // (-e) + 1)
            if rows > 1:
                n = min(n, abs(S) // item)
            steps.append(e)
        if n < 2:
            return None
        host = []
        for v, a, e in zip(vecs, sig, steps):
            if a is None:
                host.append(np.zeros((self.joints_dof, n)))
            else:
                host.append(np.array(as_strided(v, shape=(a[3], n), strides=(a[1], e), writeable=False), dtype=np.float64, order="C"))
        A, b = self.device_model.projected_batch(*(to_device(h) for h in host), friction=True)
        return {"sig": sig, "steps": steps, "n": n, "host": host, "A": A.cpu().numpy(), "b": b.cpu().numpy()}

    def _block_lookup(self, vecs, sig):
        """(block, k) when every argument is column k of the current block and carries the bytes the block was computed from."""
        blk = self._block
        if blk is None:
            return None
        k = None
        for a, b0, e in zip(sig, blk["sig"], blk["steps"]):
            if a is None:
                continue                        # a call without torques is served by a block computed with them (friction blocks)
            if b0 is None or a[1:] != b0[1:] or (a[0] - b0[0]) % e != 0:
                return None
            kk = (a[0] - b0[0]) // e
            if k is None:
                k = kk
            if kk != k or not (0 <= kk < blk["n"]):
                return None
        for v, h in zip(vecs, blk["host"]):
            if v is not None and not np.array_equal(np.asarray(v, dtype=np.float64), h[:, k], equal_nan=True):
                self._block = None              # the log was modified in place after the block was computed
                return None
        return blk, int(k)

    def _block_for(self, vecs):
        sig = self._sig(vecs)
        kind = vecs[3] is None                  # the friction loop passes no torques: each kind of call has its own history
        if sig is None:
            self._prev_sig[kind] = None
            return None
        hit = self._block_lookup(vecs, sig)
        if hit is None and self._prev_sig[kind] is not None:
            blk = self._open_block(vecs, sig, self._prev_sig[kind])
            if blk is not None:
                self._block = blk
                hit = self._block_lookup(vecs, sig)
        self._prev_sig[kind] = sig
        return hit

    # ------------------------------------------------------------------ fixed-base emulation (floating_base=False)
    def _to_floating(self, q, dq, ddq, tau, cnt):
        """Fixed-base arrays ((nd, N) or (nd,)) -> the free-flyer kernels' arrays: base pinned at the identity with zero twist,
        the three anchors in stance.  float64 numpy or CUDA tensors (same kind out)."""
        import torch
        def widen(a, head):
            if isinstance(a, torch.Tensor):
                a = a.to(torch.float64)
                h = torch.tensor(head, dtype=torch.float64, device=a.device)
                h = h.reshape(-1, *([1] * (a.dim() - 1))).expand(len(head), *a.shape[1:])
                return torch.cat([h, a], dim=0).contiguous()
            a = np.asarray(a, dtype=np.float64)
            h = np.broadcast_to(np.asarray(head, dtype=np.float64).reshape(-1, *([1] * (a.ndim - 1))), (len(head),) + a.shape[1:])
            return np.ascontiguousarray(np.concatenate([h, a], axis=0))
        z6 = [0.0] * 6
        cnt = cnt if (isinstance(cnt, torch.Tensor) or np.asarray(cnt).size) else np.zeros((0,) + np.asarray(q).shape[1:])
        out = (widen(q, [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]), widen(dq, z6), widen(ddq, z6),
               tau if isinstance(tau, torch.Tensor) else (None if tau is None else np.asarray(tau, dtype=np.float64)), widen(cnt, [1.0, 1.0, 1.0]))
        return out

    def _fixed_cols(self, friction=True):
        """Columns of the free-flyer row block that belong to the fixed-base problem: the root body's ten are dropped."""
        nbF, d = self._flat.nbodies, self.joints_dof
        cols = list(range(10, 10 * nbF))
        if friction:
            cols += list(range(10 * nbF, 10 * nbF + 2 * d))
        return np.array(cols)

    def _stats_to_fixed(self, stats_F, friction=True):
        """[G | r | s | n] of the free-flyer emulation -> the fixed-base problem's (root-body columns dropped; n = N nd: the
        reference counts the rows of ITS stack, quirk Q4)."""
        import torch
        nbF, d = self._flat.nbodies, self.joints_dof
        cF = 10 * nbF + (2 * d if friction else 0)
        idx = torch.as_tensor(self._fixed_cols(friction), device=stats_F.device)
        G = stats_F[:cF * cF].view(cF, cF).index_select(0, idx).index_select(1, idx)
        r = stats_F[cF * cF:cF * cF + cF].index_select(0, idx)
        tail = stats_F[cF * cF + cF:cF * cF + cF + 2] * torch.tensor([1.0, d / (d + 6.0)], dtype=torch.float64, device=stats_F.device)
        return torch.cat([G.reshape(-1), r, tail]).contiguous()

    def _one_sample(self, q, dq, ddq, tau, cnt):
        if self._fixed and not getattr(self, "_in_fixed", False):
            qF, dqF, ddqF, tauF, cntF = self._to_floating(np.asarray(q, dtype=np.float64), np.asarray(dq, dtype=np.float64),
                                                          np.asarray(ddq, dtype=np.float64), tau, np.asarray(cnt, dtype=np.float64))
            self._in_fixed = True
            try:
                A, b = self._one_sample(qF, dqF, ddqF, tauF, cntF)
            finally:
                self._in_fixed = False
            return A[6:][:, self._fixed_cols(True)], b[6:]
        return self._one_sample_floating(q, dq, ddq, tau, cnt)

    def _one_sample_floating(self, q, dq, ddq, tau, cnt):
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
        fl = self._flat
        for n in (fl.nq, fl.nv, fl.nv, self.joints_dof, fl.n_ee):
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
        if self._fixed:
            q, dq, ddq, tau, cnt = self._to_floating(q, dq, ddq, tau, cnt)
        dev = self._upload(q, dq, ddq, tau, cnt)
        st = self.device_model.gram_accumulate(*dev, friction=friction, weights=weights)
        return self._stats_to_fixed(st, friction) if self._fixed else st

    def tau_prediction_rmse(self, q, dq, ddq, torque, cnt, phi):
        """(total, per-joint) with the reference's formulas: total = mean_i ||e_i||^2 (no root), per joint = RMSE."""
        import torch
        phi = np.asarray(phi, dtype=np.float64)
        if self._fixed:
            q, dq, ddq, torque, cnt = self._to_floating(q, dq, ddq, torque, cnt)
            phi = np.concatenate([np.zeros(10), phi])                   # the pinned root body does not move
        dev = self._upload(q, dq, ddq, torque, cnt)
        out = self.device_model.predict_rmse(*dev, torch.as_tensor(phi)).cpu().numpy()
        if self._fixed:
            # the reference slices (y @ phi)[6:] whatever the base (src/sys_identification.py:429-430): on a fixed-base model that
            # drops the first six JOINTS; total = mean_i sum_k e_ik^2 = sum_k rmse_k^2 over the joints that are kept
            pj = out[1:][6:].copy()
            return float(np.sum(pj ** 2)), pj
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
