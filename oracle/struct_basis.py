"""TEST INFRASTRUCTURE (oracle/): numpy restatement of the STRUCTURED null-space basis the CUDA kernel gram_struct_kernel builds
(system_identification_b200/csrc/gram_struct.cuh), checked against the reference's own projector
P = I - pinv(J_c) J_c (reference src/sys_identification.py:127-135).

Per sample: swing legs give unit vectors, a stance leg with more than three joints gives the null space of its own 3 x len
Jacobian block (Householder QR of J_leg^T), and the remaining (dense) directions come from a Householder QR -- with the pinv
rank rule -- of the reduced Jacobian [J_base | T_f], T_f = J_leg Q_range, expanded through Q_range.  Q Q^T must equal P.

Only tests/ import this module; the product never does."""
import numpy as np

from oracle import dynamics as D

def house_null(Jt, tol_scale=None):
    """Householder QR of Jt (n x r) column by column with the rank rule; returns (reflectors list, rank)."""
    n, r = Jt.shape
    X = Jt.copy()
    mx = max((X[:, b] @ X[:, b] for b in range(r)), default=0.0)
    tol = 1e-13 * mx
    refl = []
    p = 0
    for b in range(r):
        x = X[:, b]
        tail2 = x[p:] @ x[p:]
        if tail2 <= tol: continue
        nt = np.sqrt(tail2); alpha = nt if x[p] >= 0 else -nt
        v = np.zeros(n); v[p:] = x[p:]; v[p] += alpha; v /= np.linalg.norm(v)
        refl.append(v)
        for b2 in range(b + 1, r):
            X[:, b2] -= 2 * v * (v @ X[:, b2])
        p += 1
    return refl, p

def struct_basis(flat, tree, q, cnt):
    nv = flat.nv
    Jc = D.contact_jacobian(tree, q, cnt, flat.ee_names)    # (3m x nv)
    m = Jc.shape[0] // 3
    stance = [k for k in range(flat.n_ee) if cnt[k] != 0]
    # chains: children of the root
    par = flat.parent
    chains = []
    for j in range(2, flat.njoints):
        if par[j] == 1:
            ch = [j]
            while True:
                kids = [k for k in range(flat.njoints) if par[k] == ch[-1]]
                if not kids: break
                assert len(kids) == 1; ch.append(kids[0])
            chains.append(ch)
    foot_chain = {}
    for k in range(flat.n_ee):
        for ci, ch in enumerate(chains):
            if flat.ee_joint[k] in ch: foot_chain[k] = (ci, ch.index(flat.ee_joint[k]) + 1)
    sparse = {ci: [] for ci in range(len(chains))}
    red_cols = []     # list of (slot, Qrange (nv x nr))
    nred = 6
    Jred_blocks = []
    for slot, k in enumerate(stance):
        rows = Jc[3 * slot:3 * slot + 3]
        if k not in foot_chain:
            Jred_blocks.append((rows[:, :6], None, 0)); continue
        ci, ln = foot_chain[k]
        cols = [4 + j for j in chains[ci][:ln]]        # idx_v = joint + 4
        Jl = rows[:, cols]                              # 3 x ln
        # QR of Jl^T (ln x 3), no rank rule: reflector skipped only on an exactly zero tail
        X = Jl.T.copy(); refl = []
        nr = min(3, ln)
        for p in range(nr):
            x = X[:, p]; tail2 = x[p:] @ x[p:]
            v = np.zeros(ln)
            if tail2 > 0:
                nt = np.sqrt(tail2); alpha = nt if x[p] >= 0 else -nt
                v[p:] = x[p:]; v[p] += alpha; v /= np.linalg.norm(v)
                for b2 in range(p, 3): X[:, b2] -= 2 * v * (v @ X[:, b2])
            refl.append(v)
        Ql = np.eye(ln)
        for v in reversed(refl): Ql -= 2 * np.outer(v, v @ Ql)
        T = Jl @ Ql[:, :nr]                              # 3 x nr  (lower triangular up to rounding)
        assert ln == nr or np.abs(Jl @ Ql[:, nr:]).max() < 1e-12
        Qfull = np.zeros((nv, ln)); Qfull[cols, :] = Ql
        Jred_blocks.append((rows[:, :6], T, nr))
        red_cols.append((slot, Qfull[:, :nr]))
        for i in range(nr, ln): sparse[ci].append(Qfull[:, i])
        for j in chains[ci][ln:]:
            e = np.zeros(nv); e[4 + j] = 1; sparse[ci].append(e)
    for ci, ch in enumerate(chains):
        if not any(foot_chain.get(k, (None,))[0] == ci for k in stance):
            for j in ch:
                e = np.zeros(nv); e[4 + j] = 1; sparse[ci].append(e)
    nred = 6 + sum(b[2] for b in Jred_blocks)
    Jred = np.zeros((3 * m, nred)); off = 6; offs = []
    for slot, (B, T, nr) in enumerate(Jred_blocks):
        Jred[3 * slot:3 * slot + 3, :6] = B
        offs.append(off)
        if nr: Jred[3 * slot:3 * slot + 3, off:off + nr] = T
        off += nr
    refl, rank = house_null(Jred.T)
    dense = []
    for k in range(rank, nred):
        x = np.zeros(nred); x[k] = 1
        for v in reversed(refl): x -= 2 * v * (v @ x)
        qv = np.zeros(nv); qv[:6] = x[:6]
        for (slot, Qr) in red_cols:
            qv += Qr @ x[offs[slot]:offs[slot] + Qr.shape[1]]
        dense.append(qv)
    return Jc, dense, sparse, chains

