// Stage 1 of the fused kernels, decomposed into short data-parallel phases over a super-batch of SB samples.
//
// The per-sample math is restated from pinocchio (upstream of reference src/sys_identification.py:113-135,395-418;
// SURVEY.md App. A) in a form whose cost is ~1/4 of walking every regressor column up the chain:
//
//   * Every row of the unprojected regressor block of body i is a 6-vector of PLUECKER coordinates (base frame, taken
//     at the base origin) dotted with the column's wrench (F; N_O):  base rows = I_6, row of ancestor joint j =
//     a_j = (m_j; z_j), z_j the joint axis, m_j = p_j x z_j its moment.  Hence for row r of the PROJECTED block
//         (P Y_i)[r] = (dl; da) . (F; N_O),   dl = P[r][0:3] + sum_{j in anc(i)} P[r][6+j] m_j,
//                                               da = P[r][3:6] + sum_{j in anc(i)} P[r][6+j] z_j,
//     i.e. the null-space projector is applied to the 6+depth Pluecker rows, not to 130 columns.
//   * Brought to the body frame, el = R_i^T (dl + da x p_i), ea = R_i^T da, the ten entries of the row are closed
//     forms in (omega, alpha, acc) of the body (bodyRegressor^T applied to (el; ea)): ~110 flops per (body, row).
//
// Phases (a group of NT threads, barriers between them; `t` is the thread's index in the group):
//   chains    lane per (sample, leaf chain): sin/cos, poses relative to the base, spatial velocity / gravity-biased
//             acceleration down the chain -> X_j = (R_j, p_j), a_j, b9_j = (omega, alpha, acc)
//   feet      thread per (sample, stance slot): world-aligned lever arm and leg columns of the contact Jacobian
//   sblocks   thread per (sample, pair of stance feet): 3x3 block of S = J_c J_c^T
//   chol      lane per sample: Cholesky S = L L^T, dependent rows dropped (pinv semantics)
//   wcols     thread per (sample, dof): column of W = L^-1 J_c by forward substitution
//   proj      thread per (sample, entry): packed lower triangle of P = I - W^T W
//   fill      thread per (sample, body, row) + per (sample, row) for the friction / torque columns -> tile rows
//
// The contact Jacobian keeps pinocchio's exact semantics for an UN-normalised logged quaternion (float32 logs are
// off unit norm by ~3e-8; assuming an orthonormal R_b moves P by ~2e-10, above the parity gate):
//   J_k = [ R_b | -[R_b r_k]x R_b | (R_b z_c) x (R_b d_kc) ].
#pragma once
#include "kinematics.cuh"

namespace sysid {

// ---- per-sample context (doubles) ------------------------------------------------------------------------
constexpr int NPACK = MAXV * (MAXV + 1) / 2;        // 171
constexpr int CX_P = 0;                             // [171] packed lower triangle of P
constexpr int CX_W = CX_P + NPACK;                  // sqrt(weight); 0 => sample contributes nothing
constexpr int CX_A = CX_W + 1;                      // [MAXD][6] Pluecker axis (m; z) of every revolute joint
constexpr int CX_X = CX_A + 6 * MAXD;               // [MAXD][12] R (9, row-major), p (3) relative to the base
constexpr int CX_B9 = CX_X + 12 * MAXD;             // [MAXB][9] omega, alpha, acc (local frame)
constexpr int CX_DQ = CX_B9 + 9 * MAXB;             // [MAXD]
constexpr int CX_TAU = CX_DQ + MAXD;                // [MAXD]
constexpr int CX_STRIDE = CX_TAU + MAXD + 1;        // 530
// temporaries living in the P slot until `proj` overwrites it
constexpr int CXT_S = CX_P;                         // [78] packed S, then L
constexpr int CXT_JL = CX_P + 78;                   // [MAXEE][MAXCH][3] leg columns of J_c
static_assert(78 + 3 * MAXEE * MAXCH <= NPACK, "temporaries fit the P slot");
static_assert(CX_STRIDE % 2 == 0, "context stride keeps 16-byte alignment");

// ---- per-sample scratch (doubles), only live inside the F phases -------------------------------------------
constexpr int SC_RB = 0;                            // [9] R_b from the raw quaternion
constexpr int SC_RF = SC_RB + 9;                    // [MAXEE][3] R_b r_k of the stance feet
constexpr int SC_META = SC_RF + 3 * MAXEE;          // 3m, then the foot index of every stance slot
constexpr int SC_WM = SC_META + 1 + MAXEE;          // [3*MAXEE][MAXV] W = L^-1 J_c
constexpr int SC_STRIDE = SC_WM + 3 * MAXEE * MAXV + 1;   // 243 (odd)

__device__ __forceinline__ int pk(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }

// ---------------------------------------------------------------------------------------------- chains
template <int SB>
__device__ __forceinline__ void phase_chains(const DevModel& M, const SampleIO& io, long long base, long long N,
                                             double* __restrict__ ctx, double* __restrict__ scr, int* s_bad, int t) {
    if (t >= SB * M.nfch) return;
    const int s = t % SB, ch = t / SB;          // sample fastest: coalesced channel loads
    const long long i = base + s;
    if (i >= N) return;
    double* c = ctx + s * CX_STRIDE;
    const long long ld = io.ld;
    double probe = 0.0;
    double Rb[9];
    {   // Eigen::Quaternion::toRotationMatrix on the raw (x, y, z, w): no normalisation, as pinocchio's free-flyer does
        const double qx = io.q[3 * ld + i], qy = io.q[4 * ld + i], qz = io.q[5 * ld + i], qw = io.q[6 * ld + i];
        probe += qx + qy + qz + qw;
        const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
        const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        Rb[0] = 1 - (tyy + tzz); Rb[1] = txy - twz; Rb[2] = txz + twy;
        Rb[3] = txy + twz; Rb[4] = 1 - (txx + tzz); Rb[5] = tyz - twx;
        Rb[6] = txz - twy; Rb[7] = tyz + twx; Rb[8] = 1 - (txx + tyy);
    }
    double v[6], a[6];
    {   // root (free-flyer): v = dq[0:6]; a = ddq[0:6] + [R_b^T (-g); 0]
        const double g0 = -M.gravity[0], g1 = -M.gravity[1], g2 = -M.gravity[2];
#pragma unroll
        for (int k = 0; k < 6; ++k) { v[k] = io.dq[k * ld + i]; a[k] = io.ddq[k * ld + i]; probe += v[k] + a[k]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] += Rb[k] * g0 + Rb[3 + k] * g1 + Rb[6 + k] * g2;
    }
    if (ch == 0) {
        double* sc = scr + s * SC_STRIDE;
#pragma unroll
        for (int k = 0; k < 9; ++k) sc[SC_RB + k] = Rb[k];
        double* b9 = c + CX_B9;
        b9[0] = v[3]; b9[1] = v[4]; b9[2] = v[5];
        b9[3] = a[3]; b9[4] = a[4]; b9[5] = a[5];
        b9[6] = a[0] + (v[4] * v[2] - v[5] * v[1]);
        b9[7] = a[1] + (v[5] * v[0] - v[3] * v[2]);
        b9[8] = a[2] + (v[3] * v[1] - v[4] * v[0]);
    }
    double R[9], p[3];
    const int len = M.fch_len[ch], own = M.fch_own[ch];
    for (int e = 0; e < len; ++e) {
        const int j = M.fch[ch][e];
        const double th = io.q[(5 + j) * ld + i], qd = io.dq[(4 + j) * ld + i], qdd = io.ddq[(4 + j) * ld + i];
        probe += th + qd + qdd;
        double sn, cs;
        sincos(th, &sn, &cs);
        double Rl[9];
        joint_rotation_compose(M, j, sn, cs, Rl);
        const double px = M.pp[j][0], py = M.pp[j][1], pz = M.pp[j][2];
        {   // motion: actInv of the parent's (v, a), then the joint's own contribution
            const double tvx = v[0] - (py * v[5] - pz * v[4]), tvy = v[1] - (pz * v[3] - px * v[5]), tvz = v[2] - (px * v[4] - py * v[3]);
            const double tax = a[0] - (py * a[5] - pz * a[4]), tay = a[1] - (pz * a[3] - px * a[5]), taz = a[2] - (px * a[4] - py * a[3]);
            double nv_[6], na[6];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                nv_[k] = Rl[k] * tvx + Rl[3 + k] * tvy + Rl[6 + k] * tvz;
                nv_[3 + k] = Rl[k] * v[3] + Rl[3 + k] * v[4] + Rl[6 + k] * v[5];
                na[k] = Rl[k] * tax + Rl[3 + k] * tay + Rl[6 + k] * taz;
                na[3 + k] = Rl[k] * a[3] + Rl[3 + k] * a[4] + Rl[6 + k] * a[5];
            }
            double wj[3];
            const int jt = M.jtype[j];
#pragma unroll
            for (int k = 0; k < 3; ++k) wj[k] = (jt == JT_RU) ? M.axis[j][k] : ((jt - JT_RX) == k ? 1.0 : 0.0);
#pragma unroll
            for (int k = 0; k < 3; ++k) nv_[3 + k] += wj[k] * qd;
            const double w0 = wj[0] * qd, w1 = wj[1] * qd, w2 = wj[2] * qd;
            na[0] += nv_[1] * w2 - nv_[2] * w1; na[1] += nv_[2] * w0 - nv_[0] * w2; na[2] += nv_[0] * w1 - nv_[1] * w0;
            na[3] += nv_[4] * w2 - nv_[5] * w1 + wj[0] * qdd; na[4] += nv_[5] * w0 - nv_[3] * w2 + wj[1] * qdd; na[5] += nv_[3] * w1 - nv_[4] * w0 + wj[2] * qdd;
#pragma unroll
            for (int k = 0; k < 6; ++k) { v[k] = nv_[k]; a[k] = na[k]; }
        }
        if (e == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = Rl[k];
            p[0] = px; p[1] = py; p[2] = pz;
        } else {
            double Rn[9];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int k = 0; k < 3; ++k) Rn[3 * r + k] = R[3 * r] * Rl[k] + R[3 * r + 1] * Rl[3 + k] + R[3 * r + 2] * Rl[6 + k];
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) p[r] += R[3 * r] * px + R[3 * r + 1] * py + R[3 * r + 2] * pz;
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = Rn[k];
        }
        if (e >= own) {
            double* X = c + CX_X + 12 * (j - 2);
#pragma unroll
            for (int k = 0; k < 9; ++k) X[k] = R[k];
            X[9] = p[0]; X[10] = p[1]; X[11] = p[2];
            double z0, z1, z2;
            const int jt = M.jtype[j];
            if (jt == JT_RU) {
                const double u0 = M.axis[j][0], u1 = M.axis[j][1], u2 = M.axis[j][2];
                z0 = R[0] * u0 + R[1] * u1 + R[2] * u2; z1 = R[3] * u0 + R[4] * u1 + R[5] * u2; z2 = R[6] * u0 + R[7] * u1 + R[8] * u2;
            } else if (jt == JT_RX) { z0 = R[0]; z1 = R[3]; z2 = R[6]; }
            else if (jt == JT_RY) { z0 = R[1]; z1 = R[4]; z2 = R[7]; }
            else { z0 = R[2]; z1 = R[5]; z2 = R[8]; }
            double* A = c + CX_A + 6 * (j - 2);
            A[0] = p[1] * z2 - p[2] * z1; A[1] = p[2] * z0 - p[0] * z2; A[2] = p[0] * z1 - p[1] * z0;
            A[3] = z0; A[4] = z1; A[5] = z2;
            double* b9 = c + CX_B9 + 9 * (j - 1);
            b9[0] = v[3]; b9[1] = v[4]; b9[2] = v[5];
            b9[3] = a[3]; b9[4] = a[4]; b9[5] = a[5];
            b9[6] = a[0] + (v[4] * v[2] - v[5] * v[1]);
            b9[7] = a[1] + (v[5] * v[0] - v[3] * v[2]);
            b9[8] = a[2] + (v[3] * v[1] - v[4] * v[0]);
        }
    }
    if (!(fabs(probe) < 1e300)) atomicOr(&s_bad[s], 2);
}

// ---------------------------------------------------------------------------------------------- feet
// Also copies dq / tau of the actuated joints into the context (threads of slot 0).
template <int SB>
__device__ __forceinline__ void phase_feet(const DevModel& M, const SampleIO& io, long long base, long long N,
                                           double* __restrict__ ctx, double* __restrict__ scr, int* s_bad, int t) {
    if (t >= SB * MAXEE) return;
    const int s = t % SB, slot = t / SB;
    const long long i = base + s;
    if (i >= N) return;
    double* c = ctx + s * CX_STRIDE;
    double* sc = scr + s * SC_STRIDE;
    const long long ld = io.ld;
    int m = 0, kf = -1;
#pragma unroll
    for (int k = 0; k < MAXEE; ++k) {
        if (k < M.n_ee) {
            const double cv = io.cnt[k * ld + i];
            if (cv != 0.0) {   // truthiness rule of the reference: state 2 counts as stance; NaN is truthy too
                if (m == slot) kf = k;
                ++m;
            }
        }
    }
    if (slot == 0) {
        sc[SC_META] = (double)(3 * m);
        double probe = 0.0;
        for (int k = 0; k < M.nd; ++k) {
            const double tk = io.tau ? io.tau[k * ld + i] : 0.0;
            probe += tk;
            c[CX_DQ + k] = io.dq[(6 + k) * ld + i];
            c[CX_TAU + k] = tk;
        }
        if (!(fabs(probe) < 1e300)) atomicOr(&s_bad[s], 2);
    }
    sc[SC_META + 1 + slot] = (double)kf;
    if (kf < 0) return;
    double Rb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rb[k] = sc[SC_RB + k];
    const int jf = M.ee_joint[kf];
    double rf[3];   // foot point in the base frame
    if (jf == 1) {
#pragma unroll
        for (int e = 0; e < 3; ++e) rf[e] = M.ee_off[kf][e];
    } else {
        const double* X = c + CX_X + 12 * (jf - 2);
#pragma unroll
        for (int e = 0; e < 3; ++e) rf[e] = X[9 + e] + X[3 * e] * M.ee_off[kf][0] + X[3 * e + 1] * M.ee_off[kf][1] + X[3 * e + 2] * M.ee_off[kf][2];
    }
#pragma unroll
    for (int e = 0; e < 3; ++e) sc[SC_RF + 3 * slot + e] = Rb[3 * e] * rf[0] + Rb[3 * e + 1] * rf[1] + Rb[3 * e + 2] * rf[2];
    const int len = M.chain_len[kf];
    for (int e = 0; e < len; ++e) {
        const int cj = M.chain[kf][e];
        const double* A = c + CX_A + 6 * (cj - 2);
        const double* X = c + CX_X + 12 * (cj - 2);
        const double ax0 = A[3], ax1 = A[4], ax2 = A[5];
        const double bx = rf[0] - X[9], by = rf[1] - X[10], bz = rf[2] - X[11];
        const double a0 = Rb[0] * ax0 + Rb[1] * ax1 + Rb[2] * ax2, a1 = Rb[3] * ax0 + Rb[4] * ax1 + Rb[5] * ax2, a2 = Rb[6] * ax0 + Rb[7] * ax1 + Rb[8] * ax2;
        const double dx = Rb[0] * bx + Rb[1] * by + Rb[2] * bz, dy = Rb[3] * bx + Rb[4] * by + Rb[5] * bz, dz = Rb[6] * bx + Rb[7] * by + Rb[8] * bz;
        double* jl = c + CXT_JL + 3 * (slot * MAXCH + e);
        jl[0] = a1 * dz - a2 * dy;
        jl[1] = a2 * dx - a0 * dz;
        jl[2] = a0 * dy - a1 * dx;
    }
}

// ---------------------------------------------------------------------------------------------- S blocks
template <int SB>
__device__ __forceinline__ void phase_sblocks(const DevModel& M, long long base, long long N,
                                              double* __restrict__ ctx, const double* __restrict__ scr, int t) {
    constexpr int NPAIR = MAXEE * (MAXEE + 1) / 2;
    if (t >= SB * NPAIR) return;
    const int s = t % SB, pr = t / SB;
    if (base + s >= N) return;
    int st = 0;
    while ((st + 1) * (st + 2) / 2 <= pr) ++st;
    const int su = pr - st * (st + 1) / 2;           // su <= st
    double* c = ctx + s * CX_STRIDE;
    const double* sc = scr + s * SC_STRIDE;
    const int m = (int)sc[SC_META] / 3;
    if (st >= m) return;
    const int kt = (int)sc[SC_META + 1 + st], ku = (int)sc[SC_META + 1 + su];
    double Rb[9], BB[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rb[k] = sc[SC_RB + k];
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) BB[3 * x + y] = Rb[3 * x] * Rb[3 * y] + Rb[3 * x + 1] * Rb[3 * y + 1] + Rb[3 * x + 2] * Rb[3 * y + 2];
    const double rt0 = sc[SC_RF + 3 * st], rt1 = sc[SC_RF + 3 * st + 1], rt2 = sc[SC_RF + 3 * st + 2];
    const double ru0 = sc[SC_RF + 3 * su], ru1 = sc[SC_RF + 3 * su + 1], ru2 = sc[SC_RF + 3 * su + 2];
    // block(t,u) = J_t J_u^T.  Base part: BB + [r_t]x BB [r_u]x^T
    double XB[9], B[9];
#pragma unroll
    for (int y = 0; y < 3; ++y) {          // XB = [r_t]x BB
        XB[y] = -rt2 * BB[3 + y] + rt1 * BB[6 + y];
        XB[3 + y] = rt2 * BB[y] - rt0 * BB[6 + y];
        XB[6 + y] = -rt1 * BB[y] + rt0 * BB[3 + y];
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) {          // B = BB - XB [r_u]x   ([r]x^T = -[r]x)
        const double m0 = XB[3 * x], m1 = XB[3 * x + 1], m2 = XB[3 * x + 2];
        B[3 * x] = BB[3 * x] - (m1 * ru2 - m2 * ru1);
        B[3 * x + 1] = BB[3 * x + 1] - (m2 * ru0 - m0 * ru2);
        B[3 * x + 2] = BB[3 * x + 2] - (m0 * ru1 - m1 * ru0);
    }
    // leg part: joints common to both foot chains (aligned at the root end of the chains)
    const int ns = M.nshared[kt][ku];
    const int lt = M.chain_len[kt], lu = M.chain_len[ku];
    for (int e = 0; e < ns; ++e) {
        const double* ja = c + CXT_JL + 3 * (st * MAXCH + (lt - ns + e));
        const double* jb = c + CXT_JL + 3 * (su * MAXCH + (lu - ns + e));
        const double a0 = ja[0], a1 = ja[1], a2 = ja[2], b0 = jb[0], b1 = jb[1], b2 = jb[2];
        B[0] += a0 * b0; B[1] += a0 * b1; B[2] += a0 * b2;
        B[3] += a1 * b0; B[4] += a1 * b1; B[5] += a1 * b2;
        B[6] += a2 * b0; B[7] += a2 * b1; B[8] += a2 * b2;
    }
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) {
            const int gi = 3 * st + x, gj = 3 * su + y;
            if (gi >= gj) c[CXT_S + tri(gi, gj)] = B[3 * x + y];
        }
}

// ---------------------------------------------------------------------------------------------- Cholesky
template <int SB>
__device__ __forceinline__ void phase_chol(long long base, long long N, double* __restrict__ ctx,
                                           const double* __restrict__ scr, int* s_bad, int t) {
    if (t >= SB || base + t >= N) return;
    double* S = ctx + t * CX_STRIDE + CXT_S;
    const int m3 = (int)scr[t * SC_STRIDE + SC_META];
    double maxdiag = 0.0;
    for (int a = 0; a < m3; ++a) maxdiag = fmax(maxdiag, S[tri(a, a)]);
    const double piv_tol = 1e-13 * maxdiag;
    int flags = 0;
    for (int a = 0; a < m3; ++a) {
        double d = S[tri(a, a)];
        for (int k = 0; k < a; ++k) { const double l = S[tri(a, k)]; d -= l * l; }
        double inv;
        if (d > piv_tol) { const double sd = sqrt(d); S[tri(a, a)] = sd; inv = 1.0 / sd; }
        else { S[tri(a, a)] = 0.0; inv = 0.0; flags |= 1; }    // row a is (numerically) dependent: drop it
        for (int b = a + 1; b < m3; ++b) {
            double vv = S[tri(b, a)];
            for (int k = 0; k < a; ++k) vv -= S[tri(b, k)] * S[tri(a, k)];
            S[tri(b, a)] = vv * inv;
        }
    }
    if (flags) atomicOr(&s_bad[t], flags);
}

// ---------------------------------------------------------------------------------------------- W columns
template <int SB>
__device__ __forceinline__ void phase_wcols(const DevModel& M, long long base, long long N,
                                            const double* __restrict__ ctx, double* __restrict__ scr, int t) {
    if (t >= SB * MAXV) return;
    const int s = t % SB, col = t / SB;
    if (base + s >= N) return;
    const double* c = ctx + s * CX_STRIDE;
    double* sc = scr + s * SC_STRIDE;
    const int m3 = (int)sc[SC_META];
    const double* L = c + CXT_S;
    double w[3 * MAXEE];
#pragma unroll
    for (int slot = 0; slot < MAXEE; ++slot) {
        if (3 * slot < m3) {
            const int kt = (int)sc[SC_META + 1 + slot];
            const double r0 = sc[SC_RF + 3 * slot], r1 = sc[SC_RF + 3 * slot + 1], r2 = sc[SC_RF + 3 * slot + 2];
            double j0 = 0.0, j1 = 0.0, j2 = 0.0;      // J_c[3 slot + x][col]
            if (col < 3) { j0 = sc[SC_RB + col]; j1 = sc[SC_RB + 3 + col]; j2 = sc[SC_RB + 6 + col]; }
            else if (col < 6) {
                const double b0 = sc[SC_RB + col - 3], b1 = sc[SC_RB + col], b2 = sc[SC_RB + col + 3];
                j0 = -(r1 * b2 - r2 * b1); j1 = -(r2 * b0 - r0 * b2); j2 = -(r0 * b1 - r1 * b0);
            } else {
                const int jn = col - 4;               // joint whose idx_v is col
                const int len = M.chain_len[kt];
                for (int e = 0; e < len; ++e)
                    if (M.chain[kt][e] == jn) {
                        const double* jl = c + CXT_JL + 3 * (slot * MAXCH + e);
                        j0 = jl[0]; j1 = jl[1]; j2 = jl[2];
                    }
            }
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const int k = 3 * slot + x;
                double val = (x == 0) ? j0 : ((x == 1) ? j1 : j2);
#pragma unroll
                for (int l = 0; l < 3 * MAXEE; ++l) if (l < k) val = fma(-L[tri(k, l)], w[l], val);
                const double lkk = L[tri(k, k)];
                const double inv = (lkk != 0.0) ? 1.0 / lkk : 0.0;     // dropped (dependent) row: W row = 0
                w[k] = val * inv;
                sc[SC_WM + k * MAXV + col] = w[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- P = I - W^T W
// NT = threads in the group; also finalises the per-sample weight and the skip flags (first SB threads).
template <int SB, int NT>
__device__ __forceinline__ void phase_proj(const SampleIO& io, long long base, long long N, double* __restrict__ ctx,
                                           const double* __restrict__ scr, const int* s_bad, int t,
                                           double& wsum, int& nflag0, int& nflag1) {
    for (int it = t; it < SB * NPACK; it += NT) {
        const int s = it % SB, e = it / SB;
        if (base + s >= N) continue;
        int r = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
        while (r * (r + 1) / 2 > e) --r;
        while ((r + 1) * (r + 2) / 2 <= e) ++r;
        const int cc = e - r * (r + 1) / 2;
        const double* sc = scr + s * SC_STRIDE;
        const int m3 = (int)sc[SC_META];
        double pv = (r == cc) ? 1.0 : 0.0;
        for (int k = 0; k < m3; ++k) pv = fma(-sc[SC_WM + k * MAXV + r], sc[SC_WM + k * MAXV + cc], pv);
        ctx[s * CX_STRIDE + CX_P + e] = pv;
    }
    if (t < SB) {
        const long long i = base + t;
        double w = 0.0;
        if (i < N) {
            const int bad = s_bad[t];
            w = io.weights ? io.weights[i] : 1.0;
            if (!(fabs(w) < 1e300) || (bad & 2)) { w = 0.0; ++nflag1; }
            if (bad & 1) ++nflag0;
            w = fmax(w, 0.0);
        }
        ctx[t * CX_STRIDE + CX_W] = sqrt(w);
        wsum += w;
    }
}

// ---------------------------------------------------------------------------------------------- tile fill
// One item = one (sample, body, row) -> ten entries of the projected row block, or one (sample, row) -> the friction
// and torque columns plus the zero padding.  TS samples starting at local sample s0 -> rows [0, TS*MAXV) of `tile`.
template <int TS, int LD, int NT>
__device__ __forceinline__ void phase_fill(const DevModel& M, const double* __restrict__ ctx, double* __restrict__ tile,
                                           int s0, int friction, int t) {
    const int nb = M.nb, np = M.nparams, nd = M.nd;
    const int per_sample = MAXV * (nb + 1);
    for (int it = t; it < TS * per_sample; it += NT) {
        const int sl = it / per_sample, rem = it - sl * per_sample;
        const int ib = rem / MAXV, r = rem - ib * MAXV;      // ib in [0, nb]: body ib+1, or nb = friction/torque item
        const double* c = ctx + (s0 + sl) * CX_STRIDE;
        const double* P = c + CX_P;
        const double wsq = c[CX_W];
        double* row = tile + (sl * MAXV + r) * LD;
        if (ib < nb) {
            double* dst = row + 10 * ib;
            if (wsq == 0.0) {
#pragma unroll
                for (int k = 0; k < 10; ++k) dst[k] = 0.0;
                continue;
            }
            const int i = ib + 1;
            double dl0 = P[pk(r, 0)], dl1 = P[pk(r, 1)], dl2 = P[pk(r, 2)], da0 = P[pk(r, 3)], da1 = P[pk(r, 4)], da2 = P[pk(r, 5)];
            for (int j = i; j > 1; j = M.parent[j]) {
                const double pj = P[pk(r, 4 + j)];
                const double* A = c + CX_A + 6 * (j - 2);
                dl0 = fma(pj, A[0], dl0); dl1 = fma(pj, A[1], dl1); dl2 = fma(pj, A[2], dl2);
                da0 = fma(pj, A[3], da0); da1 = fma(pj, A[4], da1); da2 = fma(pj, A[5], da2);
            }
            double el0, el1, el2, ea0, ea1, ea2;
            if (i > 1) {
                const double* X = c + CX_X + 12 * (i - 2);
                const double p0 = X[9], p1 = X[10], p2 = X[11];
                const double t0 = dl0 + (da1 * p2 - da2 * p1), t1 = dl1 + (da2 * p0 - da0 * p2), t2 = dl2 + (da0 * p1 - da1 * p0);
                el0 = X[0] * t0 + X[3] * t1 + X[6] * t2; el1 = X[1] * t0 + X[4] * t1 + X[7] * t2; el2 = X[2] * t0 + X[5] * t1 + X[8] * t2;
                ea0 = X[0] * da0 + X[3] * da1 + X[6] * da2; ea1 = X[1] * da0 + X[4] * da1 + X[7] * da2; ea2 = X[2] * da0 + X[5] * da1 + X[8] * da2;
            } else {
                el0 = dl0; el1 = dl1; el2 = dl2; ea0 = da0; ea1 = da1; ea2 = da2;
            }
            el0 *= wsq; el1 *= wsq; el2 *= wsq; ea0 *= wsq; ea1 *= wsq; ea2 *= wsq;
            const double* b9 = c + CX_B9 + 9 * ib;
            const double w0 = b9[0], w1 = b9[1], w2 = b9[2], al0 = b9[3], al1 = b9[4], al2 = b9[5], ac0 = b9[6], ac1 = b9[7], ac2 = b9[8];
            // mass column: el . acc
            dst[0] = el0 * ac0 + el1 * ac1 + el2 * ac2;
            // first-moment columns: -alpha x el + omega x (omega x el) + acc x ea
            const double u0 = w1 * el2 - w2 * el1, u1 = w2 * el0 - w0 * el2, u2 = w0 * el1 - w1 * el0;
            dst[1] = (w1 * u2 - w2 * u1) - (al1 * el2 - al2 * el1) + (ac1 * ea2 - ac2 * ea1);
            dst[2] = (w2 * u0 - w0 * u2) - (al2 * el0 - al0 * el2) + (ac2 * ea0 - ac0 * ea2);
            dst[3] = (w0 * u1 - w1 * u0) - (al0 * el1 - al1 * el0) + (ac0 * ea1 - ac1 * ea0);
            // inertia columns (Ixx, Ixy, Iyy, Ixz, Iyz, Izz): ea . Br(alpha)[:,k] + (ea x omega) . Br(omega)[:,k]
            const double g0 = ea1 * w2 - ea2 * w1, g1 = ea2 * w0 - ea0 * w2, g2 = ea0 * w1 - ea1 * w0;
            dst[4] = ea0 * al0 + g0 * w0;
            dst[5] = ea0 * al1 + ea1 * al0 + g0 * w1 + g1 * w0;
            dst[6] = ea1 * al1 + g1 * w1;
            dst[7] = ea0 * al2 + ea2 * al0 + g0 * w2 + g2 * w0;
            dst[8] = ea1 * al2 + ea2 * al1 + g1 * w2 + g2 * w1;
            dst[9] = ea2 * al2 + g2 * w2;
        } else {
            double tau = 0.0;
            double* dv = row + np;
            if (friction) {
                for (int jj = 0; jj < nd; ++jj) {
                    const double pj = (wsq == 0.0) ? 0.0 : P[pk(r, 6 + jj)] * wsq;
                    const double dqv = (wsq == 0.0) ? 0.0 : c[CX_DQ + jj];
                    const double sg = (dqv > 0.0) ? 1.0 : ((dqv < 0.0) ? -1.0 : (dqv == 0.0 ? 0.0 : dqv));   // numpy sign: sign(nan)=nan
                    dv[jj] = pj * dqv;
                    dv[nd + jj] = pj * sg;
                    tau = fma(pj, (wsq == 0.0) ? 0.0 : c[CX_TAU + jj], tau);
                }
                dv[2 * nd] = tau;
                for (int k = np + 2 * nd + 1; k < CW; ++k) row[k] = 0.0;
            } else {      // without friction columns the torque column follows the body columns directly
                for (int jj = 0; jj < nd; ++jj) {
                    const double pj = (wsq == 0.0) ? 0.0 : P[pk(r, 6 + jj)] * wsq;
                    tau = fma(pj, (wsq == 0.0) ? 0.0 : c[CX_TAU + jj], tau);
                }
                dv[0] = tau;
                for (int k = np + 1; k < CW; ++k) row[k] = 0.0;
            }
        }
    }
}

}  // namespace sysid
