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
//   * The null-space projector is P = Q Q^T with Q an orthonormal basis of null(J_c) (18 x nq, nq = 18 - rank J_c), so
//     A^T A = (Q^T Ytilde)^T (Q^T Ytilde): the Gram kernel contracts nq rows per sample (12-15 with one or two feet
//     down, 6 with four) instead of 18.  Row k of Q^T Ytilde has exactly the form above with column k of Q in place of
//     row r of P.
//
// Phases (a group of NT threads, barriers between them; `t` is the thread's index in the group):
//   stage     coalesced copy of the super-batch's channel values into shared memory
//   sincos    thread per (sample, joint); non-finite probe of the sample's column
//   chains    two lanes per (sample, leaf chain): poses relative to the base -> X_j = (R_j, p_j), a_j;  spatial
//             velocity / gravity-biased acceleration down the chain -> b9_j = (omega, alpha, acc)
//   feet      thread per (sample, stance slot, chain joint): world-aligned lever arm and leg columns of J_c
//   Gram kernel:  qbuild  16 lanes per sample: Householder QR of J_c^T (rank rule of pinv), reflectors to shared memory
//                 qcols   thread per (sample, basis vector): Q = H_0 ... H_{rank-1} [e_rank ..]
//                 fill_q  lane per (sample, basis vector, chain group) -> packed tile rows
//   rmse kernel:  sblocks / chol / wcols / proj  S = J_c J_c^T -> Cholesky -> W = L^-1 J_c -> packed P = I - W^T W
//                 fill_chains  lane per (sample, dof row, chain group) -> the 18 projected rows
//
// The contact Jacobian keeps pinocchio's exact semantics for an UN-normalised logged quaternion (float32 logs are
// off unit norm by ~3e-8; assuming an orthonormal R_b moves P by ~2e-10, above the parity gate):
//   J_k = [ R_b | -[R_b r_k]x R_b | (R_b z_c) x (R_b d_kc) ].
#pragma once
#include "kinematics.cuh"

namespace sysid {

// ---- per-sample context (doubles) ------------------------------------------------------------------------
constexpr int NPACK = MAXV * (MAXV + 1) / 2;        // 171
constexpr int NQMAX = MAXV - 3;                     // a sample with a stance foot has at most 15 null-space directions
constexpr int QLD = MAXV + 1;                       // pitch of a basis vector: odd, so the lanes of the tile fill (one vector each, same
                                                    // coordinate) read 16 distinct banks instead of colliding in pairs (pitch 18: k and k + 8)
constexpr int CX_Q = 0;                             // [NQMAX][QLD] orthonormal basis of null(J_c), one vector per row (Gram kernel)
constexpr int CX_P = CX_Q;                          // [171] packed lower triangle of P, same slot (rmse kernel)
constexpr int CX_W = CX_Q + NQMAX * QLD;            // sqrt(weight); 0 => sample contributes nothing
constexpr int CX_NQ = CX_W + 1;                     // number of basis vectors (= rows of the sample's block); MAXV with no stance foot: Q = I, not stored
constexpr int CX_A = CX_NQ + 2;                     // [MAXD][6] Pluecker axis (m; z) of every revolute joint (one pad before: 16-byte alignment)
constexpr int CX_X = CX_A + 6 * MAXD;               // [MAXD][12] R (9, row-major), p (3) relative to the base
constexpr int B9S = 10;                             // body-motion record: omega, alpha, acc + 1 pad (16-byte loads)
constexpr int CX_B9 = CX_X + 12 * MAXD;             // [MAXB][B9S] omega, alpha, acc (local frame)
constexpr int CX_DQ = CX_B9 + B9S * MAXB;           // [MAXD]
constexpr int CX_TAU = CX_DQ + MAXD;                // [MAXD]
constexpr int CX_STRIDE = CX_TAU + MAXD;            // 658 == 2 (mod 16): lanes of consecutive samples hit distinct banks
// temporaries living in the P slot until `proj` overwrites it
constexpr int CXT_S = CX_P;                         // [78] packed S, then L
constexpr int CXT_JL = CX_P + 78;                   // [MAXEE][MAXCH][3] leg columns of J_c
static_assert(78 + 3 * MAXEE * MAXCH <= NPACK && NPACK <= NQMAX * MAXV, "temporaries and the packed P fit the Q slot");
static_assert(CX_STRIDE % 16 == 2 && CX_A % 2 == 0 && CX_X % 2 == 0 && CX_B9 % 2 == 0, "context layout keeps 16-byte alignment");

// ---- per-sample scratch (doubles), only live inside the F phases -------------------------------------------
constexpr int SC_RB = 0;                            // [9] R_b from the raw quaternion
constexpr int SC_RF = SC_RB + 9;                    // [MAXEE][3] R_b r_k of the stance feet
constexpr int SC_META = SC_RF + 3 * MAXEE;          // 3m, then the foot index of every stance slot
constexpr int SC_WM = SC_META + 1 + MAXEE;          // [3*MAXEE][MAXV] W = L^-1 J_c (16-byte aligned rows)
constexpr int SC_SC = SC_WM + 3 * MAXEE * MAXV;     // [MAXD][2] sin, cos of the revolute joints
constexpr int SC_STRIDE = SC_SC + 2 * MAXD + 8;     // 274 == 2 (mod 16): lanes of consecutive samples hit distinct banks
constexpr int SC_RL = SC_WM;                        // [MAXD][10] joint rotations pR_j Rot(axis_j, q_j) (9, row-major; + 1 pad): formed by `sincos`
                                                    // (thread per joint) for the serial walks of `chains`; the W slot is not in use before `feet`
static_assert(10 * MAXD <= 3 * MAXEE * MAXV, "the joint rotations fit the W slot");
static_assert(SC_WM % 2 == 0 && SC_STRIDE % 16 == 2 && MAXV % 2 == 0, "W rows are read as double2, conflict-free across samples");

// ---- staged inputs of a super-batch (doubles): channel-major, element (channel, sample) at inp[channel * SB + sample]
constexpr int IN_Q = 0;                             // quaternion x, y, z, w, then the joint angles
constexpr int IN_DQ = IN_Q + 4 + MAXD;
constexpr int IN_DDQ = IN_DQ + MAXV;
constexpr int IN_TAU = IN_DDQ + MAXV;
constexpr int IN_CNT = IN_TAU + MAXD;
constexpr int IN_WGT = IN_CNT + MAXEE;
constexpr int IN_CHANNELS = IN_WGT + 1;             // 69

__device__ __forceinline__ int pk(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }

// Global address of element (staged channel ch, sample i), or nullptr for a channel this model / call does not have.
__device__ __forceinline__ const double* channel_ptr(const DevModel& M, const SampleIO& io, int ch, long long i) {
    const long long ld = io.ld;
    if (ch < IN_DQ) return (ch < 4 + M.nd) ? io.q + (3 + ch) * ld + i : nullptr;
    if (ch < IN_DDQ) return (ch - IN_DQ < M.nv) ? io.dq + (ch - IN_DQ) * ld + i : nullptr;
    if (ch < IN_TAU) return (ch - IN_DDQ < M.nv) ? io.ddq + (ch - IN_DDQ) * ld + i : nullptr;
    if (ch < IN_CNT) return (ch - IN_TAU < M.nd && io.tau) ? io.tau + (ch - IN_TAU) * ld + i : nullptr;
    if (ch < IN_WGT) return (ch - IN_CNT < M.n_ee) ? io.cnt + (ch - IN_CNT) * ld + i : nullptr;
    return io.weights ? io.weights + i : nullptr;
}

// L2 prefetch of the channel segments of a later super-batch (each segment is SB * 8 bytes: at most 3 lines of 128 B).
template <int SB, int NT>
__device__ __forceinline__ void prefetch_inputs(const DevModel& M, const SampleIO& io, long long base, long long N, int t) {
    constexpr int LINES = (SB * 8 + 127) / 128 + 1;
    for (int it = t; it < IN_CHANNELS * LINES; it += NT) {
        const int ch = it / LINES, ln = it - ch * LINES;
        const long long i = base + ln * 16;
        if (i >= N || ln * 16 >= SB + 15) continue;
        const double* p = channel_ptr(M, io, ch, i);
        if (p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
}

// ---------------------------------------------------------------------------------------------- stage
// Coalesced copy of the super-batch's channel values into shared memory (one global-memory latency per super-batch
// instead of one per joint of the serial chain walk).
template <int SB, int NT>
__device__ __forceinline__ void phase_stage(const DevModel& M, const SampleIO& io, long long base, long long N,
                                            double* __restrict__ inp, int t) {
    // One pass per input array with the channel count known at compile time (a generic channel -> pointer decode cost ~90
    // instructions per element); every load of the thread is issued before the first store, so the super-batch still pays one
    // global-memory latency.
    const long long ld = io.ld;
    const int nd = M.nd, nv = M.nv, n_ee = M.n_ee;
    constexpr int KQ = ((4 + MAXD) * SB + NT - 1) / NT, KV = (MAXV * SB + NT - 1) / NT, KT = (MAXD * SB + NT - 1) / NT, KC = (MAXEE * SB + NT - 1) / NT;
    double vq[KQ], vdq[KV], vddq[KV], vtau[KT], vcnt[KC];
    auto load = [&](const double* __restrict__ src, int nch, int nch_max, int k) {
        // element (ch, i) of src at src[ch * ld + i]; channels [nch, nch_max) and samples past N read as 0
        const int it = t + k * NT, ch = it / SB, sl = it - ch * SB;
        const long long i = base + sl;
        return (src && it < nch_max * SB && ch < nch && i < N) ? src[ch * ld + i] : 0.0;
    };
#pragma unroll
    for (int k = 0; k < KQ; ++k) vq[k] = load(io.q + 3 * ld, 4 + nd, 4 + MAXD, k);
#pragma unroll
    for (int k = 0; k < KV; ++k) { vdq[k] = load(io.dq, nv, MAXV, k); vddq[k] = load(io.ddq, nv, MAXV, k); }
#pragma unroll
    for (int k = 0; k < KT; ++k) vtau[k] = load(io.tau, nd, MAXD, k);
#pragma unroll
    for (int k = 0; k < KC; ++k) vcnt[k] = load(io.cnt, n_ee, MAXEE, k);
    double w = 1.0;
    if (t < SB && io.weights && base + t < N) w = io.weights[base + t];
#pragma unroll
    for (int k = 0; k < KQ; ++k) if (t + k * NT < (4 + MAXD) * SB) inp[IN_Q * SB + t + k * NT] = vq[k];
#pragma unroll
    for (int k = 0; k < KV; ++k) if (t + k * NT < MAXV * SB) { inp[IN_DQ * SB + t + k * NT] = vdq[k]; inp[IN_DDQ * SB + t + k * NT] = vddq[k]; }
#pragma unroll
    for (int k = 0; k < KT; ++k) if (t + k * NT < MAXD * SB) inp[IN_TAU * SB + t + k * NT] = vtau[k];
#pragma unroll
    for (int k = 0; k < KC; ++k) if (t + k * NT < MAXEE * SB) inp[IN_CNT * SB + t + k * NT] = vcnt[k];
    if (t < SB) inp[IN_WGT * SB + t] = w;
}

// ---------------------------------------------------------------------------------------------- sincos
// Thread per (sample, joint); copies the joint's dq / tau into the context; also probes every staged value of the sample's column for NaN/Inf (thread of joint 0).
template <int SB>
__device__ __forceinline__ void phase_sincos(const DevModel& M, long long base, long long N, const double* __restrict__ inp,
                                             double* __restrict__ ctx, double* __restrict__ scr, int* s_bad, int t) {
    if (t >= SB * MAXD) return;
    const int s = t % SB, k = t / SB;
    if (base + s >= N || k >= M.nd) return;
    ctx[s * CX_STRIDE + CX_DQ + k] = inp[(IN_DQ + 6 + k) * SB + s];
    ctx[s * CX_STRIDE + CX_TAU + k] = inp[(IN_TAU + k) * SB + s];
    double sn, cs;
    sincos(inp[(IN_Q + 4 + k) * SB + s], &sn, &cs);
    scr[s * SC_STRIDE + SC_SC + 2 * k] = sn;
    scr[s * SC_STRIDE + SC_SC + 2 * k + 1] = cs;
    {
        double Rl[9];
        joint_rotation_compose(M, k + 2, sn, cs, Rl);
        double2* dst = reinterpret_cast<double2*>(scr + s * SC_STRIDE + SC_RL + 10 * k);
        dst[0] = make_double2(Rl[0], Rl[1]); dst[1] = make_double2(Rl[2], Rl[3]); dst[2] = make_double2(Rl[4], Rl[5]);
        dst[3] = make_double2(Rl[6], Rl[7]); dst[4] = make_double2(Rl[8], 0.0);
    }
    if (k == 0) {
        double probe = 0.0;
#pragma unroll 4
        for (int ch = 0; ch < IN_CNT; ++ch) probe += inp[ch * SB + s];     // contacts and weights are not probed
        if (!(fabs(probe) < 1e300)) atomicOr(&s_bad[s], 2);
    }
}

__device__ __forceinline__ void load_joint_rotation(const double* __restrict__ src, double (&Rl)[9]) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
    const double2 a = s2[0], b = s2[1], c = s2[2], d = s2[3];
    Rl[0] = a.x; Rl[1] = a.y; Rl[2] = b.x; Rl[3] = b.y; Rl[4] = c.x; Rl[5] = c.y; Rl[6] = d.x; Rl[7] = d.y; Rl[8] = src[8];
}

// ---------------------------------------------------------------------------------------------- chains
// Two lanes per (sample, leaf chain), so that each keeps a small register state next to the Gram accumulators:
//   role 0 (pose)    R_j, p_j relative to the base down the chain -> X_j and the Pluecker axis a_j = (p_j x z_j; z_j)
//   role 1 (motion)  spatial velocity / gravity-biased acceleration in local frames -> b9_j = (omega, alpha, acc)
template <int SB>
__device__ __forceinline__ void phase_chains(const DevModel& M, long long base, long long N, const double* __restrict__ inp,
                                             double* __restrict__ ctx, double* __restrict__ scr, int t) {
    const int nlanes = SB * M.nfch, nl32 = (nlanes + 31) & ~31;      // roles start on warp boundaries: no divergence
    const int role = t / nl32, u = t - role * nl32;
    if (role > 1 || u >= nlanes) return;
    const int s = u % SB, ch = u / SB;          // sample fastest: conflict-free reads of the staged channels
    if (base + s >= N) return;
    double* c = ctx + s * CX_STRIDE;
    const double* in = inp + s;
    const double* scs = scr + s * SC_STRIDE + SC_SC;
    const int len = M.fch_len[ch], own = M.fch_own[ch];
    if (role == 0) {
        double R[9], p[3];
        for (int e = 0; e < len; ++e) {
            const int j = M.fch[ch][e];
            double Rl[9];
            load_joint_rotation(scr + s * SC_STRIDE + SC_RL + 10 * (j - 2), Rl);
            const double px = M.pp[j][0], py = M.pp[j][1], pz = M.pp[j][2];
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
            }
        }
        return;
    }
    double v[6], a[6];
    {
        double Rb[9];
        {   // Eigen::Quaternion::toRotationMatrix on the raw (x, y, z, w): no normalisation, as pinocchio's free-flyer does
            const double qx = in[(IN_Q + 0) * SB], qy = in[(IN_Q + 1) * SB], qz = in[(IN_Q + 2) * SB], qw = in[(IN_Q + 3) * SB];
            const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
            const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
            Rb[0] = 1 - (tyy + tzz); Rb[1] = txy - twz; Rb[2] = txz + twy;
            Rb[3] = txy + twz; Rb[4] = 1 - (txx + tzz); Rb[5] = tyz - twx;
            Rb[6] = txz - twy; Rb[7] = tyz + twx; Rb[8] = 1 - (txx + tyy);
        }
        // root (free-flyer): v = dq[0:6]; a = ddq[0:6] + [R_b^T (-g); 0]
        const double g0 = -M.gravity[0], g1 = -M.gravity[1], g2 = -M.gravity[2];
#pragma unroll
        for (int k = 0; k < 6; ++k) { v[k] = in[(IN_DQ + k) * SB]; a[k] = in[(IN_DDQ + k) * SB]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] += Rb[k] * g0 + Rb[3 + k] * g1 + Rb[6 + k] * g2;
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
    }
    for (int e = 0; e < len; ++e) {
        const int j = M.fch[ch][e];
        const double qd = in[(IN_DQ + 4 + j) * SB], qdd = in[(IN_DDQ + 4 + j) * SB];
        double Rl[9];
        load_joint_rotation(scr + s * SC_STRIDE + SC_RL + 10 * (j - 2), Rl);
        const double px = M.pp[j][0], py = M.pp[j][1], pz = M.pp[j][2];
        // actInv of the parent's (v, a), then the joint's own contribution
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
        if (e >= own) {
            double* b9 = c + CX_B9 + B9S * (j - 1);
            b9[0] = v[3]; b9[1] = v[4]; b9[2] = v[5];
            b9[3] = a[3]; b9[4] = a[4]; b9[5] = a[5];
            b9[6] = a[0] + (v[4] * v[2] - v[5] * v[1]);
            b9[7] = a[1] + (v[5] * v[0] - v[3] * v[2]);
            b9[8] = a[2] + (v[3] * v[1] - v[4] * v[0]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- feet
// Thread per (sample, stance slot, e): e < MAXCH -> leg column e of the slot's contact Jacobian rows,
// e == MAXCH -> the slot's world-aligned lever arm and the stance bookkeeping.
template <int SB, int JLOFF = CX_P + 78>
__device__ __forceinline__ void phase_feet(const DevModel& M, long long base, long long N, const double* __restrict__ inp,
                                           double* __restrict__ ctx, double* __restrict__ scr, int t) {
    // stance slots beyond the model's contact frames do not exist: the items are dealt over n_ee slots, so that a biped's 350
    // items fit one pass of 512 threads (nobody reads the slot table past n_ee)
    const int nee = M.n_ee;
    if (nee == 0) { if (t < SB) scr[t * SC_STRIDE + SC_META] = 0.0; return; }      // no contact frame: zero contact rows
    if (t >= SB * nee * (MAXCH + 1)) return;
    const int s = t % SB, u = t / SB, slot = u % nee, e = u / nee;
    if (base + s >= N) return;
    double* c = ctx + s * CX_STRIDE;
    double* sc = scr + s * SC_STRIDE;
    int m = 0, kf = -1;
#pragma unroll
    for (int k = 0; k < MAXEE; ++k) {
        if (k < M.n_ee) {
            const double cv = inp[(IN_CNT + k) * SB + s];
            if (cv != 0.0) {   // truthiness rule of the reference: state 2 counts as stance; NaN is truthy too
                if (m == slot) kf = k;
                ++m;
            }
        }
    }
    if (e == MAXCH) {
        if (slot == 0) sc[SC_META] = (double)(3 * m);
        sc[SC_META + 1 + slot] = (double)kf;
    }
    if (kf < 0) return;
    if (e < MAXCH && e >= M.chain_len[kf]) return;
    double Rb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rb[k] = sc[SC_RB + k];
    const int jf = M.ee_joint[kf];
    double rf[3];   // foot point in the base frame
    if (jf == 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) rf[k] = M.ee_off[kf][k];
    } else {
        const double* X = c + CX_X + 12 * (jf - 2);
#pragma unroll
        for (int k = 0; k < 3; ++k) rf[k] = X[9 + k] + X[3 * k] * M.ee_off[kf][0] + X[3 * k + 1] * M.ee_off[kf][1] + X[3 * k + 2] * M.ee_off[kf][2];
    }
    if (e == MAXCH) {
#pragma unroll
        for (int k = 0; k < 3; ++k) sc[SC_RF + 3 * slot + k] = Rb[3 * k] * rf[0] + Rb[3 * k + 1] * rf[1] + Rb[3 * k + 2] * rf[2];
        return;
    }
    const int cj = M.chain[kf][e];
    const double* A = c + CX_A + 6 * (cj - 2);
    const double* X = c + CX_X + 12 * (cj - 2);
    const double ax0 = A[3], ax1 = A[4], ax2 = A[5];
    const double bx = rf[0] - X[9], by = rf[1] - X[10], bz = rf[2] - X[11];
    const double a0 = Rb[0] * ax0 + Rb[1] * ax1 + Rb[2] * ax2, a1 = Rb[3] * ax0 + Rb[4] * ax1 + Rb[5] * ax2, a2 = Rb[6] * ax0 + Rb[7] * ax1 + Rb[8] * ax2;
    const double dx = Rb[0] * bx + Rb[1] * by + Rb[2] * bz, dy = Rb[3] * bx + Rb[4] * by + Rb[5] * bz, dz = Rb[6] * bx + Rb[7] * by + Rb[8] * bz;
    double* jl = c + JLOFF + 3 * (slot * MAXCH + e);      // CXT_JL unless the caller keeps the leg columns elsewhere
    jl[0] = a1 * dz - a2 * dy;
    jl[1] = a2 * dx - a0 * dz;
    jl[2] = a0 * dy - a1 * dx;
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
// Lane per sample, textbook left-looking loop.  (A 16-lanes-per-sample shuffle version and a statically unrolled
// right-looking version were measured slower on B200: the fp64 sqrt / divide sequences dominate, see profiles/.)
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
        // the diagonal keeps 1 / L[a][a] (rsqrt, 1-2 ulp): that is all the W columns need; 0 marks a dropped row
        if (d > piv_tol) inv = rsqrt(d);
        else { inv = 0.0; flags |= 1; }                        // row a is (numerically) dependent: drop it (pinv semantics)
        S[tri(a, a)] = inv;
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
                w[k] = val * L[tri(k, k)];                 // the diagonal holds 1 / L[k][k] (0 for a dropped row: W row = 0)
                sc[SC_WM + k * MAXV + col] = w[k];
            }
        }
    }
}

template <int SB>
__device__ __forceinline__ void phase_finish(long long base, long long N, const double* __restrict__ inp, double* __restrict__ ctx,
                                             const int* s_bad, int t, double* s_stat);

// ---------------------------------------------------------------------------------------------- null-space basis
// G = sum A^T A with A = P Ytilde and P = Q Q^T (Q: orthonormal basis of null(J_c), 18 x nq, nq = 18 - rank J_c), so
// A^T A = (Q^T Ytilde)^T (Q^T Ytilde): the Gram kernel contracts nq rows per sample instead of 18 -- 12..15 with one or
// two feet down, 6 with four.  W = L^-1 J_c has orthonormal rows (zero rows where a dependent contact row was dropped);
// Householder reflectors H_0 .. H_{rank-1} map them onto e_0 .. e_{rank-1}, and Q = H_0 ... H_{rank-1} [e_rank .. e_17].
// qbuild: lane per sample, reflector t overwrites W row t.   qcols: thread per (sample, basis vector).
// Householder QR of J_c^T, sixteen lanes per sample (two samples per warp): lane b holds contact row b of J_c (built
// from R_b, the world-aligned lever arms and the leg columns of the `feet` phase) in registers.  For pivot position p
// the next row whose remainder below p is not negligible (|.|^2 > 1e-13 max_b |J_b|^2: the pinv rank rule, the same
// quantity the Cholesky pivot of J_c J_c^T measures) becomes reflector p and is published in shared memory; later rows
// apply it.  No S = J J^T, no Cholesky, no W: the basis comes straight from J_c (and is better conditioned for it).
// `t` must cover whole warps: [0, 16 SB) rounded up to a multiple of 32.
template <int SB>
__device__ __forceinline__ void phase_qbuild(const DevModel& M, long long base, long long N, double* __restrict__ ctx,
                                             double* __restrict__ scr, int* s_bad, int t) {
    constexpr int MR = 3 * MAXEE;
    static_assert(MR <= 16, "one half-warp per sample");
    const unsigned full = 0xffffffffu;
    const int s = t >> 4, b = t & 15, hbase = threadIdx.x & 16;
    const bool live = (s < SB) && (base + s < N);
    double* sc = scr + (live ? s : 0) * SC_STRIDE;
    const double* c = ctx + (live ? s : 0) * CX_STRIDE;
    const int m3 = live ? (int)sc[SC_META] : 0;
    double x[MAXV];
#pragma unroll
    for (int r = 0; r < MAXV; ++r) x[r] = 0.0;
    double n2 = 0.0;
    if (b < m3) {
        // row (slot, xx) of J_c = [ R_b | -[r]x R_b | leg columns ], assembled through this lane's own (dead) W row slot
        const int slot = b / 3, xx = b - 3 * slot;
        const int kt = (int)sc[SC_META + 1 + slot];
        double* mine = sc + SC_WM + b * MAXV;
        const double r0 = sc[SC_RF + 3 * slot], r1 = sc[SC_RF + 3 * slot + 1], r2 = sc[SC_RF + 3 * slot + 2];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const double b0 = sc[SC_RB + cc], b1 = sc[SC_RB + 3 + cc], b2 = sc[SC_RB + 6 + cc];
            mine[cc] = (xx == 0) ? b0 : ((xx == 1) ? b1 : b2);
            mine[3 + cc] = -((xx == 0) ? (r1 * b2 - r2 * b1) : ((xx == 1) ? (r2 * b0 - r0 * b2) : (r0 * b1 - r1 * b0)));
        }
#pragma unroll
        for (int cc = 6; cc < MAXV; ++cc) mine[cc] = 0.0;
        const int len = M.chain_len[kt];
        for (int e = 0; e < len; ++e) mine[4 + M.chain[kt][e]] = c[CXT_JL + 3 * (slot * MAXCH + e) + xx];
#pragma unroll
        for (int r = 0; r < MAXV; ++r) { x[r] = mine[r]; n2 = fma(x[r], x[r], n2); }
    }
    double mx = n2;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(full, mx, o));
    const double tol = 1e-13 * mx;
    __syncwarp();                                            // every lane holds its row: the slots may now take reflectors
    int cur = 0, rank = 0, dropped = 0;                      // uniform per half-warp
#pragma unroll
    for (int p = 0; p < MR; ++p) {
        // next row whose remainder below position p is not negligible
        bool found = false;
        while (__any_sync(full, !found && cur < m3)) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int r = p; r < MAXV; ++r) { if (r % 3 == 0) t0 = fma(x[r], x[r], t0); else if (r % 3 == 1) t1 = fma(x[r], x[r], t1); else t2 = fma(x[r], x[r], t2); }
            const double tail2 = __shfl_sync(full, t0 + t1 + t2, hbase + min(cur, 15));
            if (!found && cur < m3) {
                if (tail2 > tol) found = true;
                else { ++cur; dropped = 1; }                // dependent row: dropped (pinv semantics)
            }
        }
        if (__any_sync(full, found)) {
            double* vrow = sc + SC_WM + p * MAXV;
            if (found && b == cur) {
                double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int r = p; r < MAXV; ++r) { if (r % 3 == 0) t0 = fma(x[r], x[r], t0); else if (r % 3 == 1) t1 = fma(x[r], x[r], t1); else t2 = fma(x[r], x[r], t2); }
                const double tail2 = t0 + t1 + t2;
                const double nt = tail2 * rsqrt(tail2);
                const double alpha = (x[p] >= 0.0) ? nt : -nt;
                const double inv = rsqrt(2.0 * (tail2 + fabs(x[p]) * nt));       // 1 / |x[p:] + alpha e_p|
#pragma unroll
                for (int r = 0; r < MAXV; ++r) vrow[r] = (r < p) ? 0.0 : ((r == p) ? (x[r] + alpha) * inv : x[r] * inv);
            }
            __syncwarp();
            if (found && b > cur && b < m3) {
                // the reflector is streamed from shared memory twice (dot, then update) instead of being held in registers
                const double2* v2 = reinterpret_cast<const double2*>(vrow);
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int r2 = p / 2; r2 < MAXV / 2; ++r2) { const double2 q = v2[r2]; d0 = fma(q.x, x[2 * r2], d0); d1 = fma(q.y, x[2 * r2 + 1], d1); }
                const double d = -2.0 * (d0 + d1);
#pragma unroll
                for (int r2 = p / 2; r2 < MAXV / 2; ++r2) { const double2 q = v2[r2]; x[2 * r2] = fma(d, q.x, x[2 * r2]); x[2 * r2 + 1] = fma(d, q.y, x[2 * r2 + 1]); }
            }
            if (found) { ++cur; ++rank; }
        }
    }
    if (b == 0 && s < SB) {
        if (live) { sc[SC_META] = (double)rank; if (dropped) atomicOr(&s_bad[s], 1); }   // SC_META: from here on the number of reflectors
        ctx[s * CX_STRIDE + CX_NQ] = live ? (double)(MAXV - rank) : 0.0;
    }
}

template <int SB>
__device__ __forceinline__ void phase_qcols(long long base, long long N, double* __restrict__ ctx, const double* __restrict__ scr, int t) {
    if (t >= SB * NQMAX) return;
    const int s = t / NQMAX, k = t - s * NQMAX;              // vector index fastest: a sample's lanes read the same reflector (broadcast)
    if (base + s >= N) return;
    const double* sc = scr + s * SC_STRIDE;
    const int rank = (int)sc[SC_META];
    if (rank == 0 || k >= MAXV - rank) return;               // no stance foot: Q = I is implicit
    double x[MAXV];
#pragma unroll
    for (int r = 0; r < MAXV; ++r) x[r] = (r == rank + k) ? 1.0 : 0.0;
    for (int p = rank - 1; p >= 0; --p) {                    // x <- H_p x, last reflector first (streamed twice: no register copy)
        const double2* v2 = reinterpret_cast<const double2*>(sc + SC_WM + p * MAXV);
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int r2 = 0; r2 < MAXV / 2; ++r2) { const double2 q = v2[r2]; d0 = fma(q.x, x[2 * r2], d0); d1 = fma(q.y, x[2 * r2 + 1], d1); }
        const double d = -2.0 * (d0 + d1);
#pragma unroll
        for (int r2 = 0; r2 < MAXV / 2; ++r2) { const double2 q = v2[r2]; x[2 * r2] = fma(d, q.x, x[2 * r2]); x[2 * r2 + 1] = fma(d, q.y, x[2 * r2 + 1]); }
    }
    double* qk = ctx + s * CX_STRIDE + CX_Q + k * QLD;       // odd pitch: scalar stores, conflict-free across the sample's lanes
#pragma unroll
    for (int r = 0; r < MAXV; ++r) qk[r] = x[r];
}

// ---------------------------------------------------------------------------------------------- P = I - W^T W
// Thread per (sample, row r): entries (r, 0..r) of the packed lower triangle, one independent accumulator per entry.
// NT = threads in the group; also finalises the per-sample weight and the skip flags (first warp of the group).
template <int SB, int NT>
__device__ __forceinline__ void phase_proj(long long base, long long N, const double* __restrict__ inp, double* __restrict__ ctx,
                                           const double* __restrict__ scr, const int* s_bad, int t, double* s_stat) {
    for (int it = t; it < SB * MAXV; it += NT) {
        const int s = it / MAXV, r = it - s * MAXV;         // row fastest: the lanes of a sample read the same W row (broadcast)
        if (base + s >= N) continue;
        const double* sc = scr + s * SC_STRIDE;
        const int m3 = (int)sc[SC_META];
        double pv[MAXV];
#pragma unroll
        for (int cc = 0; cc < MAXV; ++cc) pv[cc] = (cc == r) ? 1.0 : 0.0;
        for (int k = 0; k < m3; ++k) {
            const double2* wrow = reinterpret_cast<const double2*>(sc + SC_WM + k * MAXV);
            const double wk = -sc[SC_WM + k * MAXV + r];
#pragma unroll
            for (int c2 = 0; c2 < MAXV / 2; ++c2) {
                const double2 w2 = wrow[c2];
                pv[2 * c2] = fma(wk, w2.x, pv[2 * c2]);
                pv[2 * c2 + 1] = fma(wk, w2.y, pv[2 * c2 + 1]);
            }
        }
        double* Pr = ctx + s * CX_STRIDE + CX_P + r * (r + 1) / 2;
#pragma unroll
        for (int cc = 0; cc < MAXV; ++cc) if (cc <= r) Pr[cc] = pv[cc];
    }
    phase_finish<SB>(base, N, inp, ctx, s_bad, t, s_stat);
}

// Per-sample weight and skip flags (first warp of the group): deterministic (shuffle-tree) sums of the weights and counts.
template <int SB>
__device__ __forceinline__ void phase_finish(long long base, long long N, const double* __restrict__ inp, double* __restrict__ ctx,
                                             const int* s_bad, int t, double* s_stat) {
    static_assert(SB <= 32, "the first warp of the group finalises the weights");
    if (t < 32) {
        const long long i = base + t;
        double w = 0.0;
        int f0 = 0, f1 = 0;
        if (t < SB && i < N) {
            const int bad = s_bad[t];
            w = inp[IN_WGT * SB + t];
            if (!(fabs(w) < 1e300) || (bad & 2)) { w = 0.0; f1 = 1; }
            f0 = bad & 1;
            w = fmax(w, 0.0);
        }
        if (t < SB) ctx[t * CX_STRIDE + CX_W] = sqrt(w);
        const unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (t == 0) { s_stat[0] += w; s_stat[1] += (double)__popc(b0); s_stat[2] += (double)__popc(b1); }
    }
}

// ---------------------------------------------------------------------------------------------- tile fill
// Ten entries of one projected row of body i from its body-frame coefficients (el; ea) = bodyRegressor^T (el; ea).
__device__ __forceinline__ void body_row(const double* __restrict__ b9, double el0, double el1, double el2,
                                         double ea0, double ea1, double ea2, double* __restrict__ dst) {
    const double2* b2 = reinterpret_cast<const double2*>(b9);
    const double2 q0 = b2[0], q1 = b2[1], q2 = b2[2], q3 = b2[3];
    const double w0 = q0.x, w1 = q0.y, w2 = q1.x, al0 = q1.y, al1 = q2.x, al2 = q2.y, ac0 = q3.x, ac1 = q3.y, ac2 = b9[8];
    double o[10];
    // mass column: el . acc
    o[0] = el0 * ac0 + el1 * ac1 + el2 * ac2;
    // first-moment columns: -alpha x el + omega x (omega x el) + acc x ea
    const double u0 = w1 * el2 - w2 * el1, u1 = w2 * el0 - w0 * el2, u2 = w0 * el1 - w1 * el0;
    o[1] = (w1 * u2 - w2 * u1) - (al1 * el2 - al2 * el1) + (ac1 * ea2 - ac2 * ea1);
    o[2] = (w2 * u0 - w0 * u2) - (al2 * el0 - al0 * el2) + (ac2 * ea0 - ac0 * ea2);
    o[3] = (w0 * u1 - w1 * u0) - (al0 * el1 - al1 * el0) + (ac0 * ea1 - ac1 * ea0);
    // inertia columns (Ixx, Ixy, Iyy, Ixz, Iyz, Izz): ea . Br(alpha)[:,k] + (ea x omega) . Br(omega)[:,k]
    const double g0 = ea1 * w2 - ea2 * w1, g1 = ea2 * w0 - ea0 * w2, g2 = ea0 * w1 - ea1 * w0;
    o[4] = ea0 * al0 + g0 * w0;
    o[5] = ea0 * al1 + ea1 * al0 + g0 * w1 + g1 * w0;
    o[6] = ea1 * al1 + g1 * w1;
    o[7] = ea0 * al2 + ea2 * al0 + g0 * w2 + g2 * w0;
    o[8] = ea1 * al2 + ea2 * al1 + g1 * w2 + g2 * w1;
    o[9] = ea2 * al2 + g2 * w2;
    double2* d2 = reinterpret_cast<double2*>(dst);      // 10 * body and the row pitch are even: 16-byte aligned
#pragma unroll
    for (int k = 0; k < 5; ++k) d2[k] = make_double2(o[2 * k], o[2 * k + 1]);
}

// ---------------------------------------------------------------------------------------------- tile fill, null-space rows
// Lanes are grouped by (leaf chain, split index): a lane owns one (sample, basis vector q_k) of its group, walks the chain
// from the root accumulating (dl; da) = q_k^T [I_6; a_j ...] incrementally (6 FMAs per joint instead of re-summing the
// ancestors per body) and emits the ten entries of every body e of the chain with e % split == its split index.  One
// more group emits the root body and the friction / torque / padding columns.  Groups are padded to whole warps so that
// a warp never diverges.  Its tile row is q_k^T Ytilde, so a sample contributes nq = 18 - rank(J_c) rows, packed one
// after the other.  Returns (to every
// thread) the number of k-steps of 4 rows the round occupies; rows up to that multiple of 4 are zero-filled.
template <int TS, int LD, int NT>
__device__ __forceinline__ int phase_fill_q(const DevModel& M, const double* __restrict__ ctx, double* __restrict__ tile,
                                            int s0, int friction, int t) {
    constexpr int GL = ((TS * MAXV + 31) / 32) * 32;     // lanes per group
    static_assert(NT % 32 == 0 && GL <= NT, "groups are whole warps");
    const int split = M.fill_split, ngroups = M.nfch * split;
    const int np = M.nparams, nd = M.nd;
    // row offsets of the round's samples
    int off[TS + 1];
    off[0] = 0;
#pragma unroll
    for (int u = 0; u < TS; ++u) {
        const double* cu = ctx + (s0 + u) * CX_STRIDE;
        off[u + 1] = off[u] + ((cu[CX_W] == 0.0) ? 0 : (int)cu[CX_NQ]);
    }
    const int rows = off[TS], ksteps = (rows + 3) >> 2;
    for (int e = t; e < (4 * ksteps - rows) * CW; e += NT) tile[(rows + e / CW) * LD + (e % CW)] = 0.0;   // pad rows of the last k-step
    const int l = t % GL;
    if (l >= TS * MAXV) return ksteps;
    // basis-vector index slowest: a stance sample has at most 15 vectors, so slots 4 k + sl with k >= 15 (the group's third
    // warp) are all idle unless a sample of the round is in flight, and that warp retires at once instead of running the
    // whole walk for the five lanes a sample-major order left in it
    const int k = l / TS, sl = l - k * TS;
    int rowi = 0, nq = 0;
#pragma unroll
    for (int u = 0; u < TS; ++u) if (u == sl) { rowi = off[u] + k; nq = off[u + 1] - off[u]; }
    if (k >= nq) return ksteps;
    const double* c = ctx + (s0 + sl) * CX_STRIDE;
    const double wsq = c[CX_W];
    const bool ident = (nq == MAXV);                      // no stance foot: Q = I
    const double* Qk = c + CX_Q + k * QLD;
    double* row = tile + rowi * LD;
    auto Qe = [&](int cc) { return ident ? ((cc == k) ? 1.0 : 0.0) : Qk[cc]; };
    for (int g = t / GL; g <= ngroups; g += NT / GL) {
        if (g < ngroups) {
            const int ch = (split == 1) ? g : (g >> 1), q = (split == 1) ? 0 : (g & 1);      // split is 1 or 2
            const int len = M.fch_len[ch], own = M.fch_own[ch];
            double d[6];           // (dl; da) = q_k^T [I_6; a_j ...], pre-multiplied by sqrt(weight)
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) d[cc] = Qe(cc) * wsq;
            for (int e = 0, turn = 0; e < len; ++e, turn = (turn + 1 == split) ? 0 : turn + 1) {
                const int j = M.fch[ch][e];
                const double pj = Qe(4 + j) * wsq;
                const double2* A2 = reinterpret_cast<const double2*>(c + CX_A + 6 * (j - 2));
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) { const double2 ak = A2[cc]; d[2 * cc] = fma(pj, ak.x, d[2 * cc]); d[2 * cc + 1] = fma(pj, ak.y, d[2 * cc + 1]); }
                if (e >= own && turn == q) {
                    const double2* X2 = reinterpret_cast<const double2*>(c + CX_X + 12 * (j - 2));
                    const double2 x01 = X2[0], x23 = X2[1], x45 = X2[2], x67 = X2[3], x8p = X2[4], p12 = X2[5];
                    const double p0 = x8p.y, p1 = p12.x, p2 = p12.y;
                    const double u0 = d[0] + (d[4] * p2 - d[5] * p1), u1 = d[1] + (d[5] * p0 - d[3] * p2), u2 = d[2] + (d[3] * p1 - d[4] * p0);
                    // (el; ea) = (R^T u; R^T da), R row-major in x01..x8p
                    const double el0 = x01.x * u0 + x23.y * u1 + x67.x * u2, el1 = x01.y * u0 + x45.x * u1 + x67.y * u2, el2 = x23.x * u0 + x45.y * u1 + x8p.x * u2;
                    const double ea0 = x01.x * d[3] + x23.y * d[4] + x67.x * d[5], ea1 = x01.y * d[3] + x45.x * d[4] + x67.y * d[5], ea2 = x23.x * d[3] + x45.y * d[4] + x8p.x * d[5];
                    body_row(c + CX_B9 + B9S * (j - 1), el0, el1, el2, ea0, ea1, ea2, row + 10 * (j - 1));
                }
            }
        } else {
            // root body (its Pluecker rows are the identity, pose = identity) ...
            body_row(c + CX_B9, Qe(0) * wsq, Qe(1) * wsq, Qe(2) * wsq, Qe(3) * wsq, Qe(4) * wsq, Qe(5) * wsq, row);
            // ... and the friction / torque columns plus the zero padding, as 16-byte stores where the layout allows
            // (columns np.., np + nd.. and the torque column all start on even offsets for the robots of the envelope)
            const int ntail = friction ? 2 * nd + 1 : 1;
            if (((np | nd) & 1) == 0) {
                double tau = 0.0;
                for (int jj = 0; jj < nd; jj += 2) {
                    const double p0 = Qe(6 + jj) * wsq, p1 = Qe(7 + jj) * wsq;
                    const double2 tq = *reinterpret_cast<const double2*>(c + CX_TAU + jj);
                    tau = fma(p0, tq.x, tau); tau = fma(p1, tq.y, tau);
                    if (friction) {
                        const double2 dv = *reinterpret_cast<const double2*>(c + CX_DQ + jj);
                        const double s0v = (dv.x > 0.0) ? 1.0 : ((dv.x < 0.0) ? -1.0 : (dv.x == 0.0 ? 0.0 : dv.x));   // numpy sign: sign(nan)=nan
                        const double s1v = (dv.y > 0.0) ? 1.0 : ((dv.y < 0.0) ? -1.0 : (dv.y == 0.0 ? 0.0 : dv.y));
                        *reinterpret_cast<double2*>(row + np + jj) = make_double2(p0 * dv.x, p1 * dv.y);
                        *reinterpret_cast<double2*>(row + np + nd + jj) = make_double2(p0 * s0v, p1 * s1v);
                    }
                }
                // the torque column sits on an even offset; everything after it is padding
                *reinterpret_cast<double2*>(row + np + ntail - 1) = make_double2(tau, 0.0);
                for (int cc = np + ntail + 1; cc < CW; cc += 2) *reinterpret_cast<double2*>(row + cc) = make_double2(0.0, 0.0);
            } else {
                double tau = 0.0;
                for (int jj = 0; jj < nd; ++jj) {
                    const double pj = Qe(6 + jj) * wsq;
                    tau = fma(pj, c[CX_TAU + jj], tau);
                    if (friction) {
                        const double dqv = c[CX_DQ + jj];
                        const double sg = (dqv > 0.0) ? 1.0 : ((dqv < 0.0) ? -1.0 : (dqv == 0.0 ? 0.0 : dqv));   // numpy sign: sign(nan)=nan
                        row[np + jj] = pj * dqv;
                        row[np + nd + jj] = pj * sg;
                    }
                }
                // without friction columns the torque column follows the body columns directly
                row[np + ntail - 1] = tau;
                for (int cc = np + ntail; cc < CW; ++cc) row[cc] = 0.0;
            }
        }
    }
    return ksteps;
}

}  // namespace sysid
