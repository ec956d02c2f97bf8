// Stages 1-2 for trees OUTSIDE the fused kernel's compile-time envelope (13 bodies, nv <= 18, chains <= 6): Unitree G1-29dof
// (files/g1_description/g1_29dof.urdf: 30 bodies, nv = 35, c = 358, chains of 10) and its lock-waist variant.
//
// At c = 358 the Gram is 1 035 8x8 tiles -- 530 KB of accumulators, more than one SM's registers and tensor memory together --
// so the single-CTA-per-SM fusion of gram_kernels.cuh does not carry over.  This path is the same mathematics in the plain
// two-step form, chunk by chunk (8 192 samples) through HBM -- rows of chunk i + 1 beside the SYRK of chunk i -- generic in the tree size:
//
//   kin        thread per sample            forward kinematics: local / world placements, spatial velocity and gravity-biased
//                                           acceleration of every joint (pinocchio's first loop)
//   rows       thread per (sample, body)    bodyRegressor, walked up the ancestors: the UNPROJECTED rows Ytilde (nv x 384 per sample)
//   tail       thread per (sample, joint)   friction and torque columns of Ytilde
//   contact    thread per sample            J_c (LOCAL_WORLD_ALIGNED, raw-quaternion R_b), S = J J^T, Cholesky with the pinv rank
//                                           rule, W = L^-1 J_c  (P = I - W^T W)
//   zrows      thread per (sample, column)  Z = W Ytilde  (3 n_ee rows per sample at once; structurally zero rows of the column skipped)
//   syrk_rows  DMMA, 64 x 64 output blocks  partial Grams of  Ytilde^T Ytilde - Z^T Z  =  Ytilde^T P Ytilde, rows class by class on tile masks
//   reduce     thread per element           partials -> stats = [G | r | s | n], fixed summation order (bit-reproducible)
//
// The torque column rides along as column c, like in the fused kernel.  Not here (refused with SYSID_ERR_UNSUPPORTED for these
// models): per-sample weights, NaN skipping, the segmented (bootstrap) mode.
#pragma once
#include "kinematics.cuh"

namespace sysid {
namespace big {

constexpr int BJ = 40;        // joints including the universe
constexpr int BV = 40;        // generalised velocities
constexpr int BEE = 4;        // contact frames
constexpr int BMR = 3 * BEE;  // contact rows
constexpr int BCW = 384;      // padded row width: c + 1 <= 384, six 64-column blocks
constexpr int KIN = 36;       // kinematics record per joint: liR(9) lip(3) oR(9) op(3) v(6) a(6)
constexpr int SY_BLK = 64, SY_NB = BCW / SY_BLK, SY_NBLK = SY_NB * (SY_NB + 1) / 2;   // 21 lower-triangular output blocks
constexpr int SY_ROWS = 32, SY_LD = SY_BLK + 4;                                        // rows per staged panel; pitch == 4 (mod 16)
constexpr int SY_THREADS = 256;

struct BigModel {
    int32_t njoints, nb, nv, nq, nd, n_ee, nparams, pad_;
    int32_t parent[BJ], jtype[BJ], idx_v[BJ], idx_q[BJ];
    double axis[BJ][3], pR[BJ][9], pp[BJ][3], gravity[3];
    int32_t ee_joint[BEE];
    double ee_off[BEE][3];
};

// one bit per 8-column tile of a row of Ytilde that the row can touch at all (big_row_masks, below)
struct RowMasks { unsigned long long m[BV]; };

__device__ __forceinline__ void mat3mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3vec(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
__device__ __forceinline__ void mat3Tvec(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
__device__ __forceinline__ void cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void joint_axis(const BigModel& M, int j, double* ax) {
    const int jt = M.jtype[j];
    ax[0] = (jt == JT_RX) ? 1.0 : ((jt == JT_RU) ? M.axis[j][0] : 0.0);
    ax[1] = (jt == JT_RY) ? 1.0 : ((jt == JT_RU) ? M.axis[j][1] : 0.0);
    ax[2] = (jt == JT_RZ) ? 1.0 : ((jt == JT_RU) ? M.axis[j][2] : 0.0);
}

// ---------------------------------------------------------------------------------------------- kin
__global__ void big_kin_kernel(const __grid_constant__ BigModel M, const SampleIO io, long long base, int ns, double* __restrict__ kin) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const long long i = base + s, ld = io.ld;
    double* K = kin + (size_t)s * M.njoints * KIN;
    // universe: identity placement, zero velocity, acceleration (-g; 0)
    for (int k = 0; k < KIN; ++k) K[k] = 0.0;
    K[12] = K[16] = K[20] = 1.0;
    for (int k = 0; k < 3; ++k) K[30 + k] = -M.gravity[k];
    for (int j = 1; j < M.njoints; ++j) {
        const int lam = M.parent[j], iq = M.idx_q[j], iv = M.idx_v[j], jt = M.jtype[j];
        const double* Kp = K + (size_t)lam * KIN;
        double* Kj = K + (size_t)j * KIN;
        double Rj[9], pj[3] = {0.0, 0.0, 0.0}, vJ[6] = {0, 0, 0, 0, 0, 0}, aJ[6] = {0, 0, 0, 0, 0, 0};
        if (jt == JT_FF) {
            // Eigen::Quaternion::toRotationMatrix on the raw (x, y, z, w): no normalisation, as pinocchio's free-flyer does
            const double qx = io.q[(iq + 3) * ld + i], qy = io.q[(iq + 4) * ld + i], qz = io.q[(iq + 5) * ld + i], qw = io.q[(iq + 6) * ld + i];
            const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
            const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
            Rj[0] = 1 - (tyy + tzz); Rj[1] = txy - twz; Rj[2] = txz + twy;
            Rj[3] = txy + twz; Rj[4] = 1 - (txx + tzz); Rj[5] = tyz - twx;
            Rj[6] = txz - twy; Rj[7] = tyz + twx; Rj[8] = 1 - (txx + tyy);
            for (int k = 0; k < 3; ++k) pj[k] = io.q[(iq + k) * ld + i];
            for (int k = 0; k < 6; ++k) { vJ[k] = io.dq[(iv + k) * ld + i]; aJ[k] = io.ddq[(iv + k) * ld + i]; }
        } else {
            double sn, cs, ax[3];
            sincos(io.q[iq * ld + i], &sn, &cs);
            joint_axis(M, j, ax);
            const double t = 1.0 - cs, ux = ax[0], uy = ax[1], uz = ax[2];
            Rj[0] = 1.0 - t * (uy * uy + uz * uz); Rj[1] = t * ux * uy - sn * uz; Rj[2] = t * ux * uz + sn * uy;
            Rj[3] = t * ux * uy + sn * uz; Rj[4] = 1.0 - t * (ux * ux + uz * uz); Rj[5] = t * uy * uz - sn * ux;
            Rj[6] = t * ux * uz - sn * uy; Rj[7] = t * uy * uz + sn * ux; Rj[8] = 1.0 - t * (ux * ux + uy * uy);
            const double qd = io.dq[iv * ld + i], qdd = io.ddq[iv * ld + i];
            for (int k = 0; k < 3; ++k) { vJ[3 + k] = ax[k] * qd; aJ[3 + k] = ax[k] * qdd; }
        }
        double liR[9], lip[3], tmp[3];
        mat3mul(M.pR[j], Rj, liR);
        mat3vec(M.pR[j], pj, tmp);
        for (int k = 0; k < 3; ++k) lip[k] = M.pp[j][k] + tmp[k];
        double oR[9], op[3];
        mat3mul(Kp + 12, liR, oR);
        mat3vec(Kp + 12, lip, tmp);
        for (int k = 0; k < 3; ++k) op[k] = Kp[21 + k] + tmp[k];
        // v = vJ + liMi.actInv(v_parent) (parent = universe: zero);  a = v x vJ + aJ + liMi.actInv(a_parent)
        double v[6], a[6], c1[3], d[3], pv[6] = {0, 0, 0, 0, 0, 0}, pa[6];
        if (lam > 0) {
            cross(lip, Kp + 24 + 3, c1);
            for (int k = 0; k < 3; ++k) d[k] = Kp[24 + k] - c1[k];
            mat3Tvec(liR, d, pv); mat3Tvec(liR, Kp + 24 + 3, pv + 3);
        }
        cross(lip, Kp + 30 + 3, c1);
        for (int k = 0; k < 3; ++k) d[k] = Kp[30 + k] - c1[k];
        mat3Tvec(liR, d, pa); mat3Tvec(liR, Kp + 30 + 3, pa + 3);
        for (int k = 0; k < 6; ++k) v[k] = vJ[k] + pv[k];
        double x1[3], x2[3], x3[3];
        cross(v + 3, vJ, x1); cross(v, vJ + 3, x2); cross(v + 3, vJ + 3, x3);      // (v, w) x (v2, w2) = (w x v2 + v x w2 ; w x w2)
        for (int k = 0; k < 3; ++k) { a[k] = x1[k] + x2[k] + aJ[k] + pa[k]; a[3 + k] = x3[k] + aJ[3 + k] + pa[3 + k]; }
        for (int k = 0; k < 9; ++k) { Kj[k] = liR[k]; Kj[12 + k] = oR[k]; }
        for (int k = 0; k < 3; ++k) { Kj[9 + k] = lip[k]; Kj[21 + k] = op[k]; }
        for (int k = 0; k < 6; ++k) { Kj[24 + k] = v[k]; Kj[30 + k] = a[k]; }
    }
}

// ---------------------------------------------------------------------------------------------- rows
// Ytilde of one chunk: [ns][nv][BCW], zeroed by the caller.  Body b = joint j: its ten columns in every ancestor's row(s).
__global__ void big_rows_kernel(const __grid_constant__ BigModel M, int ns, const double* __restrict__ kin, double* __restrict__ Yt) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ns * M.nb) return;
    const int s = e / M.nb, i = 1 + e % M.nb;
    const double* K = kin + (size_t)s * M.njoints * KIN;
    double* Y = Yt + (size_t)s * M.nv * BCW + 10 * (i - 1);
    const double* vi = K + (size_t)i * KIN + 24;
    const double* ai = K + (size_t)i * KIN + 30;
    const double* w = vi + 3; const double* al = ai + 3;
    double acc[3], wxv[3];
    cross(w, vi, wxv);
    for (int k = 0; k < 3; ++k) acc[k] = ai[k] + wxv[k];
    // bodyRegressor(v, a): B (6 x 10), columns [m | h (3) | Ixx, Ixy, Iyy, Ixz, Iyz, Izz]
    double B[6][10];
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 10; ++c) B[r][c] = 0.0;
    for (int k = 0; k < 3; ++k) B[k][0] = acc[k];
    const double Sa[9] = {0, -al[2], al[1], al[2], 0, -al[0], -al[1], al[0], 0};
    const double Sw[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double Sww[9];
    mat3mul(Sw, Sw, Sww);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) B[r][1 + c] = Sa[3 * r + c] + Sww[3 * r + c];
    const double Sacc[9] = {0, -acc[2], acc[1], acc[2], 0, -acc[0], -acc[1], acc[0], 0};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) B[3 + r][1 + c] = -Sacc[3 * r + c];
    const double Bra[3][6] = {{al[0], al[1], 0, al[2], 0, 0}, {0, al[0], al[1], 0, al[2], 0}, {0, 0, 0, al[0], al[1], al[2]}};
    const double Brw[3][6] = {{w[0], w[1], 0, w[2], 0, 0}, {0, w[0], w[1], 0, w[2], 0}, {0, 0, 0, w[0], w[1], w[2]}};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 6; ++c) {
            double t = Bra[r][c];
            for (int k = 0; k < 3; ++k) t += Sw[3 * r + k] * Brw[k][c];
            B[3 + r][4 + c] = t;
        }
    for (int j = i; j > 0; j = M.parent[j]) {
        const int iv = M.idx_v[j];
        if (M.jtype[j] == JT_FF) {
            for (int r = 0; r < 6; ++r) for (int c = 0; c < 10; ++c) Y[(size_t)(iv + r) * BCW + c] = B[r][c];
        } else {
            double ax[3];
            joint_axis(M, j, ax);
            for (int c = 0; c < 10; ++c) Y[(size_t)iv * BCW + c] = ax[0] * B[3][c] + ax[1] * B[4][c] + ax[2] * B[5][c];
        }
        if (M.parent[j] > 0) {
            const double* R = K + (size_t)j * KIN; const double* p = R + 9;       // liMi[j].act(B): f' = R f, n' = R n + p x f'
            for (int c = 0; c < 10; ++c) {
                const double f[3] = {B[0][c], B[1][c], B[2][c]}, n3[3] = {B[3][c], B[4][c], B[5][c]};
                double f2[3], n2[3], cx[3];
                mat3vec(R, f, f2); mat3vec(R, n3, n2); cross(p, f2, cx);
                for (int k = 0; k < 3; ++k) { B[k][c] = f2[k]; B[3 + k][c] = n2[k] + cx[k]; }
            }
        }
    }
}

// friction / torque columns: Ytilde[6 + k][np + k] = dq_k, [np + nd + k] = sign(dq_k), [c] = tau_k
__global__ void big_tail_kernel(const __grid_constant__ BigModel M, const SampleIO io, long long base, int ns, int friction, double* __restrict__ Yt) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ns * M.nd) return;
    const int s = e / M.nd, k = e % M.nd;
    const long long i = base + s;
    double* row = Yt + ((size_t)s * M.nv + 6 + k) * BCW;
    const int np = M.nparams, c = np + (friction ? 2 * M.nd : 0);
    if (friction) {
        const double dv = io.dq[(6 + k) * io.ld + i];
        row[np + k] = dv;
        row[np + M.nd + k] = (dv > 0.0) ? 1.0 : ((dv < 0.0) ? -1.0 : (dv == 0.0 ? 0.0 : dv));       // numpy sign
    }
    row[c] = io.tau ? io.tau[k * io.ld + i] : 0.0;
}

// ---------------------------------------------------------------------------------------------- contact
// W[s][BMR][BV] = L^-1 J_c (zero rows for dropped / absent contact rows), m3[s] = number of contact rows.
__global__ void big_contact_kernel(const __grid_constant__ BigModel M, const SampleIO io, long long base, int ns, const double* __restrict__ kin,
                                   double* __restrict__ W, int* __restrict__ m3_out, int* __restrict__ rankloss) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const long long i = base + s;
    const double* K = kin + (size_t)s * M.njoints * KIN;
    const int nv = M.nv;
    double* Ws = W + (size_t)s * BMR * BV;
    for (int k = 0; k < BMR * BV; ++k) Ws[k] = 0.0;
    int m3 = 0;
    for (int f = 0; f < M.n_ee; ++f) {
        if (io.cnt[f * io.ld + i] == 0.0) continue;             // truthiness rule of the reference: state 2 counts as stance
        const int jf = M.ee_joint[f];
        const double* Kf = K + (size_t)jf * KIN;
        double tmp[3], pf[3];
        mat3vec(Kf + 12, M.ee_off[f], tmp);
        for (int k = 0; k < 3; ++k) pf[k] = Kf[21 + k] + tmp[k];
        for (int c = jf; c > 0; c = M.parent[c]) {
            const double* Kc = K + (size_t)c * KIN;
            const int iv = M.idx_v[c];
            const double d[3] = {pf[0] - Kc[21], pf[1] - Kc[22], pf[2] - Kc[23]};
            if (M.jtype[c] == JT_FF) {
                const double* R = Kc + 12;
                const double Sd[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0};
                double SR[9];
                mat3mul(Sd, R, SR);
                for (int r = 0; r < 3; ++r) for (int e2 = 0; e2 < 3; ++e2) { Ws[(m3 + r) * BV + iv + e2] = R[3 * r + e2]; Ws[(m3 + r) * BV + iv + 3 + e2] = -SR[3 * r + e2]; }
            } else {
                double ax[3], axw[3], col[3];
                joint_axis(M, c, ax);
                mat3vec(Kc + 12, ax, axw);
                cross(axw, d, col);
                for (int r = 0; r < 3; ++r) Ws[(m3 + r) * BV + iv] = col[r];
            }
        }
        m3 += 3;
    }
    m3_out[s] = m3;
    if (m3 == 0) return;
    // S = J J^T (lower, packed), Cholesky with the pinv rank rule (pivot <= 1e-13 max diag: row dropped), W = L^-1 J in place
    double S[BMR * (BMR + 1) / 2];
    double maxd = 0.0;
    for (int a = 0; a < m3; ++a)
        for (int b = 0; b <= a; ++b) {
            double t = 0.0;
            for (int k = 0; k < nv; ++k) t += Ws[a * BV + k] * Ws[b * BV + k];
            S[tri(a, b)] = t;
            if (a == b) maxd = fmax(maxd, t);
        }
    const double tol = 1e-13 * maxd;
    int lost = 0;
    for (int a = 0; a < m3; ++a) {
        double dgn = S[tri(a, a)];
        for (int k = 0; k < a; ++k) dgn -= S[tri(a, k)] * S[tri(a, k)];
        double inv;
        if (dgn > tol) inv = 1.0 / sqrt(dgn); else { inv = 0.0; lost = 1; }
        S[tri(a, a)] = inv;                                       // the diagonal keeps 1 / L[a][a]; 0 marks a dropped row
        for (int b = a + 1; b < m3; ++b) {
            double t = S[tri(b, a)];
            for (int k = 0; k < a; ++k) t -= S[tri(b, k)] * S[tri(a, k)];
            S[tri(b, a)] = t * inv;
        }
    }
    for (int col = 0; col < nv; ++col)
        for (int a = 0; a < m3; ++a) {
            double t = Ws[a * BV + col];
            for (int k = 0; k < a; ++k) t -= S[tri(a, k)] * Ws[k * BV + col];
            Ws[a * BV + col] = t * S[tri(a, a)];
        }
    if (lost && rankloss) atomicAdd(rankloss, 1);
}

// Z[s][k][col] = sum_r W[s][k][r] Ytilde[s][r][col]: thread per (sample, column), all contact rows of the sample at once -- the column
// of Ytilde is read once instead of once per contact row (the W rows are warp-wide broadcasts); same summation order per element
__global__ void big_zrows_kernel(const __grid_constant__ BigModel M, const __grid_constant__ RowMasks masks, int ns,
                                 const double* __restrict__ Yt, const double* __restrict__ W, const int* __restrict__ m3, double* __restrict__ Z) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)ns * BCW) return;
    const int col = (int)(e % BCW), s = (int)(e / BCW);
    const int m = m3[s];
    double acc[BMR];
#pragma unroll
    for (int k = 0; k < BMR; ++k) acc[k] = 0.0;
    if (m > 0) {
        const double* w = W + (size_t)s * BMR * BV;
        const double* y = Yt + (size_t)s * M.nv * BCW + col;
        for (int r = 0; r < M.nv; ++r) {
            if (!((masks.m[r] >> (col >> 3)) & 1ull)) continue;        // structurally zero in this column's tile: an exact no-op skipped
            const double yv = y[(size_t)r * BCW];
#pragma unroll
            for (int k = 0; k < BMR; ++k) if (k < m) acc[k] = fma(w[k * BV + r], yv, acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < BMR; ++k) if (k < 3 * M.n_ee) Z[((size_t)s * BMR + k) * BCW + col] = acc[k];      // the other slots are never read
}

// ---------------------------------------------------------------------------------------------- syrk
// partial[z][blk][64][64] += sign * R^T R, blk = lower-triangular 64 x 64 output block, z = slice of the chunk's samples.
// The rows of a chunk are taken CLASS BY CLASS -- class v = row v of every sample (R is
// [ns][rps][BCW]) -- because the support of a row of Ytilde is structural: joint row v holds the ten columns of every body below its
// joint, its own two friction columns and the torque column, nothing else.  masks.m[v] has one bit per 8-column tile that row v can
// touch: a CTA skips the classes that miss one of its two 64-column blocks (G1-29dof: 322 of 735 class x block pairs remain), a warp
// skips the 8 x 8 tiles whose A- or B-side tile is structurally zero (16 % of the DMMAs remain).  Panels (32 samples of one class)
// arrive through a three-stage cp.async ring in shared memory: two are in flight while one is contracted, one barrier per panel.
// Fixed order: bit-reproducible.
constexpr int SY_PF = SY_ROWS * SY_BLK / 2 / SY_THREADS;        // 16-byte copies per thread and side of a panel (4)
constexpr int SY_STAGES = 3;                                    // panels in flight per CTA (cp.async ring in shared memory)
constexpr int SY_STAGE_DOUBLES = 2 * SY_ROWS * SY_LD;
constexpr size_t SY_ROWS_SMEM = sizeof(double) * SY_STAGES * SY_STAGE_DOUBLES;      // 104 448 B: two CTAs per SM

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

__global__ void __launch_bounds__(SY_THREADS, 2)
big_syrk_rows_kernel(const double* __restrict__ R, int ns, int rps, const __grid_constant__ RowMasks masks, double sign,
                     double* __restrict__ partial) {
    extern __shared__ __align__(16) double sy_sm[];
    __shared__ int vlist[BV];
    __shared__ int s_nact;
    const int blk = blockIdx.x, nz = gridDim.y, z = blockIdx.y;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
    const int bj = blk - bi * (bi + 1) / 2;
    const bool diag = (bi == bj);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, kk = lane & 3;
    int per = (ns + nz - 1) / nz;
    per = (per + SY_ROWS - 1) / SY_ROWS * SY_ROWS;
    const int s0 = z * per, s1 = min(ns, s0 + per);
    if (threadIdx.x == 0) {
        int n = 0;
        for (int v = 0; v < rps; ++v) {
            const unsigned long long m = masks.m[v];
            if (((m >> (8 * bi)) & 0xffull) && ((m >> (8 * bj)) & 0xffull)) vlist[n++] = v;
        }
        s_nact = n;
    }
    __syncthreads();
    const int npan = (s1 > s0) ? (s1 - s0 + SY_ROWS - 1) / SY_ROWS : 0;
    const int total = s_nact * npan;
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    // panel p = 32 samples of class vlist[p / npan] -> ring stage p % SY_STAGES; always commits a group (possibly empty) so that
    // the wait below counts the same for every thread and every p
    auto issue = [&](int p) {
        if (p < total) {
            const int v = vlist[p / npan], sb = s0 + SY_ROWS * (p % npan);
            double* PA = sy_sm + (p % SY_STAGES) * SY_STAGE_DOUBLES;
            double* PB = PA + SY_ROWS * SY_LD;
            const unsigned long long m = masks.m[v];
            const unsigned ma = (unsigned)((m >> (8 * bi)) & 0xffull), mb = (unsigned)((m >> (8 * bj)) & 0xffull);
#pragma unroll
            for (int i = 0; i < SY_PF; ++i) {
                const int idx = threadIdx.x + SY_THREADS * i, rr = idx >> 5, c2 = idx & 31;
                const int smp = sb + rr;
                // only the 8-column tiles the class can touch are moved (the contraction below never reads the others)
                const bool wa = (ma >> (c2 >> 2)) & 1u, wb = !diag && ((mb >> (c2 >> 2)) & 1u);
                if (smp < s1) {
                    const double* src = R + ((size_t)smp * rps + v) * BCW;
                    if (wa) cp_async16(PA + rr * SY_LD + 2 * c2, src + bi * SY_BLK + 2 * c2);
                    if (wb) cp_async16(PB + rr * SY_LD + 2 * c2, src + bj * SY_BLK + 2 * c2);
                } else {
                    if (wa) *reinterpret_cast<double2*>(PA + rr * SY_LD + 2 * c2) = make_double2(0.0, 0.0);
                    if (wb) *reinterpret_cast<double2*>(PB + rr * SY_LD + 2 * c2) = make_double2(0.0, 0.0);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int p = 0; p < SY_STAGES - 1; ++p) issue(p);
    for (int p = 0; p < total; ++p) {
        cp_async_wait<SY_STAGES - 2>();                    // this thread's copies of panel p have landed ...
        __syncthreads();                                   // ... and everybody's; stage (p - 1) % SY_STAGES is free again
        issue(p + SY_STAGES - 1);
        const double* PA = sy_sm + (p % SY_STAGES) * SY_STAGE_DOUBLES;
        const double* Bs = diag ? PA : PA + SY_ROWS * SY_LD;
        const unsigned long long m = masks.m[vlist[p / npan]];
        const unsigned ma = (unsigned)((m >> (8 * bi)) & 0xffull), mb = (unsigned)((m >> (8 * bj)) & 0xffull);
        if ((ma >> warp) & 1u) {                           // warp-uniform: this warp's eight output rows see the class at all
#pragma unroll
            for (int ks = 0; ks < SY_ROWS / 4; ++ks) {
                const double a = PA[(4 * ks + kk) * SY_LD + 8 * warp + r];
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if ((mb >> t) & 1u) dmma884(acc[t][0], acc[t][1], a, Bs[(4 * ks + kk) * SY_LD + 8 * t + r]);
            }
        }
    }
    cp_async_wait<0>();
    double* out = partial + ((size_t)z * SY_NBLK + blk) * SY_BLK * SY_BLK;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        double2* q = reinterpret_cast<double2*>(out + (8 * warp + r) * SY_BLK + 8 * t + 2 * kk);
        double2 v2 = *q;
        v2.x += sign * acc[t][0]; v2.y += sign * acc[t][1];      // sign = +-1: exact
        *q = v2;
    }
}

// host: structural tile masks of the rows of Ytilde (see big_rows_kernel / big_tail_kernel for who writes what)
inline RowMasks big_row_masks(const BigModel& M, int friction) {
    RowMasks k;
    for (int v = 0; v < BV; ++v) k.m[v] = 0ull;
    auto set = [&](int v, int col) { k.m[v] |= 1ull << (col / 8); };
    for (int i = 1; i < M.njoints; ++i)                        // body i: its ten columns in the rows of every joint above it
        for (int j = i; j > 0; j = M.parent[j]) {
            const int iv = M.idx_v[j], nr = (M.jtype[j] == JT_FF) ? 6 : 1;
            for (int rr = 0; rr < nr; ++rr)
                for (int c = 0; c < 10; ++c) set(iv + rr, 10 * (i - 1) + c);
        }
    const int np = M.nparams, ctau = np + (friction ? 2 * M.nd : 0);
    for (int kq = 0; kq < M.nd; ++kq) {
        if (friction) { set(6 + kq, np + kq); set(6 + kq, np + M.nd + kq); }
        set(6 + kq, ctau);
    }
    return k;
}
// Z = W Ytilde: dense rows, but only 3 n_ee of the BMR row slots of a sample are ever used (the rest stay zero)
inline RowMasks big_z_masks(const BigModel& M) {
    RowMasks k;
    for (int v = 0; v < BV; ++v) k.m[v] = (v < 3 * M.n_ee) ? (1ull << (BCW / 8)) - 1ull : 0ull;
    return k;
}

// stats = [G (c x c) | r (c) | s | n] += sum_z partial (fixed order); element (i, j), i >= j, of the augmented (c+1) x (c+1) Gram
__global__ void big_reduce_kernel(const double* __restrict__ partial, int nz, int c, double n_add, double* __restrict__ stats) {
    const int ca = c + 1, total = ca * (ca + 1) / 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) {
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        const int bi = i / SY_BLK, bj = j / SY_BLK;
        const size_t off = ((size_t)(bi * (bi + 1) / 2 + bj)) * SY_BLK * SY_BLK + (size_t)(i % SY_BLK) * SY_BLK + (j % SY_BLK);
        double sum = 0.0;
        for (int z = 0; z < nz; ++z) sum += partial[(size_t)z * SY_NBLK * SY_BLK * SY_BLK + off];
        if (i < c) {
            stats[(size_t)i * c + j] += sum;
            if (i != j) stats[(size_t)j * c + i] += sum;
        } else if (j < c) stats[(size_t)c * c + j] += sum;
        else stats[(size_t)c * c + c] += sum;
    }
    if (e == 0) stats[(size_t)c * c + c + 1] += n_add;
}

// ---------------------------------------------------------------------------------------------- per-sample outputs
// A = Ytilde - W^T Z (projected rows), b = its torque column, P = I - W^T W; Y = the body columns of Ytilde.
__global__ void big_emit_kernel(const __grid_constant__ BigModel M, int ns, int ncols, int ctau, const double* __restrict__ Yt,
                                const double* __restrict__ W, const double* __restrict__ Z, const int* __restrict__ m3,
                                double* __restrict__ Yout, double* __restrict__ A, double* __restrict__ b, double* __restrict__ P) {
    const int nv = M.nv, np = M.nparams;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int width = BCW;
    if (e >= (long long)ns * nv * width) return;
    const int col = (int)(e % width), r = (int)((e / width) % nv), s = (int)(e / ((long long)width * nv));
    const double y = Yt[e];
    if (Yout) { if (col < np) Yout[((size_t)s * nv + r) * np + col] = y; }
    if (A || b) {
        double t = y;
        const int m = m3[s];
        for (int k = 0; k < m; ++k) t -= W[((size_t)s * BMR + k) * BV + r] * Z[((size_t)s * BMR + k) * BCW + col];
        if (A && col < ncols) A[((size_t)s * nv + r) * ncols + col] = t;
        if (b && col == ctau) b[(size_t)s * nv + r] = t;
    }
    if (P && col < nv) {
        double t = (r == col) ? 1.0 : 0.0;
        const int m = m3[s];
        for (int k = 0; k < m; ++k) t -= W[((size_t)s * BMR + k) * BV + r] * W[((size_t)s * BMR + k) * BV + col];
        P[((size_t)s * nv + r) * nv + col] = t;
    }
}

// evaluation pass: e2[s][k] = ((P Ytilde x)[6 + k])^2 with x = [phi; 0; -1]  (A phi - b on the joint rows)
__global__ void big_err_kernel(const __grid_constant__ BigModel M, int ns, int ctau, const double* __restrict__ Yt, const double* __restrict__ W,
                               const double* __restrict__ Z, const int* __restrict__ m3, const double* __restrict__ phi, double* __restrict__ e2) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ns * M.nd) return;
    const int s = e / M.nd, k = e % M.nd, r = 6 + k, np = M.nparams;
    const double* y = Yt + ((size_t)s * M.nv + r) * BCW;
    double t = -y[ctau];
    for (int c = 0; c < np; ++c) t = fma(y[c], phi[c], t);
    const int m = m3[s];
    for (int kk = 0; kk < m; ++kk) {
        const double* z = Z + ((size_t)s * BMR + kk) * BCW;
        double zt = -z[ctau];
        for (int c = 0; c < np; ++c) zt = fma(z[c], phi[c], zt);
        t -= W[((size_t)s * BMR + kk) * BV + r] * zt;
    }
    e2[e] = t * t;
}

// out[k] += sum over the chunk of e2[:, k], one thread per joint, fixed order
__global__ void big_err_sum_kernel(int ns, int nd, const double* __restrict__ e2, double* __restrict__ sums) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nd) return;
    double t = sums[k];
    for (int s = 0; s < ns; ++s) t += e2[(size_t)s * nd + k];
    sums[k] = t;
}

__global__ void big_err_final_kernel(int nd, long long N, const double* __restrict__ sums, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < nd; ++k) { tot += sums[k]; out[1 + k] = sqrt(sums[k] / (double)N); }
        out[0] = tot / (double)N;
    }
}

// chunk workspace layout (doubles): kin | Ytilde | Z | W | partial | e2 / sums; ints: m3 | rankloss
constexpr int BIG_CHUNK = 8192;        // samples per pass: the thread-per-sample kernels (kin, contact) are latency-bound, a chunk must fill the SMs
constexpr int BIG_NZ = 64;            // sample slices per output block: 1 344 CTAs of uneven weight (6 .. 33 row classes) balance over 296 slots
struct BigWs {
    double* kin; double* Yt; double* Z; double* W; double* partial; double* e2; int* m3; int* rankloss;
    double* kin2; double* Yt2; double* Z2; double* W2; int* m32;       // second set of the per-chunk buffers (== the first when absent)
    size_t bytes;
};
// two = true: a second set of the per-chunk buffers (kin, Ytilde, Z, W, m3) behind the first, for the statistics pass that builds the
// rows of chunk i + 1 while the SYRK of chunk i runs on another stream; big_workspace_alt() returns the view on that set
inline BigWs big_workspace(const BigModel& M, void* base, bool two = false) {
    BigWs w;
    char* p = (char*)base;
    auto take = [&](size_t n) { char* q = p; p += (n + 255) & ~(size_t)255; return q; };
    w.kin = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)M.njoints * KIN);
    w.Yt = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)M.nv * BCW);
    w.Z = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)BMR * BCW);
    w.W = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)BMR * BV);
    w.partial = (double*)take(sizeof(double) * (size_t)BIG_NZ * SY_NBLK * SY_BLK * SY_BLK);
    w.e2 = (double*)take(sizeof(double) * ((size_t)BIG_CHUNK * BV + BV));
    w.m3 = (int*)take(sizeof(int) * BIG_CHUNK);
    w.rankloss = (int*)take(sizeof(int) * 4);
    w.kin2 = w.kin; w.Yt2 = w.Yt; w.Z2 = w.Z; w.W2 = w.W; w.m32 = w.m3;
    if (two) {
        w.kin2 = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)M.njoints * KIN);
        w.Yt2 = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)M.nv * BCW);
        w.Z2 = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)BMR * BCW);
        w.W2 = (double*)take(sizeof(double) * BIG_CHUNK * (size_t)BMR * BV);
        w.m32 = (int*)take(sizeof(int) * BIG_CHUNK);
    }
    w.bytes = (size_t)(p - (char*)base);
    return w;
}
inline BigWs big_workspace_alt(const BigWs& w) {
    BigWs a = w;
    a.kin = w.kin2; a.Yt = w.Yt2; a.Z = w.Z2; a.W = w.W2; a.m3 = w.m32;
    return a;
}

}  // namespace big
}  // namespace sysid
