// Tile fill of the fused Gram kernel on the fp64 tensor pipe.
//
// Row k of a sample's block is q_k^T Ytilde (q_k: basis vector k of null(J_c), phases.cuh).  For body i that is
//     (q_k^T Ytilde_i)[c] = d_ki . W_i[:, c],   d_ki = (dl; da) = q_k[0:6] + sum_{j in anc(i)} q_k[6 + j] a_j   (6-vector),
// W_i (6 x 10) = the base-frame wrench (F; N_O) of each of the body's ten regressor columns: bodyRegressor(omega, alpha, acc)
// carried to the base frame by the force action of X_i = (R_i, p_i).  So the nq x 10 block of body i is the product
// D_i (nq x 6) W_i (6 x 10): the earlier scalar fill evaluated it entry by entry (~85 flops and ~12 shared-memory wavefronts
// per (row, body), issue- and LSU-bound: 34 % of the kernel); here it is three DMMAs per (8 rows, body):
//
//   wbuild   thread per (sample, body, column unit): W_i once per sample, 42 non-zero doubles per body (the six inertia
//            columns carry no force: F = 0), pre-multiplied by sqrt(weight)
//   proj     warp per (sample, part of the bodies): every lane keeps ITS A-fragment element of d_k for its row k --
//            two scalars, advanced by one FMA each per joint down the chain -- multiplies with the B fragments of W_i and
//            stores the C fragments into the tile.  K order (N0, N1, N2, F0 | F1, F2, -, -): the second k-step only feeds
//            the mass / first-moment columns, so a body costs 3 DMMAs per 8 rows, not 4.
//            The friction / torque columns are the product Q[6:, k]^T [diag(dq) | diag(sign dq) | tau] and go through the
//            same pipe (K = the joints, one DMMA per non-zero 4 x 8 block).
// Extra tensor work: 6 + 1.5 DMMAs per (sample, body) on top of the 735 of the contraction (+11 %); scalar work per sample
// drops from ~2400 warp instructions to ~500.
#pragma once
#include "phases.cuh"

namespace sysid {

constexpr int WB = 44;                              // doubles per (sample, body): Wn[3][10] (moment rows), Wf[3][4] (force rows), 2 zeros
constexpr int WB_F = 30, WB_Z = 42;

// ---------------------------------------------------------------------------------------------- wbuild
// Units: 0 = mass column, 1..3 = first-moment columns, 4..6 = pairs of inertia columns.  Two warps per unit (TS * MAXB <= 64
// items), so a warp never diverges on the column formulas.
template <int TS>
__device__ __forceinline__ void phase_wbuild(const DevModel& M, const double* __restrict__ ctx, double* __restrict__ Wsm,
                                             int s0, int cnt, int warp, int lane) {
    static_assert(TS * MAXB <= 64, "two warps per column unit");
    const int unit = warp >> 1, idx = ((warp & 1) << 5) + lane;
    if (unit > 6 || idx >= TS * M.nb) return;
    const int s = idx % TS, i = idx / TS;
    if (s >= cnt) return;
    const double* c = ctx + (s0 + s) * CX_STRIDE;
    const double wsq = c[CX_W];
    const double2* b2 = reinterpret_cast<const double2*>(c + CX_B9 + B9S * i);
    const double2 q0 = b2[0], q1 = b2[1], q2 = b2[2], q3 = b2[3];
    const double w0 = q0.x, w1 = q0.y, w2 = q1.x, al0 = q1.y, al1 = q2.x, al2 = q2.y, ac0 = q3.x, ac1 = q3.y, ac2 = c[CX_B9 + B9S * i + 8];
    double R[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0}, p0 = 0.0, p1 = 0.0, p2 = 0.0;
    if (i > 0) {
        const double2* X2 = reinterpret_cast<const double2*>(c + CX_X + 12 * (i - 1));       // joint i + 1 <-> slot i - 1
        const double2 x01 = X2[0], x23 = X2[1], x45 = X2[2], x67 = X2[3], x8p = X2[4], p12 = X2[5];
        R[0] = x01.x; R[1] = x01.y; R[2] = x23.x; R[3] = x23.y; R[4] = x45.x; R[5] = x45.y; R[6] = x67.x; R[7] = x67.y; R[8] = x8p.x;
        p0 = x8p.y; p1 = p12.x; p2 = p12.y;
    }
    double* W = Wsm + (s * MAXB + i) * WB;
    if (unit == 0) *reinterpret_cast<double2*>(W + WB_Z) = make_double2(0.0, 0.0);        // the zero slot "no contribution" loads read
    if (unit < 4) {
        double f0, f1, f2, n0, n1, n2;
        if (unit == 0) { f0 = ac0; f1 = ac1; f2 = ac2; n0 = n1 = n2 = 0.0; }
        else {
            // column of [alpha]x + [omega]x [omega]x and of -[acc]x for the unit vector e
            const double e0 = (unit == 1) ? 1.0 : 0.0, e1 = (unit == 2) ? 1.0 : 0.0, e2 = (unit == 3) ? 1.0 : 0.0;
            const double u0 = w1 * e2 - w2 * e1, u1 = w2 * e0 - w0 * e2, u2 = w0 * e1 - w1 * e0;      // omega x e
            f0 = (al1 * e2 - al2 * e1) + (w1 * u2 - w2 * u1);
            f1 = (al2 * e0 - al0 * e2) + (w2 * u0 - w0 * u2);
            f2 = (al0 * e1 - al1 * e0) + (w0 * u1 - w1 * u0);
            n0 = e1 * ac2 - e2 * ac1; n1 = e2 * ac0 - e0 * ac2; n2 = e0 * ac1 - e1 * ac0;            // e x acc
        }
        const double F0 = R[0] * f0 + R[1] * f1 + R[2] * f2, F1 = R[3] * f0 + R[4] * f1 + R[5] * f2, F2 = R[6] * f0 + R[7] * f1 + R[8] * f2;
        const double N0 = R[0] * n0 + R[1] * n1 + R[2] * n2 + (p1 * F2 - p2 * F1);
        const double N1 = R[3] * n0 + R[4] * n1 + R[5] * n2 + (p2 * F0 - p0 * F2);
        const double N2 = R[6] * n0 + R[7] * n1 + R[8] * n2 + (p0 * F1 - p1 * F0);
        W[unit] = N0 * wsq; W[10 + unit] = N1 * wsq; W[20 + unit] = N2 * wsq;
        W[WB_F + unit] = F0 * wsq; W[WB_F + 4 + unit] = F1 * wsq; W[WB_F + 8 + unit] = F2 * wsq;
    } else {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int k = 2 * (unit - 4) + v;          // 0..5 <-> Ixx, Ixy, Iyy, Ixz, Iyz, Izz: S_k = e_a e_b^T + e_b e_a^T (a != b) or e_a e_a^T
            const int a = (k == 2 || k == 4) ? 1 : ((k == 5) ? 2 : 0), b = (k == 0) ? 0 : ((k == 1 || k == 2) ? 1 : 2);
            const double ala = (a == 0) ? al0 : ((a == 1) ? al1 : al2), alb = (b == 0) ? al0 : ((b == 1) ? al1 : al2);
            const double wa = (a == 0) ? w0 : ((a == 1) ? w1 : w2), wb = (b == 0) ? w0 : ((b == 1) ? w1 : w2);
            // S_k u = e_a u_b + e_b u_a (a != b), e_a u_a (a == b)
            double sa[3] = {0.0, 0.0, 0.0}, sw[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                if (x == a) { sa[x] += alb; sw[x] += wb; }
                if (x == b && a != b) { sa[x] += ala; sw[x] += wa; }
            }
            const double n0 = sa[0] + (w1 * sw[2] - w2 * sw[1]), n1 = sa[1] + (w2 * sw[0] - w0 * sw[2]), n2 = sa[2] + (w0 * sw[1] - w1 * sw[0]);
            W[4 + k] = (R[0] * n0 + R[1] * n1 + R[2] * n2) * wsq;
            W[14 + k] = (R[3] * n0 + R[4] * n1 + R[5] * n2) * wsq;
            W[24 + k] = (R[6] * n0 + R[7] * n1 + R[8] * n2) * wsq;
        }
    }
}

// ---------------------------------------------------------------------------------------------- proj
// One pass over the part's bodies for NM (1 or 2) groups of 8 rows of one sample.  Written for instruction count: every
// per-lane index is fixed before the loops, "no contribution" is a load from a zero slot (W[42..43] of every body, written by
// wbuild) instead of a predicate, and a sample without a stance foot (Q = I, not stored) reads the rows of an identity table.
//   Qrow[m]   this lane's basis vector (row 8 (mt0 + m) + lane / 4 of the sample's block; clamped for rows >= nq, never stored)
//   trow[m]   tile row of that basis vector, + 2 (lane % 4)
template <int NM, int LD>
__device__ __forceinline__ void proj_pass(const DevModel& M, const double* __restrict__ c, const double* __restrict__ Wb,
                                          const double* const (&Qrow)[NM], double* const (&trow)[NM], const bool (&ok)[NM],
                                          int part, int friction, int lane, unsigned itemword) {
    const unsigned full = 0xffffffffu;
    const int r = lane >> 2, kk = lane & 3;
    const int offA = (kk < 3) ? 3 + kk : 0;                 // d component that meets K rows (N0, N1, N2, F0) ...
    const int offB = (kk < 2) ? 1 + kk : 0;                 // ... and (F1, F2, -, -): lanes kk >= 2 meet zero rows of W, any finite value does
    const int o00 = (kk < 3) ? kk * 10 + r : ((r < 4) ? WB_F + r : WB_Z);
    const int o10 = (kk < 2 && r < 4) ? WB_F + 4 * (1 + kk) + r : WB_Z;
    const int o01 = (kk < 3 && r < 2) ? kk * 10 + 8 + r : WB_Z;
    const double* pA = c + CX_A - 12 + offA;                // a_j[offA] = pA[6 j]
    const double* pB = c + CX_A - 12 + offB;
    double d0A[NM], d0B[NM], dA[NM], dB[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) { d0A[m] = Qrow[m][offA]; d0B[m] = Qrow[m][offB]; dA[m] = d0A[m]; dB[m] = d0B[m]; }
    auto emit = [&](int i) {
        const double* W = Wb + i * WB;
        const double b00 = W[o00], b10 = W[o10], b01 = W[o01];
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
            dmma884(c0, c1, dA[m], b00);
            dmma884(e0, e1, dA[m], b01);
            dmma884(c0, c1, dB[m], b10);
            if (ok[m]) {
                *reinterpret_cast<double2*>(trow[m] + 10 * i) = make_double2(c0, c1);
                if (kk == 0) *reinterpret_cast<double2*>(trow[m] + 10 * i + 8) = make_double2(e0, e1);
            }
        }
    };
    if (M.proj_root[part]) emit(0);
    const int nitems = M.proj_n[part];
    for (int it = 0; it < nitems; ++it) {
        const unsigned w = __shfl_sync(full, itemword, it);
        const int j = (int)(w & 0xffu);
        if (w & 0x200u) {
#pragma unroll
            for (int m = 0; m < NM; ++m) { dA[m] = d0A[m]; dB[m] = d0B[m]; }
        }
        const double aA = pA[6 * j], aB = pB[6 * j];
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const double qj = Qrow[m][4 + j];
            dA[m] = fma(qj, aA, dA[m]);
            dB[m] = fma(qj, aB, dB[m]);
        }
        if (w & 0x100u) emit(j - 1);
    }
    // friction / torque / padding columns: Q[6:, k]^T [diag(dq) | diag(sign dq) | tau], 8 columns per group from column nparams
    unsigned mask = M.proj_tail[part];
    if (mask) {
        const int np = M.nparams, nd = M.nd;
        const double wsq = c[CX_W];
        const int tcol = friction ? 2 * nd : 0;             // the torque column (relative to nparams)
        double vq[3], vs[3], vt[3], qa[NM][3];
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            const int j = 4 * ks + kk;
            const bool valid = j < nd;
            const double dv = valid ? c[CX_DQ + j] : 0.0;
            vq[ks] = dv * wsq;
            vs[ks] = ((dv > 0.0) ? 1.0 : ((dv < 0.0) ? -1.0 : (dv == 0.0 ? 0.0 : dv))) * wsq;       // numpy sign: sign(nan) = nan
            vt[ks] = valid ? c[CX_TAU + j] * wsq : 0.0;
#pragma unroll
            for (int m = 0; m < NM; ++m) qa[m][ks] = valid ? Qrow[m][6 + j] : 0.0;
        }
        while (mask) {
            const int nt = __ffs(mask) - 1;
            mask &= mask - 1;
            const unsigned ksm = M.proj_tailks[friction ? 1 : 0][nt];
            const int col = 8 * nt + r;                     // B operand: this lane's column of the group
            double c0[NM], c1[NM];
#pragma unroll
            for (int m = 0; m < NM; ++m) { c0[m] = 0.0; c1[m] = 0.0; }
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                if (!((ksm >> ks) & 1u)) continue;
                const int j = 4 * ks + kk;
                double b = (col == tcol) ? vt[ks] : 0.0;
                if (friction) b = (col == j) ? vq[ks] : ((col == nd + j) ? vs[ks] : b);
#pragma unroll
                for (int m = 0; m < NM; ++m) dmma884(c0[m], c1[m], qa[m][ks], b);
            }
            if (np + 8 * nt + 2 * kk + 1 < LD) {
#pragma unroll
                for (int m = 0; m < NM; ++m) if (ok[m]) *reinterpret_cast<double2*>(trow[m] + np + 8 * nt) = make_double2(c0[m], c1[m]);
            }
        }
    }
}

// off[u] = first tile row of the round's sample u (off[cnt] = rows of the round).  Warp -> (sample = warp % 4, part = warp / 4).
// ident18: an 18 x 18 identity in shared memory (pitch MAXV).
template <int TS, int LD>
__device__ __forceinline__ void phase_proj_mma(const DevModel& M, const double* __restrict__ ctx, const double* __restrict__ Wsm,
                                               const double* __restrict__ ident18, double* __restrict__ tile, int s0, int cnt,
                                               const int (&off)[TS + 1], int friction, int warp, int lane) {
    static_assert(TS == 4, "warp -> (sample, part) mapping");
    const int s = warp & 3, part = warp >> 2;
    if (s >= cnt) return;
    int row0 = 0, nq = 0;
#pragma unroll
    for (int u = 0; u < TS; ++u) if (u == s) { row0 = off[u]; nq = off[u + 1] - off[u]; }
    if (nq == 0) return;
    const double* c = ctx + (s0 + s) * CX_STRIDE;
    const double* Wb = Wsm + s * MAXB * WB;
    const unsigned itemword = (lane < PROJ_MAXITEMS) ? M.proj_item[part][lane] : 0u;
    const int r = lane >> 2, kk = lane & 3;
    // rows of the basis: Q (pitch QLD) for a stance sample, the identity table for a sample in flight (Q = I is not stored)
    const double* Qb = (nq == MAXV) ? ident18 : c + CX_Q;
    const int qld = (nq == MAXV) ? MAXV : QLD;
    double* tb = tile + row0 * LD + 2 * kk;
    for (int mt = 0; 8 * mt < nq; mt += 2) {
        const int ra = 8 * mt + r, rb = ra + 8;
        if (8 * (mt + 1) < nq) {
            const double* const Qrow[2] = {Qb + (ra < nq ? ra : 0) * qld, Qb + (rb < nq ? rb : 0) * qld};
            double* const trow[2] = {tb + ra * LD, tb + rb * LD};
            const bool ok[2] = {ra < nq, rb < nq};
            proj_pass<2, LD>(M, c, Wb, Qrow, trow, ok, part, friction, lane, itemword);
        } else {
            const double* const Qrow[1] = {Qb + (ra < nq ? ra : 0) * qld};
            double* const trow[1] = {tb + ra * LD};
            const bool ok[1] = {ra < nq};
            proj_pass<1, LD>(M, c, Wb, Qrow, trow, ok, part, friction, lane, itemword);
        }
    }
}

}  // namespace sysid
