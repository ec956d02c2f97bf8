// Stage 1+2, structured form: the fused regressor -> projector -> Gram kernel for legged trees (every child of the root heads a
// simple chain with at most one contact frame: Solo-12, Spot, G1-12dof and the fixed-base emulation).
//
// What it restates (reference src/sys_identification.py:119-135, 401-418 and the stacking of demo/solo_identification.py:79-84):
// G = sum_i Ytilde_i^T P_i Ytilde_i with P_i = I - pinv(J_c) J_c = Q Q^T for ANY orthonormal basis Q of null(J_c).  The basis is
// chosen so that most of its vectors are SPARSE:
//   * a swing leg contributes the unit vectors of its joints;
//   * a stance leg with more than three joints contributes the null space of its own 3 x len Jacobian block (joint motions
//     that leave the foot where it is): Householder QR of J_leg^T gives Q_leg = [range (3) | null (len - 3)];
//   * what remains is dense: with u_f the range coordinates of stance leg f, the reduced Jacobian [J_base | T_f] (T_f = J_leg
//     Q_leg,range, 3 x 3 triangular) has a null space of dimension 6 (more only if contact rows are dependent), found by the same
//     Householder QR + pinv rank rule the unstructured kernel applies to J_c itself (phase_qbuild) and expanded through Q_leg,range.
// Rows of a sparse vector touch only the bodies, friction columns of its own leg and the torque column, so with the tile columns
// permuted to [class A legs | class B legs | torque, root body] they update 55 of the 210 Gram tiles.  Per sample: 6 dense rows +
// (nq - 6) sparse rows instead of nq dense rows -- on the G1 log 424 DMMAs instead of 735, and the tile fill walks one leg instead
// of the whole tree for the sparse rows.
//
// Layout of a round (ST_TS samples): dense rows in a 160-column tile, the sparse rows of each class in an 80-column tile (72 class
// columns + the block of the torque column); row descriptors are built once per super-batch by one warp (prefix sums by shuffles).
// The fill is task-parallel, one warp per task, lane = tile row: (dense rows, run of two bodies of a leg), (dense rows, root body +
// torque column), (class rows, run of two bodies).  DESIGN.md section 4.1 has the measurements behind each of these choices.
#pragma once
#include "gram_kernels.cuh"

namespace sysid {

#include "gram_tiles_struct.inc"
static_assert(STILES_MAX_NT == 14, "tensor-memory parking moves 56 registers per thread");

// k-steps of the M loops are NOT unrolled: the sixteen per-warp code variants are re-fetched every round (measured on the 1 M-sample
// G1 log: unroll 1 / 2 / 4 = 64.6 / 61.6 / 50.9 Msamples/s)
#ifndef SYSID_ST_MMA_UNROLL
#define SYSID_ST_MMA_UNROLL 1
#endif
constexpr int ST_MMA_UNROLL = SYSID_ST_MMA_UNROLL;
#ifndef SYSID_ST_SB
#define SYSID_ST_SB 25
#endif
#ifndef SYSID_ST_TS
#define SYSID_ST_TS 5
#endif
constexpr int ST_TS = SYSID_ST_TS;                // samples per round
constexpr int ST_SB = SYSID_ST_SB;                // samples per super-batch (F phases)
constexpr int ST_NR = ST_SB / ST_TS;              // rounds per super-batch
constexpr int ST_DROWS = 32;                      // dense rows per M pass (6 per sample unless contact rows are dependent)
constexpr int ST_SROWS = (6 * ST_TS + 3) & ~3;    // sparse rows per class and round (whole k-steps)
constexpr int ST_SLD = 84;                        // pitch of a sparse row: 72 class columns + 8 (torque block); == 4 (mod 16)
constexpr int ST_DMAX = ST_TS * NQMAX;            // dense row descriptors per round
constexpr int ST_TILE_D = ST_DROWS * TILE_LD;
constexpr int ST_TILE_S = ST_SROWS * ST_SLD;
constexpr int ST_FSCR = ST_SB * (SC_STRIDE + IN_CHANNELS);
constexpr int ST_FRONT = (ST_TILE_D + 2 * ST_TILE_S > ST_FSCR) ? ST_TILE_D + 2 * ST_TILE_S : ST_FSCR;
constexpr size_t ST_SMEM_BYTES = sizeof(double) * (ST_FRONT + ST_SB * CX_STRIDE);
constexpr int ST_TAUCOL = 144, ST_ROOTCOL = 146;
static_assert((GRAM_WARPS - 2) * 32 >= ST_SB * NQMAX, "qcols runs on warps 1 .. 14");
static_assert(ST_SROWS <= 32 && ST_DROWS <= 32, "one lane per tile row in the fill");
static_assert(ST_SB % ST_TS == 0 && ST_SB <= 32 && ST_SLD % 16 == 4 && ST_SMEM_BYTES + 6144 <= 232448, "layout");
static_assert(IN_DDQ - IN_DQ + MAXV >= 3 * MAXD, "the dead dq / ddq channels hold the range coefficients of the legs");

__device__ __forceinline__ int st_rec_leg(uint32_t r) { return r & 3; }
__device__ __forceinline__ int st_rec_pos(uint32_t r) { return (r >> 2) & 7; }
__device__ __forceinline__ int st_rec_cls(uint32_t r) { return (r >> 5) & 1; }
__device__ __forceinline__ int st_rec_slot(uint32_t r) { return (r >> 6) & 7; }
__device__ __forceinline__ int st_rec_bcol(uint32_t r) { return (r >> 9) & 255; }
__device__ __forceinline__ int st_rec_fvcol(uint32_t r) { return (r >> 17) & 255; }

// row descriptor: sample (5) | vector (5) << 5; vector 0 .. 14 = row of the sample's CX_Q, 16 + u = unit vector of joint u + 2
// (a row of the CTA's identity table, s_unit)
__device__ __forceinline__ uint32_t st_desc(int s, int qsel) { return (uint32_t)s | ((uint32_t)qsel << 5); }
__device__ __forceinline__ const double* st_vector(uint32_t desc, const double* __restrict__ c, const double* __restrict__ unit) {
    const int qsel = (desc >> 5) & 31;
    return (qsel & 16) ? unit + (qsel & 15) * QLD : c + CX_Q + qsel * QLD;
}

// ---------------------------------------------------------------------------------------------- feet + leg QR (one phase)
// Items [0, SB n_ee): (sample, stance slot) -- the slot's contact frame, its world-aligned lever arm and the stance bookkeeping
// (what phase_feet's e == MAXCH items do).  Items after that: (sample, leg, column i of Q_leg).  A column thread reads the
// contact flags itself, forms the leg columns of its foot's contact-Jacobian rows (phase_feet's other items: here every column
// thread forms all of them, 45 flops each, instead of waiting one barrier interval for them), then: stance leg: Householder QR of
// J_leg^T (len x 3, len = joints between the foot and the root, 3..6), recomputed by each of the leg's column threads (150
// flops; the alternative is one thread per leg running all six columns one after the other while fifteen warps wait).  Column
// thread i < 3 writes the range coefficients [joint][i] into dead input channels, i >= 3 the null vector i - 3 straight into the
// basis (rows sbase ..); thread 0 also writes T = J_leg Q_range into the leg part of the slot's three rows of the reduced
// Jacobian (scratch rows, where qbuild_red expects them) and the joints' reduced offset (-1: swing leg / joint beyond the foot).
template <int SB>
__device__ __forceinline__ void phase_feet_legqr(const DevModel& M, long long base, long long N, double* __restrict__ ctx,
                                                 double* __restrict__ scr, double* __restrict__ inp, int t) {
    const int nee = M.n_ee, nslot_items = SB * nee;
    if (nee == 0 && t < SB) scr[t * SC_STRIDE + SC_META] = 0.0;      // no contact frame: zero contact rows
    if (t < nslot_items) {
        const int s = t % SB, slot = t / SB;
        if (base + s >= N) return;
        double* sc = scr + s * SC_STRIDE;
        int m = 0, kf = -1;
#pragma unroll
        for (int k = 0; k < MAXEE; ++k)
            if (k < nee && inp[(IN_CNT + k) * SB + s] != 0.0) { if (m == slot) kf = k; ++m; }    // truthiness rule of the reference (state 2, NaN: stance)
        if (slot == 0) sc[SC_META] = (double)(3 * m);
        sc[SC_META + 1 + slot] = (double)kf;
        if (kf < 0) return;
        const int jf = M.ee_joint[kf];
        double rf[3];
        if (jf == 1) { rf[0] = M.ee_off[kf][0]; rf[1] = M.ee_off[kf][1]; rf[2] = M.ee_off[kf][2]; }
        else {
            const double* X = ctx + s * CX_STRIDE + CX_X + 12 * (jf - 2);
#pragma unroll
            for (int k = 0; k < 3; ++k) rf[k] = X[9 + k] + X[3 * k] * M.ee_off[kf][0] + X[3 * k + 1] * M.ee_off[kf][1] + X[3 * k + 2] * M.ee_off[kf][2];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) sc[SC_RF + 3 * slot + k] = sc[SC_RB + 3 * k] * rf[0] + sc[SC_RB + 3 * k + 1] * rf[1] + sc[SC_RB + 3 * k + 2] * rf[2];
        return;
    }
    const int tt = t - nslot_items, nleg = M.nfch;
    if (tt >= SB * nleg * MAXCH) return;
    const int s = tt % SB, u = tt / SB, leg = u % nleg, i = u / nleg;
    if (base + s >= N) return;
    double* c = ctx + s * CX_STRIDE;
    double* sc = scr + s * SC_STRIDE;
    const int kfoot = M.st_cfoot[leg];
    // stance slots in contact-frame order: this leg's slot, the reduced-coordinate offset and the stored-vector base before it
    int m = 0, slot = -1, off = 0, sbase = 0;
#pragma unroll
    for (int k = 0; k < MAXEE; ++k) {
        if (k < nee && inp[(IN_CNT + k) * SB + s] != 0.0) {
            const int ln = M.chain_len[k];
            if (k == kfoot) slot = m;
            else if (kfoot >= 0 && k < kfoot && ln > 0) { off += 3; sbase += ln - 3; }
            if (leg == 0 && i == 0 && ln == 0) {
                // a contact frame on the root body (fixed-base emulation): its rows have no leg part
#pragma unroll
                for (int x = 0; x < 3; ++x)
                    for (int cc = 6; cc < MAXV; ++cc) sc[SC_WM + (3 * m + x) * MAXV + cc] = 0.0;
            }
            ++m;
        }
    }
    const int flen = M.fch_len[leg], j0 = M.fch[leg][0];           // the leg's joints are j0 .. j0 + flen - 1 (root -> leaf)
    if (slot < 0) {
        if (i == 0) for (int e = 0; e < flen; ++e) inp[(IN_TAU + (j0 + e - 2)) * SB + s] = -1.0;
        return;
    }
    const int len = M.chain_len[kfoot];                              // 3 .. 6; chain[kfoot][e] = j0 + len - 1 - e
    if (i >= len) return;
    if (i == 0) {
        for (int e = len; e < flen; ++e) inp[(IN_TAU + (j0 + e - 2)) * SB + s] = -1.0;
        for (int e = 0; e < len; ++e) inp[(IN_TAU + (j0 + e - 2)) * SB + s] = (double)off;
    }
    // leg columns of the foot's three contact-Jacobian rows: (R_b z_c) x (R_b (r_f - p_c)), c = joint e of the foot chain
    double a[MAXCH][3], v[3][MAXCH];
    {
        double Rb[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rb[k] = sc[SC_RB + k];
        const int jf = M.ee_joint[kfoot];
        const double* Xf = c + CX_X + 12 * (jf - 2);
        double rf[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) rf[k] = Xf[9 + k] + Xf[3 * k] * M.ee_off[kfoot][0] + Xf[3 * k + 1] * M.ee_off[kfoot][1] + Xf[3 * k + 2] * M.ee_off[kfoot][2];
#pragma unroll
        for (int e = 0; e < MAXCH; ++e) {
            if (e < len) {
                const int cj = j0 + len - 1 - e;
                const double* A = c + CX_A + 6 * (cj - 2);
                const double* X = c + CX_X + 12 * (cj - 2);
                const double ax0 = A[3], ax1 = A[4], ax2 = A[5];
                const double bx = rf[0] - X[9], by = rf[1] - X[10], bz = rf[2] - X[11];
                const double a0 = Rb[0] * ax0 + Rb[1] * ax1 + Rb[2] * ax2, a1 = Rb[3] * ax0 + Rb[4] * ax1 + Rb[5] * ax2, a2 = Rb[6] * ax0 + Rb[7] * ax1 + Rb[8] * ax2;
                const double dx = Rb[0] * bx + Rb[1] * by + Rb[2] * bz, dy = Rb[3] * bx + Rb[4] * by + Rb[5] * bz, dz = Rb[6] * bx + Rb[7] * by + Rb[8] * bz;
                a[e][0] = a1 * dz - a2 * dy; a[e][1] = a2 * dx - a0 * dz; a[e][2] = a0 * dy - a1 * dx;
            } else { a[e][0] = 0.0; a[e][1] = 0.0; a[e][2] = 0.0; }
        }
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        double tail2 = 0.0;
#pragma unroll
        for (int e = p; e < MAXCH; ++e) tail2 = fma(a[e][p], a[e][p], tail2);
        double alpha = 0.0, inv = 0.0;
        if (tail2 > 0.0) {
            const double nt = tail2 * rsqrt(tail2);
            alpha = (a[p][p] >= 0.0) ? nt : -nt;
            inv = rsqrt(2.0 * (tail2 + fabs(a[p][p]) * nt));
        }
#pragma unroll
        for (int e = 0; e < MAXCH; ++e) v[p][e] = (e < p) ? 0.0 : ((e == p) ? (a[e][p] + alpha) * inv : a[e][p] * inv);
#pragma unroll
        for (int b = p; b < 3; ++b) {
            double d = 0.0;
#pragma unroll
            for (int e = p; e < MAXCH; ++e) d = fma(v[p][e], a[e][b], d);
            d *= -2.0;
#pragma unroll
            for (int e = p; e < MAXCH; ++e) a[e][b] = fma(d, v[p][e], a[e][b]);
        }
    }
    if (i == 0) {
        // T[x][y] = R[y][x], y <= x: the leg part of row (slot, x) of the reduced Jacobian
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double* row = sc + SC_WM + (3 * slot + x) * MAXV;
            for (int cc = 6; cc < MAXV; ++cc) row[cc] = 0.0;
#pragma unroll
            for (int y = 0; y < 3; ++y) if (y <= x) row[6 + off + y] = a[y][x];
        }
    }
    // column i of Q_leg = H_0 H_1 H_2
    double x[MAXCH];
#pragma unroll
    for (int e = 0; e < MAXCH; ++e) x[e] = (e == i) ? 1.0 : 0.0;
#pragma unroll
    for (int p = 2; p >= 0; --p) {
        double d = 0.0;
#pragma unroll
        for (int e = p; e < MAXCH; ++e) d = fma(v[p][e], x[e], d);
        d *= -2.0;
#pragma unroll
        for (int e = p; e < MAXCH; ++e) x[e] = fma(d, v[p][e], x[e]);
    }
    if (i < 3) {
#pragma unroll
        for (int e = 0; e < MAXCH; ++e) if (e < len) inp[(IN_DQ + 3 * (j0 + len - 1 - e - 2) + i) * SB + s] = x[e];
    } else {
        double* qrow = c + CX_Q + (sbase + i - 3) * QLD;
        for (int cc = 0; cc < MAXV; ++cc) qrow[cc] = 0.0;
#pragma unroll
        for (int e = 0; e < MAXCH; ++e) if (e < len) qrow[4 + j0 + len - 1 - e] = x[e];
    }
}

// ---------------------------------------------------------------------------------------------- reduced QR
// phase_qbuild on the reduced Jacobian [J_base | T_f ...] (same sixteen-lanes-per-sample Householder QR, same rank rule: the row
// norms and elimination remainders are those of J_c, the leg parts having only been rotated).  Lane b assembles the base part of
// its row; the leg part was written by phase_feet_legqr.  Leaves rank in SC_META, the number of dense vectors in CX_NQ and the number
// of stored sparse vectors in CX_NQ + 1.
// The pivot loop is ROLLED: after pivot p every lane shifts its row one coordinate to the left, so the pivot coordinate is
// always x[0] and the register array keeps static indices (unrolled over the pivots this phase was 30 KB of code that ran once
// per super-batch, i.e. always from a cold instruction cache).
// LPS lanes per sample (lane b holds contact row b): 16 in general, 8 for a model with at most two contact frames (four samples
// per warp: the phase is issue-bound, every lane of a warp pays for the rows of the others).
template <int SB, int NRED, int LPS>
__device__ __forceinline__ void phase_qbuild_red(const DevModel& M, long long base, long long N, double* __restrict__ ctx,
                                                 double* __restrict__ scr, int* s_bad, int t) {
    constexpr int MR = (3 * MAXEE < LPS) ? 3 * MAXEE : LPS;
    const unsigned full = 0xffffffffu;
    const int s = t / LPS, b = t % LPS, hbase = threadIdx.x & 31 & ~(LPS - 1);
    const bool live = (s < SB) && (base + s < N);
    double* sc = scr + (live ? s : 0) * SC_STRIDE;
    const int m3 = live ? (int)sc[SC_META] : 0;
    double x[NRED];                                         // columns NRED.. of the reduced Jacobian are zero for this model
#pragma unroll
    for (int r = 0; r < NRED; ++r) x[r] = 0.0;
    double n2 = 0.0;
    if (b < m3) {
        const int slot = b / 3, xx = b - 3 * slot;
        double* mine = sc + SC_WM + b * MAXV;
        const double r0 = sc[SC_RF + 3 * slot], r1 = sc[SC_RF + 3 * slot + 1], r2 = sc[SC_RF + 3 * slot + 2];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const double b0 = sc[SC_RB + cc], b1 = sc[SC_RB + 3 + cc], b2 = sc[SC_RB + 6 + cc];
            mine[cc] = (xx == 0) ? b0 : ((xx == 1) ? b1 : b2);
            mine[3 + cc] = -((xx == 0) ? (r1 * b2 - r2 * b1) : ((xx == 1) ? (r2 * b0 - r0 * b2) : (r0 * b1 - r1 * b0)));
        }
#pragma unroll
        for (int r = 0; r < NRED; ++r) { x[r] = mine[r]; n2 = fma(x[r], x[r], n2); }
    }
    double mx = n2;
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(full, mx, o));
    const double tol = 1e-13 * mx;
    __syncwarp();                                            // every lane holds its row: the slots may now take reflectors
    int cur = 0, rank = 0, dropped = 0;                      // uniform per sample (LPS lanes)
#pragma unroll 1
    for (int p = 0; p < (MR < NRED ? MR : NRED); ++p) {
        if (!__any_sync(full, cur < m3)) break;
        // x[r] holds coordinate p + r of the lane's row (zeros past NRED - p).  Next row whose remainder is not negligible:
        bool found = false;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int r = 0; r < NRED; ++r) { if (r % 3 == 0) t0 = fma(x[r], x[r], t0); else if (r % 3 == 1) t1 = fma(x[r], x[r], t1); else t2 = fma(x[r], x[r], t2); }
        const double mytail2 = t0 + t1 + t2;
        while (__any_sync(full, !found && cur < m3)) {
            const double tail2 = __shfl_sync(full, mytail2, hbase + min(cur, LPS - 1));
            if (!found && cur < m3) {
                if (tail2 > tol) found = true;
                else { ++cur; dropped = 1; }                // dependent row: dropped (pinv semantics)
            }
        }
        if (__any_sync(full, found)) {
            double* vrow = sc + SC_WM + p * MAXV;
            if (found && b == cur) {
                const double nt = mytail2 * rsqrt(mytail2);
                const double alpha = (x[0] >= 0.0) ? nt : -nt;
                const double inv = rsqrt(2.0 * (mytail2 + fabs(x[0]) * nt));       // 1 / |x + alpha e_0|
                for (int r = 0; r < p; ++r) vrow[r] = 0.0;
#pragma unroll
                for (int r = 0; r < NRED; ++r) if (p + r < NRED) vrow[p + r] = (r == 0) ? (x[0] + alpha) * inv : x[r] * inv;
            }
            __syncwarp();
            if (found && b > cur && b < m3) {
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int r = 0; r < NRED; ++r) { const double q = (p + r < NRED) ? vrow[p + r] : 0.0; if (r & 1) d1 = fma(q, x[r], d1); else d0 = fma(q, x[r], d0); }
                const double d = -2.0 * (d0 + d1);
#pragma unroll
                for (int r = 0; r < NRED; ++r) { const double q = (p + r < NRED) ? vrow[p + r] : 0.0; x[r] = fma(d, q, x[r]); }
            }
            if (found) { ++cur; ++rank; }
        }
        // coordinate p is done for every remaining row
#pragma unroll
        for (int r = 0; r + 1 < NRED; ++r) x[r] = x[r + 1];
        x[NRED - 1] = 0.0;
    }
    if (b == 0 && s < SB) {
        int nd = 0, nsv = 0;
        if (live) {
            int nred = 6;
            const int m = m3 / 3;
            for (int sl = 0; sl < MAXEE; ++sl)
                if (sl < m) {
                    const int ln = M.chain_len[(int)sc[SC_META + 1 + sl]];
                    if (ln > 0) { nred += 3; nsv += ln - 3; }
                }
            nd = nred - rank;
            int flags = dropped ? 1 : 0;
            if (nd + nsv > NQMAX) { nd = 0; nsv = 0; flags |= 2; }      // cannot happen while R_b has rank 3 (a garbage quaternion): skip the sample
            sc[SC_META] = (double)rank;
            if (flags) atomicOr(&s_bad[s], flags);
        }
        ctx[s * CX_STRIDE + CX_NQ] = (double)nd;
        ctx[s * CX_STRIDE + CX_NQ + 1] = (double)nsv;
    }
}

// ---------------------------------------------------------------------------------------------- dense vectors
// Thread per (sample, dense vector k): y = H_0 ... H_{rank-1} e_{rank+k} in reduced coordinates, then q = [y_base; Q_range y_f].
template <int SB, int NRED>
__device__ __forceinline__ void phase_qcols_struct(const DevModel& M, long long base, long long N, double* __restrict__ ctx,
                                                   const double* __restrict__ scr, const double* __restrict__ inp, int t) {
    if (t < 0 || t >= SB * NQMAX) return;
    const int s = t / NQMAX, k = t - s * NQMAX;
    if (base + s >= N) return;
    const double* sc = scr + s * SC_STRIDE;
    double* c = ctx + s * CX_STRIDE;
    const int rank = (int)sc[SC_META], nd = (int)c[CX_NQ], nsv = (int)c[CX_NQ + 1];
    if (k >= nd) return;
    double x[NRED];
#pragma unroll
    for (int r = 0; r < NRED; ++r) x[r] = (r == rank + k) ? 1.0 : 0.0;
    for (int p = rank - 1; p >= 0; --p) {
        const double2* v2 = reinterpret_cast<const double2*>(sc + SC_WM + p * MAXV);
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int r2 = 0; r2 < NRED / 2; ++r2) { const double2 q = v2[r2]; d0 = fma(q.x, x[2 * r2], d0); d1 = fma(q.y, x[2 * r2 + 1], d1); }
        const double d = -2.0 * (d0 + d1);
#pragma unroll
        for (int r2 = 0; r2 < NRED / 2; ++r2) { const double2 q = v2[r2]; x[2 * r2] = fma(d, q.x, x[2 * r2]); x[2 * r2 + 1] = fma(d, q.y, x[2 * r2 + 1]); }
    }
    double* qk = c + CX_Q + (nsv + k) * QLD;
#pragma unroll
    for (int r = 6; r < NRED; ++r) qk[r] = x[r];                 // reduced leg coordinates, read back below at run-time offsets
    double qq[MAXD];
#pragma unroll
    for (int jd = 0; jd < MAXD; ++jd) {
        double v = 0.0;
        if (jd < M.nd) {
            const double ro = inp[(IN_TAU + jd) * SB + s];
            if (ro >= 0.0) {
                const int r = 6 + (int)ro;
                v = inp[(IN_DQ + 3 * jd) * SB + s] * qk[r] + inp[(IN_DQ + 3 * jd + 1) * SB + s] * qk[r + 1] + inp[(IN_DQ + 3 * jd + 2) * SB + s] * qk[r + 2];
            }
        }
        qq[jd] = v;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) qk[r] = x[r];
#pragma unroll
    for (int jd = 0; jd < MAXD; ++jd) qk[6 + jd] = qq[jd];
}

// ---------------------------------------------------------------------------------------------- row descriptors
// One warp, lane = sample of the super-batch; runs beside phase_finish / phase_qcols_struct, so it applies the skip rule of
// phase_finish itself (non-finite inputs or weight, weight <= 0: no rows).
template <int SB>
__device__ __forceinline__ void phase_rowdesc(const DevModel& M, long long base, long long N, const double* __restrict__ ctx,
                                              const double* __restrict__ scr, const double* __restrict__ inp, const int* s_bad, int lane,
                                              uint32_t (*descD)[ST_DMAX], uint32_t (*descS)[ST_NR][ST_SROWS], int (*cnt)[4]) {
    const unsigned full = 0xffffffffu;
    const int s = lane;
    bool live = (s < SB) && (base + s < N);
    if (live) {
        const double w = inp[IN_WGT * SB + s];
        live = (fabs(w) < 1e300) && !(s_bad[s] & 2) && (w > 0.0);
    }
    const double* c = ctx + (s < SB ? s : 0) * CX_STRIDE;
    const double* sc = scr + (s < SB ? s : 0) * SC_STRIDE;
    int nd = 0, nsvtot = 0, stmask = 0;
    if (live) {
        nd = (int)c[CX_NQ]; nsvtot = (int)c[CX_NQ + 1];
        for (int sl = 0; sl < M.n_ee; ++sl) { const int kf = (int)sc[SC_META + 1 + sl]; if (kf >= 0) stmask |= 1 << kf; }
    }
    int nsv[ST_MAXLEG], nun[ST_MAXLEG], sb[ST_MAXLEG], us[ST_MAXLEG], nA = 0, nB = 0;      // static indices only: registers
#pragma unroll
    for (int leg = 0; leg < ST_MAXLEG; ++leg) {
        nsv[leg] = 0; nun[leg] = 0; sb[leg] = 0; us[leg] = 0;
        if (live && leg < M.nfch) {
            const int kf = M.st_cfoot[leg];
            const bool stance = kf >= 0 && ((stmask >> kf) & 1);
            const int flen = M.fch_len[leg];
            if (stance) {
                const int ln = M.chain_len[kf];
                nsv[leg] = ln - 3; nun[leg] = flen - ln; us[leg] = ln;
                for (int k2 = 0; k2 < MAXEE; ++k2) if (k2 < kf && ((stmask >> k2) & 1) && M.chain_len[k2] > 0) sb[leg] += M.chain_len[k2] - 3;
            } else nun[leg] = flen;
            if (M.st_ccls[leg]) nB += nsv[leg] + nun[leg]; else nA += nsv[leg] + nun[leg];
        }
    }
    // exclusive prefix inside the round (ST_TS consecutive lanes)
    int incD = nd, incA = nA, incB = nB;
#pragma unroll
    for (int k = 1; k < ST_TS; ++k) {
        const int uD = __shfl_up_sync(full, nd, k), uA = __shfl_up_sync(full, nA, k), uB = __shfl_up_sync(full, nB, k);
        if (lane % ST_TS >= k) { incD += uD; incA += uA; incB += uB; }
    }
    if (s >= SB) return;
    const int rd = s / ST_TS;
    if (lane % ST_TS == ST_TS - 1) { cnt[rd][0] = incD; cnt[rd][1] = incA; cnt[rd][2] = incB; }
    const int exD = incD - nd;
    for (int k = 0; k < nd; ++k) descD[rd][exD + k] = st_desc(s, nsvtot + k);
    int oA = incA - nA, oB = incB - nB;
#pragma unroll
    for (int leg = 0; leg < ST_MAXLEG; ++leg) {
        if (leg < M.nfch) {
            const int X = M.st_ccls[leg];
            int o = X ? oB : oA;
            uint32_t* dst = descS[X][rd];
            for (int i = 0; i < nsv[leg]; ++i) dst[o++] = st_desc(s, sb[leg] + i);
            for (int i = 0; i < nun[leg]; ++i) dst[o++] = st_desc(s, 16 + M.fch[leg][0] - 2 + us[leg] + i);
            if (X) oB = o; else oA = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------- tile fill
__device__ __forceinline__ double st_sign(double v) { return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : (v == 0.0 ? 0.0 : v)); }   // numpy sign: sign(nan) = nan

// ten entries of body (joint j) for the row whose Pluecker combination is d, into dst
__device__ __forceinline__ void st_emit(const double* __restrict__ c, int j, const double (&d)[6], double* __restrict__ dst) {
    const double2* X2 = reinterpret_cast<const double2*>(c + CX_X + 12 * (j - 2));
    const double2 x01 = X2[0], x23 = X2[1], x45 = X2[2], x67 = X2[3], x8p = X2[4], p12 = X2[5];
    const double p0 = x8p.y, p1 = p12.x, p2 = p12.y;
    const double u0 = d[0] + (d[4] * p2 - d[5] * p1), u1 = d[1] + (d[5] * p0 - d[3] * p2), u2 = d[2] + (d[3] * p1 - d[4] * p0);
    const double el0 = x01.x * u0 + x23.y * u1 + x67.x * u2, el1 = x01.y * u0 + x45.x * u1 + x67.y * u2, el2 = x23.x * u0 + x45.y * u1 + x8p.x * u2;
    const double ea0 = x01.x * d[3] + x23.y * d[4] + x67.x * d[5], ea1 = x01.y * d[3] + x45.x * d[4] + x67.y * d[5], ea2 = x23.x * d[3] + x45.y * d[4] + x8p.x * d[5];
    body_row(c + CX_B9 + B9S * (j - 1), el0, el1, el2, ea0, ea1, ea2, dst);
}

// One pass of the round's tile fill: dense rows [d0, d0 + Dn) of the round and, when SA / SBn are given, the class rows.
// A WARP takes one task = a run of (at most ST_GROUP) consecutive bodies of one leg for all rows of a tile (lane = row): the pose,
// body motion and Pluecker axes of a sample are then read by all of its rows at once (shared-memory broadcast), the walk down the
// leg is warp-uniform, and the bodies of a run share it (a walk per body made 21 joint steps per six-joint leg, runs of two make
// 12).  Every row -- dense, stored sparse or unit vector (a row of the identity table `unit`) -- is a coefficient vector of 18
// numbers, so ONE code path serves them all; a unit vector simply multiplies zeros.  Tasks (M.st_task, sorted by cost on the host):
// kind 0 dense rows x run; 1 dense rows x (root body, torque column, padding); 2 / 3 class rows x run (the run that holds class slot
// 0 also writes the row's torque block).  Warp w takes tasks w and 2 NW - 1 - w of the sorted list (heavy with light).
template <int NT>
__device__ __forceinline__ void phase_fill_struct(const DevModel& M, const double* __restrict__ ctx, const double* __restrict__ unit,
                                                  double* __restrict__ tileD, double* __restrict__ tileA, double* __restrict__ tileB,
                                                  const uint32_t* __restrict__ descD, const uint32_t* __restrict__ descA,
                                                  const uint32_t* __restrict__ descB, int Dn, int SA, int SBn, int friction, int t) {
    constexpr int NW = NT / 32;
    const int nd = M.nd, warp = t >> 5, lane = t & 31;
    const int ksD = (Dn + 3) >> 2, ksA = (SA + 3) >> 2, ksB = (SBn + 3) >> 2;
    // pad rows of the last k-step of each tile
    for (int e = t; e < (4 * ksD - Dn) * CW; e += NT) tileD[(Dn + e / CW) * TILE_LD + (e % CW)] = 0.0;
    for (int e = t; e < (4 * ksA - SA) * 80; e += NT) tileA[(SA + e / 80) * ST_SLD + (e % 80)] = 0.0;
    for (int e = t; e < (4 * ksB - SBn) * 80; e += NT) tileB[(SBn + e / 80) * ST_SLD + (e % 80)] = 0.0;
    const int ntask = M.st_ntask;
    for (int pass = 0; pass < 2; ++pass) {
        const int ti = pass ? 2 * NW - 1 - warp : warp;
        if (ti >= ntask) continue;
        const uint32_t code = M.st_task[ti];
        const int kind = code & 15;
        if (kind == 1) {
            // ---- dense rows: root body, torque column, padding
            if (lane >= Dn) continue;
            const uint32_t desc = descD[lane];
            const double* c = ctx + (desc & 31) * CX_STRIDE;
            const double* Qk = st_vector(desc, c, unit);
            const double wsq = c[CX_W];
            double* row = tileD + lane * TILE_LD;
            body_row(c + CX_B9, Qk[0] * wsq, Qk[1] * wsq, Qk[2] * wsq, Qk[3] * wsq, Qk[4] * wsq, Qk[5] * wsq, row + ST_ROOTCOL);
            double tau = 0.0;
            for (int jj = 0; jj < nd; ++jj) tau = fma(Qk[6 + jj] * wsq, c[CX_TAU + jj], tau);
            *reinterpret_cast<double2*>(row + ST_TAUCOL) = make_double2(tau, 0.0);
            *reinterpret_cast<double2*>(row + ST_ROOTCOL + 10) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2*>(row + ST_ROOTCOL + 12) = make_double2(0.0, 0.0);
            // class slots no joint uses (trees with fewer than twelve joints)
            for (int X = 0; X < 2; ++X)
                for (int e = M.st_nslot[X]; e < 6; ++e) {
                    double* z = row + 72 * X;
#pragma unroll
                    for (int cc = 0; cc < 10; ++cc) z[10 * e + cc] = 0.0;
                    z[60 + e] = 0.0; z[66 + e] = 0.0;
                }
            continue;
        }
        // ---- rows of one tile x a run of bodies (warp-uniform: tile, pitch, columns, joints)
        const int X = kind - 2;                                     // -2: dense tile
        const int j0 = (code >> 4) & 15, cnt = (code >> 8) & 15, slot0 = (code >> 12) & 15;
        const int nrows = (kind == 0) ? Dn : (X ? SBn : SA);
        if (lane >= nrows) continue;
        const uint32_t desc = ((kind == 0) ? descD : (X ? descB : descA))[lane];
        double* row = (kind == 0) ? tileD + lane * TILE_LD : (X ? tileB : tileA) + lane * ST_SLD;
        const double* c = ctx + (desc & 31) * CX_STRIDE;
        const double* Qk = st_vector(desc, c, unit);
        const double wsq = c[CX_W];
        const int pos0 = st_rec_pos(M.st_jrec[j0]), jf = j0 - pos0;
        double d[6];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) d[cc] = Qk[cc] * wsq;
#pragma unroll
        for (int e2 = 0; e2 < MAXCH - 1; ++e2) {                     // joints above the run: walked, not emitted (warp-uniform guards)
            if (e2 < pos0) {
                const double pj = Qk[4 + jf + e2] * wsq;
                const double2* A2 = reinterpret_cast<const double2*>(c + CX_A + 6 * (jf + e2 - 2));
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) { const double2 ak = A2[cc]; d[2 * cc] = fma(pj, ak.x, d[2 * cc]); d[2 * cc + 1] = fma(pj, ak.y, d[2 * cc + 1]); }
            }
        }
        for (int b = 0; b < cnt; ++b) {
            const int j = j0 + b;
            const double pj = Qk[4 + j] * wsq;
            const double2* A2 = reinterpret_cast<const double2*>(c + CX_A + 6 * (j - 2));
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) { const double2 ak = A2[cc]; d[2 * cc] = fma(pj, ak.x, d[2 * cc]); d[2 * cc + 1] = fma(pj, ak.y, d[2 * cc + 1]); }
            const uint32_t rec = M.st_jrec[j];
            const int bcol = (kind == 0) ? st_rec_bcol(rec) : 10 * (slot0 + b), fvcol = (kind == 0) ? st_rec_fvcol(rec) : 60 + slot0 + b;
            st_emit(c, j, d, row + bcol);
            const double dv = c[CX_DQ + j - 2];
            row[fvcol] = friction ? pj * dv : 0.0;
            row[fvcol + 6] = friction ? pj * st_sign(dv) : 0.0;
        }
        if ((code >> 16) & 1) {
            // torque entry of the class row (coefficients outside its leg are zero), the rest of the torque block, unused slots
            const int ns = M.st_nslot[X];
            double tau = 0.0;
            for (int e2 = 0; e2 < ns; ++e2) { const int j2 = M.st_slotjoint[X][e2]; tau = fma(Qk[4 + j2] * wsq, c[CX_TAU + j2 - 2], tau); }
            double2* t2 = reinterpret_cast<double2*>(row + 72);
            t2[0] = make_double2(tau, 0.0); t2[1] = make_double2(0.0, 0.0); t2[2] = make_double2(0.0, 0.0); t2[3] = make_double2(0.0, 0.0);
            for (int e3 = ns; e3 < 6; ++e3) {
#pragma unroll
                for (int cc = 0; cc < 10; ++cc) row[10 * e3 + cc] = 0.0;
                row[60 + e3] = 0.0; row[66 + e3] = 0.0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- M phase
// k-steps are software-pipelined by hand (the fragments of k-step ks + 1 are loaded before the DMMAs of k-step ks issue, then moved)
// instead of unrolled: unrolling doubles the sixteen per-warp code variants, which are re-fetched every round.
template <int W>
__device__ __forceinline__ void st_mma_warp(const double* __restrict__ tileD, int ksD, const double* __restrict__ tileA, int ksA,
                                            const double* __restrict__ tileB, int ksB, int lane, double (&acc)[STILES_MAX_NT][2]) {
    using T = STiles<W>;
    if (ksD > 0) {
        const double* base = tileD + (lane & 3) * TILE_LD + (lane >> 2);
        double frag[T::NG], nxt[T::NG];
#pragma unroll
        for (int g = 0; g < T::NG; ++g) frag[g] = base[8 * T::G(g)];
#pragma unroll 1
        for (int ks = 0; ks < ksD; ++ks) {
            const double* nb = base + (ks + 1 < ksD ? ks + 1 : ks) * 4 * TILE_LD;
#pragma unroll
            for (int g = 0; g < T::NG; ++g) nxt[g] = nb[8 * T::G(g)];
#pragma unroll
            for (int t = 0; t < T::NT; ++t) dmma884(acc[t][0], acc[t][1], frag[T::IA(t)], frag[T::IB(t)]);
#pragma unroll
            for (int g = 0; g < T::NG; ++g) frag[g] = nxt[g];
        }
    }
    if constexpr (T::NTA > 0) {
        if (ksA > 0) {
            const double* base = tileA + (lane & 3) * ST_SLD + (lane >> 2);
            double frag[T::NGA], nxt[T::NGA];
#pragma unroll
            for (int g = 0; g < T::NGA; ++g) frag[g] = base[8 * T::GA(g)];
#pragma unroll 1
            for (int ks = 0; ks < ksA; ++ks) {
                const double* nb = base + (ks + 1 < ksA ? ks + 1 : ks) * 4 * ST_SLD;
#pragma unroll
                for (int g = 0; g < T::NGA; ++g) nxt[g] = nb[8 * T::GA(g)];
#pragma unroll
                for (int t = 0; t < T::NTA; ++t) dmma884(acc[T::TA(t)][0], acc[T::TA(t)][1], frag[T::IAA(t)], frag[T::IBA(t)]);
#pragma unroll
                for (int g = 0; g < T::NGA; ++g) frag[g] = nxt[g];
            }
        }
    }
    if constexpr (T::NTB > 0) {
        if (ksB > 0) {
            const double* base = tileB + (lane & 3) * ST_SLD + (lane >> 2);
            double frag[T::NGB], nxt[T::NGB];
#pragma unroll
            for (int g = 0; g < T::NGB; ++g) frag[g] = base[8 * T::GB(g)];
#pragma unroll 1
            for (int ks = 0; ks < ksB; ++ks) {
                const double* nb = base + (ks + 1 < ksB ? ks + 1 : ks) * 4 * ST_SLD;
#pragma unroll
                for (int g = 0; g < T::NGB; ++g) nxt[g] = nb[8 * T::GB(g)];
#pragma unroll
                for (int t = 0; t < T::NTB; ++t) dmma884(acc[T::TB(t)][0], acc[T::TB(t)][1], frag[T::IAB(t)], frag[T::IBB(t)]);
#pragma unroll
                for (int g = 0; g < T::NGB; ++g) frag[g] = nxt[g];
            }
        }
    }
}
template <int W = 0>
__device__ __forceinline__ void st_mma_dispatch(int w, const double* tileD, int ksD, const double* tileA, int ksA, const double* tileB,
                                                int ksB, int lane, double (&acc)[STILES_MAX_NT][2]) {
    if constexpr (W < GRAM_WARPS) {
        if (w == W) st_mma_warp<W>(tileD, ksD, tileA, ksA, tileB, ksB, lane, acc);
        else st_mma_dispatch<W + 1>(w, tileD, ksD, tileA, ksA, tileB, ksB, lane, acc);
    }
}
template <int W = 0>
__device__ __forceinline__ void st_store_dispatch(int w, double* __restrict__ partial, int lane, const double (&acc)[STILES_MAX_NT][2]) {
    if constexpr (W < GRAM_WARPS) {
        if (w == W) {
            using T = STiles<W>;
#pragma unroll
            for (int t = 0; t < T::NT; ++t)
                *reinterpret_cast<double2*>(partial + T::ID(t) * 64 + (lane >> 2) * 8 + 2 * (lane & 3)) = make_double2(acc[t][0], acc[t][1]);
        } else st_store_dispatch<W + 1>(w, partial, lane, acc);
    }
}

template <int W = 0>
__device__ __forceinline__ void st_load_dispatch(int w, const double* __restrict__ partial, int lane, double (&acc)[STILES_MAX_NT][2]) {
    if constexpr (W < GRAM_WARPS) {
        if (w == W) {
            using T = STiles<W>;
#pragma unroll
            for (int t = 0; t < T::NT; ++t) {
                const double2 v = *reinterpret_cast<const double2*>(partial + T::ID(t) * 64 + (lane >> 2) * 8 + 2 * (lane & 3));
                acc[t][0] = v.x; acc[t][1] = v.y;
            }
        } else st_load_dispatch<W + 1>(w, partial, lane, acc);
    }
}

// ---------------------------------------------------------------------------------------------- the kernel
// Same launch shape, arguments, partial-Gram layout (tile coordinates) and segmented mode as gram_fused_kernel; the columns of
// the partial Grams are in TILE order, which gram_reduce_kernel undoes (ColMap).
template <bool SEG>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_struct_kernel(const __grid_constant__ DevModel M, const GramArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* tileD = smem;
    double* tileA = smem + ST_TILE_D;
    double* tileB = tileA + ST_TILE_S;
    double* scr = smem;                        // aliases the tiles: only live during the F phases
    double* inp = smem + ST_SB * SC_STRIDE;
    double* ctx = smem + ST_FRONT;
    __shared__ double s_stat[3];
    __shared__ int s_bad[ST_SB];
    __shared__ uint32_t s_descD[ST_NR][ST_DMAX];
    __shared__ uint32_t s_descS[2][ST_NR][ST_SROWS];
    __shared__ int s_cnt[ST_NR][4];
    __shared__ uint32_t s_tmem;
    __shared__ double s_unit[MAXD * QLD];      // identity table: row u = the unit vector of joint u + 2 (coefficient 6 + u)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = tid;
    const bool resume = !SEG && args.accumulate;       // continue from this CTA's partial Gram (host streaming)
    if (tid < 3) s_stat[tid] = resume ? args.partial[(size_t)blockIdx.x * PARTIAL_DOUBLES + GRAM_NTILES * 64 + tid] : 0.0;
    if (tid < ST_SB) s_bad[tid] = 0;
    if (tid < MAXD * QLD) s_unit[tid] = (tid % QLD == 6 + tid / QLD) ? 1.0 : 0.0;
    if (warp == 0) tmem_alloc(&s_tmem, TMEM_PARK_COLS);
    tmem_fence_before_sync();
    double acc[STILES_MAX_NT][2];
    const bool nred12 = M.st_nred <= 12;       // reduced Jacobian has 6 + 3 (legs with a contact frame) columns: 12 for a biped, 18 for a quadruped
    const long long nseg = SEG ? (args.N + args.seg_len - 1) / args.seg_len : 1;
#ifdef SYSID_PHASE_CLOCKS
    long long clkF = 0, clkC = 0, clkM = 0, clk0, clkSub[6] = {0, 0, 0, 0, 0, 0};
#endif
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tpark = s_tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
#ifdef SYSID_PHASE_CLOCKS
    clk0 = clock64();
#endif
    for (long long seg = SEG ? blockIdx.x : 0; seg < nseg; seg += SEG ? gridDim.x : 1) {
    const long long seg_base = SEG ? seg * args.seg_len : 0;
    const long long Nlim = SEG ? min(args.N, seg_base + args.seg_len) : args.N;
    const long long nsb = (Nlim - seg_base + ST_SB - 1) / ST_SB;
    double* partial = args.partial + (size_t)(SEG ? seg : (long long)blockIdx.x) * PARTIAL_DOUBLES;
    bool first_sb = true;
    for (long long sb = SEG ? 0 : blockIdx.x; sb < nsb; sb += SEG ? 1 : gridDim.x) {
        const long long base = seg_base + sb * ST_SB;
        // ---- F phases
        phase_stage<ST_SB, GRAM_THREADS>(M, args.io, base, Nlim, inp, t);
        __syncthreads();
        for (int it = t; it < ST_SB * MAXD; it += GRAM_THREADS) phase_sincos<ST_SB>(M, base, Nlim, inp, ctx, scr, s_bad, it);
        __syncthreads();
        F_TICK(0)
        for (int it = t; it < 2 * ((ST_SB * M.nfch + 31) & ~31); it += GRAM_THREADS) phase_chains<ST_SB>(M, base, Nlim, inp, ctx, scr, it);
        __syncthreads();
        F_TICK(1)
        for (int it = t; it < ST_SB * (M.n_ee + M.nfch * MAXCH); it += GRAM_THREADS) phase_feet_legqr<ST_SB>(M, base, Nlim, ctx, scr, inp, it);
        __syncthreads();
        F_TICK(2)
        if (nred12 && M.n_ee <= 2) { for (int it = t; it < ((8 * ST_SB + 31) & ~31); it += GRAM_THREADS) phase_qbuild_red<ST_SB, 12, 8>(M, base, Nlim, ctx, scr, s_bad, it); }
        else if (nred12) { for (int it = t; it < ((16 * ST_SB + 31) & ~31); it += GRAM_THREADS) phase_qbuild_red<ST_SB, 12, 16>(M, base, Nlim, ctx, scr, s_bad, it); }
        else { for (int it = t; it < ((16 * ST_SB + 31) & ~31); it += GRAM_THREADS) phase_qbuild_red<ST_SB, MAXV, 16>(M, base, Nlim, ctx, scr, s_bad, it); }
        __syncthreads();
        F_TICK(3)
        if (warp == 0) phase_finish<ST_SB>(base, Nlim, inp, ctx, s_bad, t, s_stat);
        else if (warp == GRAM_WARPS - 1) phase_rowdesc<ST_SB>(M, base, Nlim, ctx, scr, inp, s_bad, lane, s_descD, s_descS, s_cnt);
        else if (nred12) phase_qcols_struct<ST_SB, 12>(M, base, Nlim, ctx, scr, inp, t - 32);
        else phase_qcols_struct<ST_SB, MAXV>(M, base, Nlim, ctx, scr, inp, t - 32);
        __syncthreads();
        if (first_sb) {
#pragma unroll
            for (int k = 0; k < STILES_MAX_NT; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }
            if (resume) st_load_dispatch(warp, partial, lane, acc);
            tmem_park<STILES_MAX_NT>(tpark, acc);
        }
        if (t < ST_SB) s_bad[t] = 0;
        PHASE_TICK(clkF)
        first_sb = false;
        prefetch_inputs<ST_SB, GRAM_THREADS>(M, args.io, seg_base + (sb + (SEG ? 1 : (long long)gridDim.x)) * ST_SB, Nlim, t);
        const int nrd = (int)min((long long)ST_NR, (Nlim - base + ST_TS - 1) / ST_TS);
        for (int rd = 0; rd < nrd; ++rd) {
            const int Dtot = s_cnt[rd][0], SA = s_cnt[rd][1], SBn = s_cnt[rd][2];
            for (int d0 = 0; d0 == 0 || d0 < Dtot; d0 += ST_DROWS) {
                const int Dn = min(ST_DROWS, Dtot - d0);
                const bool firstpass = d0 == 0;
                phase_fill_struct<GRAM_THREADS>(M, ctx, s_unit, tileD, tileA, tileB, &s_descD[rd][d0], s_descS[0][rd], s_descS[1][rd], Dn,
                                                firstpass ? SA : 0, firstpass ? SBn : 0, args.friction, t);
                __syncthreads();
                PHASE_TICK(clkC)
                tmem_unpark<STILES_MAX_NT>(tpark, acc);
                st_mma_dispatch(warp, tileD, (Dn + 3) >> 2, tileA, firstpass ? (SA + 3) >> 2 : 0, tileB, firstpass ? (SBn + 3) >> 2 : 0, lane, acc);
                tmem_park<STILES_MAX_NT>(tpark, acc);
                __syncthreads();
                PHASE_TICK(clkM)
            }
        }
    }
    if (!first_sb) tmem_unpark<STILES_MAX_NT>(tpark, acc);
    else {
#pragma unroll
        for (int k = 0; k < STILES_MAX_NT; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }
    }
    st_store_dispatch(warp, partial, lane, acc);
    __syncthreads();
    if (tid == 0) {
        partial[GRAM_NTILES * 64 + 0] = s_stat[0];
        partial[GRAM_NTILES * 64 + 1] = s_stat[1];
        partial[GRAM_NTILES * 64 + 2] = s_stat[2];
        if (SEG) { s_stat[0] = 0.0; s_stat[1] = 0.0; s_stat[2] = 0.0; }
    }
    __syncthreads();
    }   // segments
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_tmem, TMEM_PARK_COLS);
#ifdef SYSID_PHASE_CLOCKS
    if (tid == 0 && !SEG) {
        double* partial = args.partial + (size_t)blockIdx.x * PARTIAL_DOUBLES;
        partial[GRAM_NTILES * 64 + 3] = (double)clkF; partial[GRAM_NTILES * 64 + 4] = (double)clkC; partial[GRAM_NTILES * 64 + 5] = (double)clkM;
        for (int k = 0; k < 6; ++k) partial[GRAM_NTILES * 64 + 6 + k] = (double)clkSub[k];
    }
#endif
}

}  // namespace sysid
