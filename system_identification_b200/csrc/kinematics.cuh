// Per-sample kinematics: stage 1 of the hot path.
//
//   forward_sample()  one THREAD per sample ("F phase"): joint sines/cosines, poses relative to the base,
//                     contact Jacobian J, its Gram S = J J^T, Cholesky S = L L^T, and W = L^-1 J (the null-space
//                     projector is P = I - J^T S^-1 J = I - W^T W); then spatial velocity / gravity-biased
//                     acceleration of every body.  Results go to a small per-sample context in shared memory.
//   column_item()     one THREAD per (sample, regressor column) ("C phase"): the 6-vector column of the body
//                     regressor, walked up the kinematic chain (S_j^T B, then liMi[j].act), projected with
//                     P = I - W^T W on the fly.  Produces one column (MAXV values) of the projected row block
//                     A_i = P [Y | S^T diag(dq) | S^T diag(sign dq) | S^T tau].
//
// Algorithms restated from pinocchio (upstream of reference src/sys_identification.py:395,406,113-135):
// computeJointTorqueRegressor / bodyRegressor / getFrameJacobian(LOCAL_WORLD_ALIGNED); see SURVEY.md App. A.
// Motions are [linear; angular], forces [f; n]; all quantities in LOCAL joint frames.
//
// The contact Jacobian only involves position DIFFERENCES, so the base position drops out and the poses are
// accumulated relative to the base; the result is then rotated by R_b, the matrix Eigen builds from the
// UN-normalised logged quaternion (float32 logs are off unit norm by ~3e-8, which moves P by ~2e-10 -- above the
// parity gate -- if R_b were assumed orthonormal):  J_k = [ R_b | -[r_k]x R_b | (R_b a_c) x (R_b d_kc) ],
// r_k = R_b (foot in base frame), a_c / d_kc = joint axis / lever arm in the base frame.
#pragma once
#include "model.cuh"

namespace sysid {

// ---- per-sample context (doubles), sample-major in shared memory -----------------------------------------
constexpr int CTX_POSE = 0;                    // [MAXD][12] R(9) p(3) of every revolute joint frame relative to the base
constexpr int CTX_B9 = CTX_POSE + 12 * MAXD;   // [MAXB][9]  omega(3), alpha(3), acc(3) of each body (local frame)
constexpr int CTX_DQ = CTX_B9 + 9 * MAXB;      // [MAXD]     joint velocities
constexpr int CTX_TAU = CTX_DQ + MAXD;         // [MAXD]     joint torques
constexpr int CTX_W = CTX_TAU + MAXD;          // sqrt(weight) (0 => sample skipped)
constexpr int CTX_M3 = CTX_W + 1;              // number of contact rows 3m (as a double)
constexpr int CTX_WM = CTX_M3 + 2;             // [3*MAXEE][MAXV] W = L^-1 J_c, row-major: P = I - W^T W (offset even => 16 B rows)
constexpr int CTX_STRIDE = CTX_WM + 3 * MAXEE * MAXV + 2;   // == 2 (mod 4): at most 2-way conflicts for the F-phase accesses
static_assert(CTX_POSE % 2 == 0 && CTX_WM % 2 == 0 && CTX_STRIDE % 4 == 2, "context layout");

// ---- F-phase scratch (doubles per lane), lane-minor: element e of lane l at scr[e * SCR_LANES + l] --------
constexpr int SCR_SC = 0;                            // [MAXD][2] sin, cos of the revolute joints
constexpr int SCR_RF = SCR_SC + 2 * MAXD;            // [MAXEE][3] stance-foot lever arms R_b r_k
constexpr int SCR_JL = SCR_RF + 3 * MAXEE;           // [MAXEE][MAXCH][3] leg columns of J
constexpr int SCR_S = SCR_JL + 3 * MAXEE * MAXCH;    // [78] packed lower triangle of S, then its Cholesky factor L
constexpr int SCR_DOUBLES = SCR_S + 78;              // 186
constexpr int SCR_VA = SCR_RF;                       // second pass reuses everything after the sines: [MAXB][12] v(6) a(6)
static_assert(SCR_VA + 12 * MAXB <= SCR_DOUBLES, "scratch reuse");
constexpr int SCR_LANES = 32;

struct SampleIO {
    const double* q; const double* dq; const double* ddq; const double* tau; const double* cnt;
    const double* weights;   // nullable
    long long ld;
};

#define SCR(e) scr[(e) * SCR_LANES]

__device__ __forceinline__ void joint_rotation_compose(const DevModel& M, int j, double s, double c, double R[9]) {
    // R = pR[j] * Rot(axis_j, angle)   (row-major)
    const double* p = M.pR[j];
    const int jt = M.jtype[j];
    if (jt == JT_RX) {
#pragma unroll
        for (int r = 0; r < 3; ++r) { R[3 * r] = p[3 * r]; R[3 * r + 1] = c * p[3 * r + 1] + s * p[3 * r + 2]; R[3 * r + 2] = c * p[3 * r + 2] - s * p[3 * r + 1]; }
    } else if (jt == JT_RY) {
#pragma unroll
        for (int r = 0; r < 3; ++r) { R[3 * r] = c * p[3 * r] - s * p[3 * r + 2]; R[3 * r + 1] = p[3 * r + 1]; R[3 * r + 2] = s * p[3 * r] + c * p[3 * r + 2]; }
    } else if (jt == JT_RZ) {
#pragma unroll
        for (int r = 0; r < 3; ++r) { R[3 * r] = c * p[3 * r] + s * p[3 * r + 1]; R[3 * r + 1] = c * p[3 * r + 1] - s * p[3 * r]; R[3 * r + 2] = p[3 * r + 2]; }
    } else {  // Rodrigues about a unit axis u: I + s [u]x + (1-c) [u]x^2
        const double ux = M.axis[j][0], uy = M.axis[j][1], uz = M.axis[j][2], t = 1.0 - c;
        const double Q[9] = {1.0 - t * (uy * uy + uz * uz), t * ux * uy - s * uz, t * ux * uz + s * uy,
                             t * ux * uy + s * uz, 1.0 - t * (ux * ux + uz * uz), t * uy * uz - s * ux,
                             t * ux * uz - s * uy, t * uy * uz + s * ux, 1.0 - t * (ux * ux + uy * uy)};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) R[3 * r + k] = p[3 * r] * Q[k] + p[3 * r + 1] * Q[3 + k] + p[3 * r + 2] * Q[6 + k];
    }
}

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

// One thread = one sample.  `scr` already points at this lane's column of the scratch; `ctx` at this sample's context.
// Returns flags: bit0 = contact Jacobian lost rank (dependent row dropped), bit1 = non-finite input (sample skipped).
__device__ __noinline__ int forward_sample(const DevModel& M, const SampleIO& io, long long i, double* scr, double* ctx) {
    int flags = 0;
    const int nb = M.nb, nd = M.nd;
    // ---------------- pass 1: joint angles, poses relative to the base --------------------------------------
    double finite_probe = 0.0;
    double Rb[9];
    {   // Eigen::Quaternion::toRotationMatrix on the raw (x, y, z, w): no normalisation, as pinocchio's free-flyer does
        const double qx = io.q[3 * io.ld + i], qy = io.q[4 * io.ld + i], qz = io.q[5 * io.ld + i], qw = io.q[6 * io.ld + i];
        finite_probe += qx + qy + qz + qw;
        const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
        const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        Rb[0] = 1 - (tyy + tzz); Rb[1] = txy - twz; Rb[2] = txz + twy;
        Rb[3] = txy + twz; Rb[4] = 1 - (txx + tzz); Rb[5] = tyz - twx;
        Rb[6] = txz - twy; Rb[7] = tyz + twx; Rb[8] = 1 - (txx + tyy);
    }
    for (int j = 2; j <= nb; ++j) {
        const double th = io.q[(7 + (j - 2)) * io.ld + i];
        double s, c;
        sincos(th, &s, &c);
        finite_probe += th;
        SCR(SCR_SC + 2 * (j - 2)) = s;
        SCR(SCR_SC + 2 * (j - 2) + 1) = c;
        double R[9];
        joint_rotation_compose(M, j, s, c, R);
        const int lam = M.parent[j];
        double* pose = ctx + CTX_POSE + 12 * (j - 2);
        if (lam == 1) {
#pragma unroll
            for (int k = 0; k < 9; ++k) pose[k] = R[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) pose[9 + k] = M.pp[j][k];
        } else {
            const double* pp_ = ctx + CTX_POSE + 12 * (lam - 2);
            double PR[9], Pp[3];
#pragma unroll
            for (int k = 0; k < 9; ++k) PR[k] = pp_[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) Pp[k] = pp_[9 + k];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int k = 0; k < 3; ++k) pose[3 * r + k] = PR[3 * r] * R[k] + PR[3 * r + 1] * R[3 + k] + PR[3 * r + 2] * R[6 + k];
                pose[9 + r] = Pp[r] + PR[3 * r] * M.pp[j][0] + PR[3 * r + 1] * M.pp[j][1] + PR[3 * r + 2] * M.pp[j][2];
            }
        }
    }
    // ---------------- contact Jacobian, world-aligned (rows of getFrameJacobian(LOCAL_WORLD_ALIGNED)[0:3]) -----
    int foot_of[MAXEE];   // stance slot -> foot index
    int m = 0;
#pragma unroll
    for (int k = 0; k < MAXEE; ++k) foot_of[k] = 0;
#pragma unroll
    for (int k = 0; k < MAXEE; ++k) {
        if (k < M.n_ee) {
            const double cv = io.cnt[k * io.ld + i];
            if (cv != 0.0) {   // truthiness rule of the reference: state 2 counts as stance; NaN is truthy too
#pragma unroll
                for (int t = 0; t < MAXEE; ++t) if (t == m) foot_of[t] = k;
                ++m;
            }
        }
    }
    for (int t = 0; t < m; ++t) {
        int k = 0;
#pragma unroll
        for (int u = 0; u < MAXEE; ++u) if (u == t) k = foot_of[u];
        const int jf = M.ee_joint[k];
        double rf[3];   // foot point in the base frame
        if (jf == 1) {
#pragma unroll
            for (int e = 0; e < 3; ++e) rf[e] = M.ee_off[k][e];
        } else {
            const double* po = ctx + CTX_POSE + 12 * (jf - 2);
#pragma unroll
            for (int e = 0; e < 3; ++e)
                rf[e] = po[9 + e] + po[3 * e] * M.ee_off[k][0] + po[3 * e + 1] * M.ee_off[k][1] + po[3 * e + 2] * M.ee_off[k][2];
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) SCR(SCR_RF + 3 * t + e) = Rb[3 * e] * rf[0] + Rb[3 * e + 1] * rf[1] + Rb[3 * e + 2] * rf[2];
        const int len = M.chain_len[k];
        for (int e = 0; e < len; ++e) {
            const int cj = M.chain[k][e];
            const double* po = ctx + CTX_POSE + 12 * (cj - 2);
            double ax[3];   // joint axis in the base frame
            const int jt = M.jtype[cj];
            if (jt == JT_RU) {
#pragma unroll
                for (int r = 0; r < 3; ++r) ax[r] = po[3 * r] * M.axis[cj][0] + po[3 * r + 1] * M.axis[cj][1] + po[3 * r + 2] * M.axis[cj][2];
            } else {
                const int col = jt - JT_RX;
#pragma unroll
                for (int r = 0; r < 3; ++r) ax[r] = po[3 * r + col];
            }
            const double bx = rf[0] - po[9], by = rf[1] - po[10], bz = rf[2] - po[11];
            const double a0 = Rb[0] * ax[0] + Rb[1] * ax[1] + Rb[2] * ax[2], a1 = Rb[3] * ax[0] + Rb[4] * ax[1] + Rb[5] * ax[2], a2 = Rb[6] * ax[0] + Rb[7] * ax[1] + Rb[8] * ax[2];
            const double dx = Rb[0] * bx + Rb[1] * by + Rb[2] * bz, dy = Rb[3] * bx + Rb[4] * by + Rb[5] * bz, dz = Rb[6] * bx + Rb[7] * by + Rb[8] * bz;
            const int jo = SCR_JL + 3 * (t * MAXCH + e);
            SCR(jo) = a1 * dz - a2 * dy;
            SCR(jo + 1) = a2 * dx - a0 * dz;
            SCR(jo + 2) = a0 * dy - a1 * dx;
        }
    }
    // ---------------- S = J J^T (3m x 3m), packed lower --------------------------------------------------------
    const int m3 = 3 * m;
    double BB[9];   // R_b R_b^T: identity up to the quaternion's deviation from unit norm, kept for parity
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) BB[3 * x + y] = Rb[3 * x] * Rb[3 * y] + Rb[3 * x + 1] * Rb[3 * y + 1] + Rb[3 * x + 2] * Rb[3 * y + 2];
    for (int t = 0; t < m; ++t) {
        int kt = 0;
#pragma unroll
        for (int u = 0; u < MAXEE; ++u) if (u == t) kt = foot_of[u];
        const double rt0 = SCR(SCR_RF + 3 * t), rt1 = SCR(SCR_RF + 3 * t + 1), rt2 = SCR(SCR_RF + 3 * t + 2);
        for (int u2 = 0; u2 <= t; ++u2) {
            int ku = 0;
#pragma unroll
            for (int u = 0; u < MAXEE; ++u) if (u == u2) ku = foot_of[u];
            const double ru0 = SCR(SCR_RF + 3 * u2), ru1 = SCR(SCR_RF + 3 * u2 + 1), ru2 = SCR(SCR_RF + 3 * u2 + 2);
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
                const int ot = SCR_JL + 3 * (t * MAXCH + (lt - ns + e));
                const int ou = SCR_JL + 3 * (u2 * MAXCH + (lu - ns + e));
                const double a0 = SCR(ot), a1 = SCR(ot + 1), a2 = SCR(ot + 2);
                const double b0 = SCR(ou), b1 = SCR(ou + 1), b2 = SCR(ou + 2);
                B[0] += a0 * b0; B[1] += a0 * b1; B[2] += a0 * b2;
                B[3] += a1 * b0; B[4] += a1 * b1; B[5] += a1 * b2;
                B[6] += a2 * b0; B[7] += a2 * b1; B[8] += a2 * b2;
            }
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int y = 0; y < 3; ++y) {
                    const int gi = 3 * t + x, gj = 3 * u2 + y;
                    if (gi >= gj) SCR(SCR_S + tri(gi, gj)) = B[3 * x + y];
                }
        }
    }
    // ---------------- Cholesky S = L L^T in place; dependent rows are dropped (pinv semantics) ---------------
    double maxdiag = 0.0;
    for (int a = 0; a < m3; ++a) maxdiag = fmax(maxdiag, SCR(SCR_S + tri(a, a)));
    const double piv_tol = 1e-13 * maxdiag;
    for (int a = 0; a < m3; ++a) {
        double d = SCR(SCR_S + tri(a, a));
        for (int k = 0; k < a; ++k) { const double l = SCR(SCR_S + tri(a, k)); d -= l * l; }
        double inv;
        if (d > piv_tol) { const double sd = sqrt(d); SCR(SCR_S + tri(a, a)) = sd; inv = 1.0 / sd; }
        else { SCR(SCR_S + tri(a, a)) = 0.0; inv = 0.0; flags |= 1; }    // row a is (numerically) dependent: drop it
        for (int b = a + 1; b < m3; ++b) {
            double v = SCR(SCR_S + tri(b, a));
            for (int k = 0; k < a; ++k) v -= SCR(SCR_S + tri(b, k)) * SCR(SCR_S + tri(a, k));
            SCR(SCR_S + tri(b, a)) = v * inv;
        }
    }
    // ---------------- W = L^-1 J by forward substitution, row by row, straight into the context ----------------
    // P = I - J^T (J J^T)^-1 J = I - W^T W.  Row (t, x) of J: [ R_b[x][:] | (-[r_t]x R_b)[x][:] | leg columns ].
    ctx[CTX_M3] = (double)m3;
    double* Wm = ctx + CTX_WM;
    for (int t = 0; t < m; ++t) {
        int kt = 0;
#pragma unroll
        for (int uu = 0; uu < MAXEE; ++uu) if (uu == t) kt = foot_of[uu];
        const double r0 = SCR(SCR_RF + 3 * t), r1 = SCR(SCR_RF + 3 * t + 1), r2 = SCR(SCR_RF + 3 * t + 2);
        const int len = M.chain_len[kt];
        for (int x = 0; x < 3; ++x) {
            const int k = 3 * t + x;
            double row[MAXV];
#pragma unroll
            for (int c2 = 0; c2 < MAXV; ++c2) row[c2] = 0.0;
            // base columns: R_b[x][:] and (-[r]x R_b)[x][:] = -(r x R_b[:,c])_x
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
                row[c2] = (x == 0) ? Rb[c2] : ((x == 1) ? Rb[3 + c2] : Rb[6 + c2]);
                const double b0 = Rb[c2], b1 = Rb[3 + c2], b2 = Rb[6 + c2];
                // (r x b)_x
                const double cx = (x == 0) ? (r1 * b2 - r2 * b1) : ((x == 1) ? (r2 * b0 - r0 * b2) : (r0 * b1 - r1 * b0));
                row[3 + c2] = -cx;
            }
            // leg columns: scatter by joint id (compile-time register index via the unrolled compare)
            for (int e = 0; e < len; ++e) {
                const int cidx = M.chain[kt][e] + 4;           // idx_v of the joint
                const double val = SCR(SCR_JL + 3 * (t * MAXCH + e) + x);
#pragma unroll
                for (int c2 = 6; c2 < MAXV; ++c2) if (c2 == cidx) row[c2] = val;
            }
            // row_k <- (row_k - sum_{l<k} L[k][l] W[l]) / L[k][k]
            for (int l = 0; l < k; ++l) {
                const double lkl = SCR(SCR_S + tri(k, l));
                const double* Wl = Wm + l * MAXV;
#pragma unroll
                for (int c2 = 0; c2 < MAXV; ++c2) row[c2] = fma(-lkl, Wl[c2], row[c2]);
            }
            const double lkk = SCR(SCR_S + tri(k, k));
            const double inv = (lkk != 0.0) ? 1.0 / lkk : 0.0;     // dropped (dependent) row: W row = 0
            double* Wk = Wm + k * MAXV;
#pragma unroll
            for (int c2 = 0; c2 < MAXV; ++c2) Wk[c2] = row[c2] * inv;
        }
    }
    // ---------------- pass 2: spatial velocities and gravity-biased accelerations (local frames) ---------------
    {
        // root (free-flyer): v = dq[0:6]; a = ddq[0:6] + [R_b^T (-g); 0], R_b from the UN-normalised quaternion (Eigen)
        const double g0 = -M.gravity[0], g1 = -M.gravity[1], g2 = -M.gravity[2];
        double v[6], a[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { v[k] = io.dq[k * io.ld + i]; a[k] = io.ddq[k * io.ld + i]; finite_probe += v[k] + a[k]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] += Rb[k] * g0 + Rb[3 + k] * g1 + Rb[6 + k] * g2;   // R_b^T (-g)
#pragma unroll
        for (int k = 0; k < 6; ++k) { SCR(SCR_VA + k) = v[k]; SCR(SCR_VA + 6 + k) = a[k]; }
        double* b9 = ctx + CTX_B9;
        b9[0] = v[3]; b9[1] = v[4]; b9[2] = v[5];
        b9[3] = a[3]; b9[4] = a[4]; b9[5] = a[5];
        b9[6] = a[0] + (v[4] * v[2] - v[5] * v[1]);
        b9[7] = a[1] + (v[5] * v[0] - v[3] * v[2]);
        b9[8] = a[2] + (v[3] * v[1] - v[4] * v[0]);
    }
    for (int j = 2; j <= nb; ++j) {
        const double s = SCR(SCR_SC + 2 * (j - 2)), c = SCR(SCR_SC + 2 * (j - 2) + 1);
        double R[9];
        joint_rotation_compose(M, j, s, c, R);
        const double px = M.pp[j][0], py = M.pp[j][1], pz = M.pp[j][2];
        const int lam = M.parent[j];
        const int po = SCR_VA + 12 * (lam - 1);
        double pv[6], pa[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { pv[k] = SCR(po + k); pa[k] = SCR(po + 6 + k); }
        // actInv: [R^T (v - p x w); R^T w]
        const double tvx = pv[0] - (py * pv[5] - pz * pv[4]), tvy = pv[1] - (pz * pv[3] - px * pv[5]), tvz = pv[2] - (px * pv[4] - py * pv[3]);
        const double tax = pa[0] - (py * pa[5] - pz * pa[4]), tay = pa[1] - (pz * pa[3] - px * pa[5]), taz = pa[2] - (px * pa[4] - py * pa[3]);
        double v[6], a[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = R[k] * tvx + R[3 + k] * tvy + R[6 + k] * tvz;
            v[3 + k] = R[k] * pv[3] + R[3 + k] * pv[4] + R[6 + k] * pv[5];
            a[k] = R[k] * tax + R[3 + k] * tay + R[6 + k] * taz;
            a[3 + k] = R[k] * pa[3] + R[3 + k] * pa[4] + R[6 + k] * pa[5];
        }
        const double qd = io.dq[(6 + (j - 2)) * io.ld + i], qdd = io.ddq[(6 + (j - 2)) * io.ld + i];
        finite_probe += qd + qdd;
        double wj[3];
        const int jt = M.jtype[j];
#pragma unroll
        for (int k = 0; k < 3; ++k) wj[k] = (jt == JT_RU) ? M.axis[j][k] : ((jt - JT_RX) == k ? 1.0 : 0.0);
        // v_i = actInv(v_parent) + (0; wj qd)
#pragma unroll
        for (int k = 0; k < 3; ++k) v[3 + k] += wj[k] * qd;
        // a_i = actInv(a_parent) + v_i x vJ + (0; wj qdd), with vJ = (0; wj qd): v x vJ = (v_lin x wJ ; w x wJ)
        const double w0 = wj[0] * qd, w1 = wj[1] * qd, w2 = wj[2] * qd;
        a[0] += v[1] * w2 - v[2] * w1; a[1] += v[2] * w0 - v[0] * w2; a[2] += v[0] * w1 - v[1] * w0;
        a[3] += v[4] * w2 - v[5] * w1 + wj[0] * qdd; a[4] += v[5] * w0 - v[3] * w2 + wj[1] * qdd; a[5] += v[3] * w1 - v[4] * w0 + wj[2] * qdd;
        const int o = SCR_VA + 12 * (j - 1);
#pragma unroll
        for (int k = 0; k < 6; ++k) { SCR(o + k) = v[k]; SCR(o + 6 + k) = a[k]; }
        double* b9 = ctx + CTX_B9 + 9 * (j - 1);
        b9[0] = v[3]; b9[1] = v[4]; b9[2] = v[5];
        b9[3] = a[3]; b9[4] = a[4]; b9[5] = a[5];
        b9[6] = a[0] + (v[4] * v[2] - v[5] * v[1]);
        b9[7] = a[1] + (v[5] * v[0] - v[3] * v[2]);
        b9[8] = a[2] + (v[3] * v[1] - v[4] * v[0]);
    }
    for (int k = 0; k < nd; ++k) {
        const double dqk = io.dq[(6 + k) * io.ld + i];
        const double tk = io.tau ? io.tau[k * io.ld + i] : 0.0;
        finite_probe += tk;
        ctx[CTX_DQ + k] = dqk;
        ctx[CTX_TAU + k] = tk;
    }
    double wgt = io.weights ? io.weights[i] : 1.0;
    if (!(fabs(finite_probe) < 1e300)) { flags |= 2; wgt = 0.0; }   // NaN/Inf anywhere: skip the sample
    ctx[CTX_W] = sqrt(fmax(wgt, 0.0));
    return flags;
}
#undef SCR

// One thread = one column of one sample's projected row block  A_i = P [Y | S^T diag(dq) | S^T diag(sign dq) | S^T tau],
// P = I - W^T W.  col in [0, CW): body columns, then viscous, Coulomb, then the tau column.
// The sparse unprojected column y is walked up the chain once; t = W y is accumulated on the fly; the result is
//     out[r] = (y_r for the statically known rows) - sum_k W[k][r] t[k],        r < MAXV
// plus up to MAXCH (row, value) pairs -- the joint rows, known only at run time -- that the caller ADDS to its
// destination after writing out[] (the thread owns the column, so the read-modify-write is private).
// PROJECT=false yields the raw regressor column (W ignored).  Returns the pair count, or -1 for a padding column.
template <bool PROJECT>
__device__ __forceinline__ int column_item(const DevModel& M, const double* ctx, int col, double out[MAXV],
                                           int prow[MAXCH], double pval[MAXCH]) {
    const int np = M.nparams, nd = M.nd;
    const double* Wm = ctx + CTX_WM;
    const int m3 = PROJECT ? (int)ctx[CTX_M3] : 0;
    double t[3 * MAXEE];
#pragma unroll
    for (int k = 0; k < 3 * MAXEE; ++k) t[k] = 0.0;
#pragma unroll
    for (int r = 0; r < MAXV; ++r) out[r] = 0.0;
#pragma unroll
    for (int e = 0; e < MAXCH; ++e) { prow[e] = -1; pval[e] = 0.0; }
    int npairs = 0;
    if (col < np) {
        const int i = col / 10 + 1, k = col - 10 * (i - 1);
        const double* b9 = ctx + CTX_B9 + 9 * (i - 1);
        const double w0 = b9[0], w1 = b9[1], w2 = b9[2], al0 = b9[3], al1 = b9[4], al2 = b9[5], ac0 = b9[6], ac1 = b9[7], ac2 = b9[8];
        double f0 = 0, f1 = 0, f2 = 0, n0 = 0, n1 = 0, n2 = 0;
        if (k == 0) { f0 = ac0; f1 = ac1; f2 = ac2; }
        else if (k < 4) {
            // f = alpha x e + w x (w x e);  n = e x acc
            const double e0 = (k == 1), e1 = (k == 2), e2 = (k == 3);
            const double c0 = w1 * e2 - w2 * e1, c1 = w2 * e0 - w0 * e2, c2 = w0 * e1 - w1 * e0;   // w x e
            f0 = al1 * e2 - al2 * e1 + (w1 * c2 - w2 * c1);
            f1 = al2 * e0 - al0 * e2 + (w2 * c0 - w0 * c2);
            f2 = al0 * e1 - al1 * e0 + (w0 * c1 - w1 * c0);
            n0 = e1 * ac2 - e2 * ac1; n1 = e2 * ac0 - e0 * ac2; n2 = e0 * ac1 - e1 * ac0;
        } else {
            // n = Br(alpha)[:,k'] + w x Br(w)[:,k'],  Br(u) columns (Ixx,Ixy,Iyy,Ixz,Iyz,Izz):
            //   (u0,0,0) (u1,u0,0) (0,u1,0) (u2,0,u0) (0,u2,u1) (0,0,u2)
            double a0, a1, a2, b0, b1, b2;
            switch (k) {
                case 4: a0 = al0; a1 = 0; a2 = 0; b0 = w0; b1 = 0; b2 = 0; break;
                case 5: a0 = al1; a1 = al0; a2 = 0; b0 = w1; b1 = w0; b2 = 0; break;
                case 6: a0 = 0; a1 = al1; a2 = 0; b0 = 0; b1 = w1; b2 = 0; break;
                case 7: a0 = al2; a1 = 0; a2 = al0; b0 = w2; b1 = 0; b2 = w0; break;
                case 8: a0 = 0; a1 = al2; a2 = al1; b0 = 0; b1 = w2; b2 = w1; break;
                default: a0 = 0; a1 = 0; a2 = al2; b0 = 0; b1 = 0; b2 = w2; break;
            }
            n0 = a0 + (w1 * b2 - w2 * b1); n1 = a1 + (w2 * b0 - w0 * b2); n2 = a2 + (w0 * b1 - w1 * b0);
        }
        // Jacobian-transpose form in the BASE frame: rotate the column wrench to base axes at the body origin,
        //   F = R_i f,  N = R_i n   (R_i, p_i = pose of body i relative to the base, from the F phase);
        // the row of ancestor joint j (axis z_j through p_j, both in the base frame) is  z_j . (N + (p_i - p_j) x F),
        // and the six free-flyer rows are (F, N + p_i x F).  Identical to walking S_j^T B up the chain with liMi[j].act.
        double F0 = f0, F1 = f1, F2 = f2, N0 = n0, N1 = n1, N2 = n2, bp0 = 0.0, bp1 = 0.0, bp2 = 0.0;
        if (i > 1) {
            const double2* pq = reinterpret_cast<const double2*>(ctx + CTX_POSE + 12 * (i - 2));
            const double2 a01 = pq[0], a23 = pq[1], a45 = pq[2], a67 = pq[3], a8p = pq[4], p12 = pq[5];
            F0 = a01.x * f0 + a01.y * f1 + a23.x * f2; F1 = a23.y * f0 + a45.x * f1 + a45.y * f2; F2 = a67.x * f0 + a67.y * f1 + a8p.x * f2;
            N0 = a01.x * n0 + a01.y * n1 + a23.x * n2; N1 = a23.y * n0 + a45.x * n1 + a45.y * n2; N2 = a67.x * n0 + a67.y * n1 + a8p.x * n2;
            bp0 = a8p.y; bp1 = p12.x; bp2 = p12.y;
        }
        int j = i;
#pragma unroll
        for (int e = 0; e < MAXCH; ++e) {
            if (j > 1) {
                const int jt = M.jtype[j];
                const double* pj = ctx + CTX_POSE + 12 * (j - 2);
                double z0, z1, z2;
                if (jt == JT_RU) {
                    const double u0 = M.axis[j][0], u1 = M.axis[j][1], u2 = M.axis[j][2];
                    z0 = pj[0] * u0 + pj[1] * u1 + pj[2] * u2; z1 = pj[3] * u0 + pj[4] * u1 + pj[5] * u2; z2 = pj[6] * u0 + pj[7] * u1 + pj[8] * u2;
                } else {
                    const int cc = jt - JT_RX;
                    z0 = pj[cc]; z1 = pj[3 + cc]; z2 = pj[6 + cc];
                }
                const double r0 = bp0 - pj[9], r1 = bp1 - pj[10], r2 = bp2 - pj[11];
                double val = z0 * (N0 + (r1 * F2 - r2 * F1)) + z1 * (N1 + (r2 * F0 - r0 * F2)) + z2 * (N2 + (r0 * F1 - r1 * F0));
                if (e == 0) {   // the body's own joint: S_i^T B in the body frame, exactly (keeps the structural zeros exact)
                    val = (jt == JT_RX) ? n0 : (jt == JT_RY) ? n1 : (jt == JT_RZ) ? n2
                                                             : (M.axis[j][0] * n0 + M.axis[j][1] * n1 + M.axis[j][2] * n2);
                }
                const int row = 6 + (j - 2);
                prow[e] = row; pval[e] = val; npairs = e + 1;
                if (PROJECT) {
#pragma unroll
                    for (int kk = 0; kk < 3 * MAXEE; ++kk) if (kk < m3) t[kk] = fma(Wm[kk * MAXV + row], val, t[kk]);
                }
                j = M.parent[j];
            }
        }
        // free-flyer root: rows 0..5 = (F; N + p_i x F)
        f0 = F0; f1 = F1; f2 = F2;
        n0 = N0 + (bp1 * F2 - bp2 * F1); n1 = N1 + (bp2 * F0 - bp0 * F2); n2 = N2 + (bp0 * F1 - bp1 * F0);
        out[0] = f0; out[1] = f1; out[2] = f2; out[3] = n0; out[4] = n1; out[5] = n2;
        if (PROJECT) {
#pragma unroll
            for (int kk = 0; kk < 3 * MAXEE; ++kk) {
                if (kk < m3) {
                    const double2* w2p = reinterpret_cast<const double2*>(Wm + kk * MAXV);
                    const double2 wa = w2p[0], wb = w2p[1], wc = w2p[2];
                    t[kk] += wa.x * f0 + wa.y * f1 + wb.x * f2 + wb.y * n0 + wc.x * n1 + wc.y * n2;
                }
            }
        }
    } else if (!PROJECT) {
        return -1;
    } else if (col < np + 2 * nd) {
        const int jj = (col - np) % nd;
        const bool coulomb = (col - np) >= nd;
        const double dqv = ctx[CTX_DQ + jj];
        const double sc = coulomb ? ((dqv > 0.0) ? 1.0 : ((dqv < 0.0) ? -1.0 : (dqv == 0.0 ? 0.0 : dqv))) : dqv;   // numpy sign: sign(nan)=nan
        const int row = 6 + jj;
        prow[0] = row; pval[0] = sc; npairs = 1;
#pragma unroll
        for (int kk = 0; kk < 3 * MAXEE; ++kk) if (kk < m3) t[kk] = Wm[kk * MAXV + row] * sc;
    } else if (col == np + 2 * nd) {
#pragma unroll
        for (int jj = 0; jj < MAXD; ++jj) {
            const double tj = (jj < nd) ? ctx[CTX_TAU + jj] : 0.0;
            out[6 + jj] = tj;
#pragma unroll
            for (int kk = 0; kk < 3 * MAXEE; ++kk) if (kk < m3) t[kk] = fma(Wm[kk * MAXV + 6 + jj], tj, t[kk]);
        }
    } else {
        return -1;
    }
    if (PROJECT) {
        // out -= W^T t
#pragma unroll
        for (int kk = 0; kk < 3 * MAXEE; ++kk) {
            if (kk < m3) {
                const double tk = t[kk];
                const double2* w2p = reinterpret_cast<const double2*>(Wm + kk * MAXV);
#pragma unroll
                for (int r2 = 0; r2 < MAXV / 2; ++r2) {
                    const double2 wv = w2p[r2];
                    out[2 * r2] = fma(-wv.x, tk, out[2 * r2]);
                    out[2 * r2 + 1] = fma(-wv.y, tk, out[2 * r2 + 1]);
                }
            }
        }
    }
    return npairs;
}

}  // namespace sysid
