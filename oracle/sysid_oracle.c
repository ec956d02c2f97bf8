/* ORACLE (test infrastructure, not product code) -- C restatement of stages 1-2 of the hot path, used as the
 * CPU baseline (bench.py cpu_baseline / --impl reference) and cross-checked against oracle/dynamics.py in tests.
 * PARITY UNPINNED against pinocchio/numpy themselves (see oracle/dynamics.py header).
 *
 * It follows the reference's per-sample arithmetic, dense and unfused, exactly as the demo loop performs it:
 *   Y = computeJointTorqueRegressor(q, dq, ddq)                 reference src/sys_identification.py:406 (upstream pinocchio)
 *   J_c = stacked getFrameJacobian(LOCAL_WORLD_ALIGNED)[0:3]     reference src/sys_identification.py:119-129
 *   P = I - pinv(J_c) J_c                                        reference src/sys_identification.py:131-135
 *   A_i = P [Y | S^T diag(dq_j) | S^T diag(sign dq_j)], b_i = P S^T tau   reference src/sys_identification.py:401-418
 *   G += A_i^T A_i, r += A_i^T b_i, s += b_i^T b_i               what MOSEK forms from the stack (demo/solo_identification.py:79-84)
 * OpenMP over samples with one private accumulator per thread.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle/sysid_oracle.c -o oracle/_build/libsysid_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXJ 40
#define MAXV 40
#define MAXEE 4

typedef struct {
    int32_t njoints, n_ee;
    const int32_t* parent;   /* [njoints] */
    const int32_t* jtype;    /* 0 FF, 1 RX, 2 RY, 3 RZ, 4 RU */
    const double* axis;      /* [njoints*3] */
    const double* place_R;   /* [njoints*9] */
    const double* place_p;   /* [njoints*3] */
    const int32_t* ee_joint; /* [n_ee] */
    const double* ee_offset; /* [n_ee*3] */
    double gravity[3];
} oracle_tree;

static void matmul3(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void matvec3(const double* A, const double* x, double* y) {
    for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
static void matTvec3(const double* A, const double* x, double* y) {
    for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
static void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

static void joint_rot(const oracle_tree* T, int j, const double* q, int iq, double* R, double* p) {
    int jt = T->jtype[j];
    p[0] = p[1] = p[2] = 0.0;
    if (jt == 0) {
        double x = q[iq + 3], y = q[iq + 4], z = q[iq + 5], w = q[iq + 6];
        double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w;
        double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
        R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
        R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
        R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
        p[0] = q[iq]; p[1] = q[iq + 1]; p[2] = q[iq + 2];
        return;
    }
    double c = cos(q[iq]), s = sin(q[iq]);
    if (jt == 1) { double M[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(R, M, sizeof M); }
    else if (jt == 2) { double M[9] = {c, 0, s, 0, 1, 0, -s, 0, c}; memcpy(R, M, sizeof M); }
    else if (jt == 3) { double M[9] = {c, -s, 0, s, c, 0, 0, 0, 1}; memcpy(R, M, sizeof M); }
    else {
        const double* u = T->axis + 3 * j; double t = 1 - c;
        double M[9] = {1 - t * (u[1] * u[1] + u[2] * u[2]), t * u[0] * u[1] - s * u[2], t * u[0] * u[2] + s * u[1],
                       t * u[0] * u[1] + s * u[2], 1 - t * (u[0] * u[0] + u[2] * u[2]), t * u[1] * u[2] - s * u[0],
                       t * u[0] * u[2] - s * u[1], t * u[1] * u[2] + s * u[0], 1 - t * (u[0] * u[0] + u[1] * u[1])};
        memcpy(R, M, sizeof M);
    }
}

static void joint_axis(const oracle_tree* T, int j, double* ax) {
    int jt = T->jtype[j];
    ax[0] = ax[1] = ax[2] = 0.0;
    if (jt == 4) { ax[0] = T->axis[3 * j]; ax[1] = T->axis[3 * j + 1]; ax[2] = T->axis[3 * j + 2]; }
    else if (jt >= 1) ax[jt - 1] = 1.0;
}

/* symmetric eigen-decomposition by cyclic Jacobi: A (n x n, destroyed -> diagonal), V columns = eigenvectors */
static void jacobi_eig(double* A, double* V, int n) {
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, tot = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { tot += A[i * n + j] * A[i * n + j]; if (i != j) off += A[i * n + j] * A[i * n + j]; }
        if (off <= 1e-30 * tot) break;
        for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) {
            double apq = A[p * n + q];
            if (apq == 0.0) continue;
            double th = (A[q * n + q] - A[p * n + p]) / (2 * apq);
            double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
            double c = 1 / sqrt(t * t + 1), s = t * c;
            for (int k = 0; k < n; ++k) { double x = A[k * n + p], y = A[k * n + q]; A[k * n + p] = c * x - s * y; A[k * n + q] = s * x + c * y; }
            for (int k = 0; k < n; ++k) { double x = A[p * n + k], y = A[q * n + k]; A[p * n + k] = c * x - s * y; A[q * n + k] = s * x + c * y; }
            for (int k = 0; k < n; ++k) { double x = V[k * n + p], y = V[k * n + q]; V[k * n + p] = c * x - s * y; V[k * n + q] = s * x + c * y; }
        }
    }
}

/* one sample: A (nv x ncols, row-major), b (nv); optionally Y (nv x np) and P (nv x nv) */
static void sample_rows(const oracle_tree* T, int nv, int np, int nd, int friction,
                        const double* q, const double* dq, const double* ddq, const double* tau, const double* cnt,
                        double* A, double* b, double* Yout, double* Pout) {
    const int n = T->njoints;
    double liR[MAXJ][9], lip[MAXJ][3], oR[MAXJ][9], op[MAXJ][3], v[MAXJ][6], a[MAXJ][6];
    static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(oR[0], I3, sizeof I3); op[0][0] = op[0][1] = op[0][2] = 0;
    memset(v[0], 0, sizeof v[0]); memset(a[0], 0, sizeof a[0]);
    for (int k = 0; k < 3; ++k) a[0][k] = -T->gravity[k];
    /* forward pass */
    for (int i = 1; i < n; ++i) {
        int lam = T->parent[i];
        int iq = (i == 1) ? 0 : 7 + (i - 2), iv = (i == 1) ? 0 : 6 + (i - 2);
        double Rj[9], pj[3], tmp[3];
        joint_rot(T, i, q, iq, Rj, pj);
        matmul3(T->place_R + 9 * i, Rj, liR[i]);
        matvec3(T->place_R + 9 * i, pj, tmp);
        for (int k = 0; k < 3; ++k) lip[i][k] = T->place_p[3 * i + k] + tmp[k];
        matmul3(oR[lam], liR[i], oR[i]);
        matvec3(oR[lam], lip[i], tmp);
        for (int k = 0; k < 3; ++k) op[i][k] = op[lam][k] + tmp[k];
        double vJ[6] = {0, 0, 0, 0, 0, 0}, aJ[6] = {0, 0, 0, 0, 0, 0};
        if (T->jtype[i] == 0) { for (int k = 0; k < 6; ++k) { vJ[k] = dq[iv + k]; aJ[k] = ddq[iv + k]; } }
        else { double ax[3]; joint_axis(T, i, ax); for (int k = 0; k < 3; ++k) { vJ[3 + k] = ax[k] * dq[iv]; aJ[3 + k] = ax[k] * ddq[iv]; } }
        /* actInv of the parent's motion */
        double pv[6] = {0, 0, 0, 0, 0, 0}, pa[6], c1[3], d[3];
        if (lam > 0) {
            cross3(lip[i], v[lam] + 3, c1);
            for (int k = 0; k < 3; ++k) d[k] = v[lam][k] - c1[k];
            matTvec3(liR[i], d, pv); matTvec3(liR[i], v[lam] + 3, pv + 3);
        }
        cross3(lip[i], a[lam] + 3, c1);
        for (int k = 0; k < 3; ++k) d[k] = a[lam][k] - c1[k];
        matTvec3(liR[i], d, pa); matTvec3(liR[i], a[lam] + 3, pa + 3);
        for (int k = 0; k < 6; ++k) v[i][k] = vJ[k] + pv[k];
        /* a = v x vJ + aJ + actInv(a_parent);  (v,w) x (v2,w2) = (w x v2 + v x w2 ; w x w2) */
        double x1[3], x2[3], x3[3];
        cross3(v[i] + 3, vJ, x1); cross3(v[i], vJ + 3, x2); cross3(v[i] + 3, vJ + 3, x3);
        for (int k = 0; k < 3; ++k) { a[i][k] = x1[k] + x2[k] + aJ[k] + pa[k]; a[i][3 + k] = x3[k] + aJ[3 + k] + pa[3 + k]; }
    }
    /* backward pass: dense Y */
    double Y[MAXV][400];
    for (int r = 0; r < nv; ++r) memset(Y[r], 0, sizeof(double) * np);
    for (int i = n - 1; i >= 1; --i) {
        double B[6][10];
        memset(B, 0, sizeof B);
        const double* vl = v[i]; const double* w = v[i] + 3; const double* al = a[i]; const double* aa = a[i] + 3;
        double acc[3], wxv[3];
        cross3(w, vl, wxv);
        for (int k = 0; k < 3; ++k) acc[k] = al[k] + wxv[k];
        for (int k = 0; k < 3; ++k) B[k][0] = acc[k];
        /* [alpha]x + [w]x[w]x */
        double Sa[9] = {0, -aa[2], aa[1], aa[2], 0, -aa[0], -aa[1], aa[0], 0};
        double Sw[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
        double Sww[9]; matmul3(Sw, Sw, Sww);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) B[r][1 + c] = Sa[3 * r + c] + Sww[3 * r + c];
        double Sacc[9] = {0, -acc[2], acc[1], acc[2], 0, -acc[0], -acc[1], acc[0], 0};
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) B[3 + r][1 + c] = -Sacc[3 * r + c];
        double Bra[3][6] = {{aa[0], aa[1], 0, aa[2], 0, 0}, {0, aa[0], aa[1], 0, aa[2], 0}, {0, 0, 0, aa[0], aa[1], aa[2]}};
        double Brw[3][6] = {{w[0], w[1], 0, w[2], 0, 0}, {0, w[0], w[1], 0, w[2], 0}, {0, 0, 0, w[0], w[1], w[2]}};
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 6; ++c) {
            double s = Bra[r][c];
            for (int k = 0; k < 3; ++k) s += Sw[3 * r + k] * Brw[k][c];
            B[3 + r][4 + c] = s;
        }
        int j = i;
        while (j > 0) {
            int iv = (j == 1) ? 0 : 6 + (j - 2);
            if (T->jtype[j] == 0) { for (int r = 0; r < 6; ++r) for (int c = 0; c < 10; ++c) Y[iv + r][10 * (i - 1) + c] = B[r][c]; }
            else { double ax[3]; joint_axis(T, j, ax); for (int c = 0; c < 10; ++c) Y[iv][10 * (i - 1) + c] = ax[0] * B[3][c] + ax[1] * B[4][c] + ax[2] * B[5][c]; }
            if (T->parent[j] > 0) {
                for (int c = 0; c < 10; ++c) {
                    double f[3] = {B[0][c], B[1][c], B[2][c]}, m3[3] = {B[3][c], B[4][c], B[5][c]}, f2[3], m2[3], cx[3];
                    matvec3(liR[j], f, f2); matvec3(liR[j], m3, m2); cross3(lip[j], f2, cx);
                    for (int k = 0; k < 3; ++k) { B[k][c] = f2[k]; B[3 + k][c] = m2[k] + cx[k]; }
                }
            }
            j = T->parent[j];
        }
    }
    if (Yout) for (int r = 0; r < nv; ++r) memcpy(Yout + (size_t)r * np, Y[r], sizeof(double) * np);
    /* contact Jacobian (rows of stance feet; truthiness rule) and projector */
    double J[3 * MAXEE][MAXV];
    int m3 = 0;
    for (int k = 0; k < T->n_ee; ++k) {
        if (cnt[k] == 0.0) continue;
        int jf = T->ee_joint[k];
        double tmp[3], pf[3];
        matvec3(oR[jf], T->ee_offset + 3 * k, tmp);
        for (int e = 0; e < 3; ++e) pf[e] = op[jf][e] + tmp[e];
        for (int r = 0; r < 3; ++r) memset(J[m3 + r], 0, sizeof(double) * nv);
        for (int c = jf; c > 0; c = T->parent[c]) {
            int iv = (c == 1) ? 0 : 6 + (c - 2);
            double d[3] = {pf[0] - op[c][0], pf[1] - op[c][1], pf[2] - op[c][2]};
            if (T->jtype[c] == 0) {
                for (int r = 0; r < 3; ++r) for (int e = 0; e < 3; ++e) J[m3 + r][iv + e] = oR[c][3 * r + e];
                /* -[d]x R */
                double Sd[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0}, SR[9];
                matmul3(Sd, oR[c], SR);
                for (int r = 0; r < 3; ++r) for (int e = 0; e < 3; ++e) J[m3 + r][iv + 3 + e] = -SR[3 * r + e];
            } else {
                double ax[3], axw[3], col[3];
                joint_axis(T, c, ax); matvec3(oR[c], ax, axw); cross3(axw, d, col);
                for (int r = 0; r < 3; ++r) J[m3 + r][iv] = col[r];
            }
        }
        m3 += 3;
    }
    double P[MAXV][MAXV];
    for (int r = 0; r < nv; ++r) for (int c = 0; c < nv; ++c) P[r][c] = (r == c);
    if (m3 > 0) {
        /* pinv(J) J = J^T U diag(1/lambda) U^T J over eigenpairs of J J^T with lambda above the cutoff */
        double S[144], U[144];
        for (int r = 0; r < m3; ++r) for (int c = 0; c < m3; ++c) { double s = 0; for (int k = 0; k < nv; ++k) s += J[r][k] * J[c][k]; S[r * m3 + c] = s; }
        jacobi_eig(S, U, m3);
        double lmax = 0;
        for (int e = 0; e < m3; ++e) if (S[e * m3 + e] > lmax) lmax = S[e * m3 + e];
        for (int e = 0; e < m3; ++e) {
            double lam = S[e * m3 + e];
            if (!(lam > 1e-13 * lmax)) continue;
            double wv[MAXV];
            for (int k = 0; k < nv; ++k) { double s = 0; for (int r = 0; r < m3; ++r) s += U[r * m3 + e] * J[r][k]; wv[k] = s; }
            for (int r = 0; r < nv; ++r) for (int c = 0; c < nv; ++c) P[r][c] -= wv[r] * wv[c] / lam;
        }
    }
    if (Pout) for (int r = 0; r < nv; ++r) for (int c = 0; c < nv; ++c) Pout[r * nv + c] = P[r][c];
    /* A = P [Y | S^T diag(dq_j) | S^T diag(sign dq_j)],  b = P S^T tau */
    const int ncols = np + (friction ? 2 * nd : 0);
    for (int r = 0; r < nv; ++r) {
        double* Ar = A + (size_t)r * ncols;
        for (int c = 0; c < np; ++c) Ar[c] = 0.0;
        for (int k = 0; k < nv; ++k) { double prk = P[r][k]; if (prk == 0.0) continue; const double* Yk = Y[k]; for (int c = 0; c < np; ++c) Ar[c] += prk * Yk[c]; }
        if (friction) for (int jn = 0; jn < nd; ++jn) {
            double dv = dq[6 + jn], sg = (dv > 0) - (dv < 0);
            Ar[np + jn] = P[r][6 + jn] * dv;
            Ar[np + nd + jn] = P[r][6 + jn] * sg;
        }
        double s = 0;
        for (int jn = 0; jn < nd; ++jn) s += P[r][6 + jn] * tau[jn];
        b[r] = s;
    }
}

/* channel-major inputs (channels x N, leading dimension ld), like the product ABI.
 * stats = [G (c x c) | r (c) | s | n], ADDED into.  Returns the number of threads used. */
int oracle_gram_accumulate(const oracle_tree* T, const double* q, const double* dq, const double* ddq, const double* tau,
                           const double* cnt, int64_t N, int64_t ld, int friction, int nthreads, double* stats) {
    const int nb = T->njoints - 1, nv = 6 + nb - 1, nq = 7 + nb - 1, nd = nb - 1, np = 10 * nb;
    const int c = np + (friction ? 2 * nd : 0);
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();
#endif
        double* G = (double*)calloc((size_t)c * c + c + 2, sizeof(double));
        double* A = (double*)malloc(sizeof(double) * (size_t)nv * c);
        double bq[MAXV], qs[64], dqs[MAXV], ddqs[MAXV], taus[MAXV], cs[MAXEE];
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            for (int k = 0; k < nq; ++k) qs[k] = q[k * ld + i];
            for (int k = 0; k < nv; ++k) { dqs[k] = dq[k * ld + i]; ddqs[k] = ddq[k * ld + i]; }
            for (int k = 0; k < nd; ++k) taus[k] = tau[k * ld + i];
            for (int k = 0; k < T->n_ee; ++k) cs[k] = cnt[k * ld + i];
            sample_rows(T, nv, np, nd, friction, qs, dqs, ddqs, taus, cs, A, bq, NULL, NULL);
            for (int r = 0; r < nv; ++r) {
                const double* Ar = A + (size_t)r * c;
                for (int a2 = 0; a2 < c; ++a2) {
                    const double ar = Ar[a2];
                    double* Gr = G + (size_t)a2 * c;
                    for (int b2 = 0; b2 <= a2; ++b2) Gr[b2] += ar * Ar[b2];
                    G[(size_t)c * c + a2] += ar * bq[r];
                }
                G[(size_t)c * c + c] += bq[r] * bq[r];
            }
        }
#pragma omp critical
        {
            for (int a2 = 0; a2 < c; ++a2) for (int b2 = 0; b2 <= a2; ++b2) {
                stats[(size_t)a2 * c + b2] += G[(size_t)a2 * c + b2];
                if (a2 != b2) stats[(size_t)b2 * c + a2] += G[(size_t)a2 * c + b2];
            }
            for (int a2 = 0; a2 <= c; ++a2) stats[(size_t)c * c + a2] += G[(size_t)c * c + a2];
        }
        free(G); free(A);
    }
    stats[(size_t)c * c + c + 1] += (double)nv * (double)N;
    return used;
}

/* per-sample blocks for cross-checks: A (N x nv x ncols), b (N x nv), Y (N x nv x np, nullable), P (N x nv x nv, nullable) */
void oracle_sample_blocks(const oracle_tree* T, const double* q, const double* dq, const double* ddq, const double* tau,
                          const double* cnt, int64_t N, int64_t ld, int friction, double* A, double* b, double* Y, double* P) {
    const int nb = T->njoints - 1, nv = 6 + nb - 1, nq = 7 + nb - 1, nd = nb - 1, np = 10 * nb;
    const int c = np + (friction ? 2 * nd : 0);
    for (int64_t i = 0; i < N; ++i) {
        double qs[64], dqs[MAXV], ddqs[MAXV], taus[MAXV], cs[MAXEE];
        for (int k = 0; k < nq; ++k) qs[k] = q[k * ld + i];
        for (int k = 0; k < nv; ++k) { dqs[k] = dq[k * ld + i]; ddqs[k] = ddq[k * ld + i]; }
        for (int k = 0; k < nd; ++k) taus[k] = tau[k * ld + i];
        for (int k = 0; k < T->n_ee; ++k) cs[k] = cnt[k * ld + i];
        sample_rows(T, nv, np, nd, friction, qs, dqs, ddqs, taus, cs, A + (size_t)i * nv * c, b + (size_t)i * nv,
                    Y ? Y + (size_t)i * nv * np : NULL, P ? P + (size_t)i * nv * nv : NULL);
    }
}

/* Torque-prediction error pass, reference src/sys_identification.py:421-437 (print_tau_prediction_rmse):
 *   predicted_i = (P Y phi)[6:], measured_i = (P S^T tau)[6:];  out[0] = mean_i ||e_i||^2 (quirk Q7: no root),
 *   out[1 + k] = sqrt(mean_i e_ik^2).  phi multiplies the pinocchio-ordered regressor as it is (quirk Q1). */
int oracle_tau_rmse(const oracle_tree* T, const double* q, const double* dq, const double* ddq, const double* tau,
                    const double* cnt, int64_t N, int64_t ld, const double* phi, int nthreads, double* out) {
    const int nb = T->njoints - 1, nv = 6 + nb - 1, nq = 7 + nb - 1, nd = nb - 1, np = 10 * nb;
    int used = 1;
    double tot = 0.0, pj[MAXV];
    for (int k = 0; k < MAXV; ++k) pj[k] = 0.0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();
#endif
        double* A = (double*)malloc(sizeof(double) * (size_t)nv * np);
        double bq[MAXV], qs[64], dqs[MAXV], ddqs[MAXV], taus[MAXV], cs[MAXEE], ltot = 0.0, lpj[MAXV];
        for (int k = 0; k < MAXV; ++k) lpj[k] = 0.0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            for (int k = 0; k < nq; ++k) qs[k] = q[k * ld + i];
            for (int k = 0; k < nv; ++k) { dqs[k] = dq[k * ld + i]; ddqs[k] = ddq[k * ld + i]; }
            for (int k = 0; k < nd; ++k) taus[k] = tau[k * ld + i];
            for (int k = 0; k < T->n_ee; ++k) cs[k] = cnt[k * ld + i];
            sample_rows(T, nv, np, nd, 0, qs, dqs, ddqs, taus, cs, A, bq, NULL, NULL);
            for (int k = 0; k < nd; ++k) {
                const double* Ar = A + (size_t)(6 + k) * np;
                double e = 0.0;
                for (int c = 0; c < np; ++c) e += Ar[c] * phi[c];
                e -= bq[6 + k];
                ltot += e * e; lpj[k] += e * e;
            }
        }
#pragma omp critical
        {
            tot += ltot;
            for (int k = 0; k < nd; ++k) pj[k] += lpj[k];
        }
        free(A);
    }
    out[0] = tot / (double)N;
    for (int k = 0; k < nd; ++k) out[1 + k] = sqrt(pj[k] / (double)N);
    return used;
}
