// the two public entry points of the reference that no demo calls.
//
//   tsqr_kernel          Solver.solve_llsq_svd (reference src/solver.py:32-39) takes the SVD of the whole stacked regressor
//                        (rows x c, rows = 18 N).  Its Gram form cannot honour pinv's 1e-15 cutoff on a rank-deficient
//                        matrix (the Gram squares the condition number), so the stack is reduced by a communication-
//                        avoiding QR instead: every CTA folds its blocks of 64 rows into a running (c+1) x (c+1) triangle
//                        [R z; 0 rho] of the augmented matrix [A | b] by Householder reflections, the triangles are
//                        merged by the same kernel, and the SVD is taken of the final c x c R (same singular values and
//                        right singular vectors as the stack; z = Q^T b).
//   consistency_kernel   SystemIdentification.get_physical_consistency (reference src/sys_identification.py:324-389) for a
//                        batch of parameter vectors: per link the smallest eigenvalue of I_bar (3x3), the spatial inertia
//                        (6x6), the pseudo inertia J (4x4) and the CoM matrix C (4x4), and tr(J Q).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sysid {

constexpr int TSQR_MAXC = 160;          // c + 1 <= 160 (the Gram kernel's padded width)
constexpr int TSQR_ROWS = 64;           // rows folded per step
constexpr int TSQR_THREADS = 192;       // one thread per column (6 warps), columns >= 160 do not exist
constexpr int TSQR_LDB = TSQR_MAXC + 1; // odd pitch: the column-parallel updates walk rows without bank conflicts
static_assert(TSQR_THREADS >= TSQR_MAXC, "one thread per column");

__host__ __device__ inline size_t tsqr_tri(int i, int ca) { return (size_t)i * ca - (size_t)i * (i - 1) / 2; }   // offset of R[i][i]
inline size_t tsqr_smem_bytes(int ca) { return sizeof(double) * (tsqr_tri(ca, ca) + (size_t)TSQR_ROWS * TSQR_LDB + 8); }

struct TsqrArgs {
    const double* A; long long lda; int ncolsA;      // rows x ncolsA, row-major
    const double* b;                                  // optional extra column (rows), appended as column ncolsA
    long long rows; int ca;                           // ca = columns of the augmented matrix
    double* out;                                      // [gridDim][ca x ca] dense row-major upper triangles
};

__global__ void __launch_bounds__(TSQR_THREADS, 1)
tsqr_kernel(const TsqrArgs g) {
    extern __shared__ __align__(16) double sm[];
    const int ca = g.ca, tid = threadIdx.x;
    double* R = sm;                                   // packed upper triangle, row i at tsqr_tri(i), entries j = i .. ca-1
    double* B = sm + tsqr_tri(ca, ca);                // [TSQR_ROWS][TSQR_LDB]
    double* sh = B + TSQR_ROWS * TSQR_LDB;            // sh[0] = sigma
    for (int e = tid; e < (int)tsqr_tri(ca, ca); e += TSQR_THREADS) R[e] = 0.0;
    const long long nblk = (g.rows + TSQR_ROWS - 1) / TSQR_ROWS;
    // contiguous ranges of blocks per CTA: the merge order is fixed by the grid, the result is deterministic
    const long long per = (nblk + gridDim.x - 1) / gridDim.x;
    const long long b0 = (long long)blockIdx.x * per, b1 = (b0 + per < nblk) ? b0 + per : nblk;
    for (long long blk = b0; blk < b1; ++blk) {
        const long long r0 = blk * TSQR_ROWS;
        __syncthreads();
        for (int e = tid; e < TSQR_ROWS * ca; e += TSQR_THREADS) {
            const int r = e / ca, col = e - r * ca;
            const long long gr = r0 + r;
            double v = 0.0;
            if (gr < g.rows) v = (col < g.ncolsA) ? g.A[gr * g.lda + col] : ((g.b && col == g.ncolsA) ? g.b[gr] : 0.0);
            B[r * TSQR_LDB + col] = v;
        }
        __syncthreads();
        for (int j = 0; j < ca; ++j) {
            // d_k = x . B[:, k] with x = B[:, j]; thread j's own product is sigma = |x|^2
            double d = 0.0;
            double* Rj = R + tsqr_tri(j, ca) - j;                 // Rj[k] = R[j][k]
            const double alpha = Rj[j];                           // read by everyone BEFORE thread j overwrites it below
            if (tid >= j && tid < ca) {
#pragma unroll 8
                for (int r = 0; r < TSQR_ROWS; ++r) d = fma(B[r * TSQR_LDB + j], B[r * TSQR_LDB + tid], d);
                if (tid == j) sh[0] = d;
            }
            __syncthreads();
            const double sigma = sh[0];
            // sigma == 0: the column is already reduced.  Rows of a rank-deficient triangle carry rounding noise that
            // squares at every merge level (1e-16 -> 1e-32 -> ...): below 1e-280 it is dropped instead of being allowed
            // to underflow inside 2 / (v0^2 + sigma)
            if (sigma > 1e-280 && tid >= j && tid < ca) {
                const double nrm = sqrt(fma(alpha, alpha, sigma));
                const double beta = (alpha > 0.0) ? -nrm : nrm;
                const double v0 = alpha - beta;                   // no cancellation: opposite signs
                const double tau = 2.0 / fma(v0, v0, sigma);
                if (tid == j) Rj[j] = beta;
                else {
                    const double w = tau * fma(v0, Rj[tid], d);
                    Rj[tid] = fma(-w, v0, Rj[tid]);
#pragma unroll 8
                    for (int r = 0; r < TSQR_ROWS; ++r) B[r * TSQR_LDB + tid] = fma(-w, B[r * TSQR_LDB + j], B[r * TSQR_LDB + tid]);
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    double* out = g.out + (size_t)blockIdx.x * ca * ca;
    for (int e = tid; e < ca * ca; e += TSQR_THREADS) {
        const int i = e / ca, k = e - i * ca;
        out[e] = (k >= i) ? R[tsqr_tri(i, ca) - i + k] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------ physical consistency
// Cyclic Jacobi eigenvalues of a symmetric n x n matrix (n <= 6, row pitch ld, local memory); returns the smallest.
// Kept out of line and un-unrolled on purpose: inlined into the caller with the 6 x 6 array promoted to registers, nvcc
// 12.9 -O3 for sm_100a produced a wrong rotation sequence (a diagonal entry zeroed; tools/jtest.cu reproduces it, host and
// this form agree to the last digit), so the matrix stays addressable and the loops stay loops.
__host__ __device__ __noinline__ inline double jacobi_min_eig_flat(double* a, int n, int ld) {
#pragma unroll 1
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll 1
        for (int p = 0; p < n; ++p) {
            diag += a[p * ld + p] * a[p * ld + p];
            for (int q = p + 1; q < n; ++q) off += a[p * ld + q] * a[p * ld + q];
        }
        if (off <= 1e-60 || off <= 1e-34 * diag) break;
#pragma unroll 1
        for (int p = 0; p < n - 1; ++p)
#pragma unroll 1
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * ld + q];
                if (apq == 0.0) continue;
                const double theta = (a[q * ld + q] - a[p * ld + p]) / (2.0 * apq);
                const double t = ((theta >= 0.0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll 1
                for (int k = 0; k < n; ++k) {                      // A <- A J
                    const double akp = a[k * ld + p], akq = a[k * ld + q];
                    a[k * ld + p] = c * akp - s * akq; a[k * ld + q] = s * akp + c * akq;
                }
#pragma unroll 1
                for (int k = 0; k < n; ++k) {                      // A <- J^T A
                    const double apk = a[p * ld + k], aqk = a[q * ld + k];
                    a[p * ld + k] = c * apk - s * aqk; a[q * ld + k] = s * apk + c * aqk;
                }
            }
    }
    double m = a[0];
    for (int p = 1; p < n; ++p) m = fmin(m, a[p * ld + p]);
    return m;
}
template <int NMAX>
__host__ __device__ inline double jacobi_min_eig(double (&a)[NMAX][NMAX], int n) { return jacobi_min_eig_flat(&a[0][0], n, NMAX); }

__device__ __forceinline__ double f32r(double x) { return (double)__double2float_rn(x); }     // the reference's np.float32 matrices

struct ConsistencyArgs {
    const double* phi; long long phi_stride; int batch; int num_links;
    const double* semi_axes; const double* centers;    // [num_links][3] each (device)
    double* out;                                        // [batch][5][num_links]: I_bar, I (6x6), J, C, tr(J Q)
};

// one thread per (problem, link)
__global__ void consistency_kernel(const ConsistencyArgs g) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)g.batch * g.num_links) return;
    const int pb = (int)(e / g.num_links), l = (int)(e - (long long)pb * g.num_links);
    const double* p = g.phi + (size_t)pb * g.phi_stride + 10 * l;
    const double m = p[0], h[3] = {p[1], p[2], p[3]};
    const double Ib[3][3] = {{p[4], p[5], p[6]}, {p[5], p[7], p[8]}, {p[6], p[8], p[9]}};     // reference order xx xy xz yy yz zz
    const double s[3] = {g.semi_axes[3 * l], g.semi_axes[3 * l + 1], g.semi_axes[3 * l + 2]};
    const double c[3] = {g.centers[3 * l], g.centers[3 * l + 1], g.centers[3 * l + 2]};
    double* out = g.out + (size_t)pb * 5 * g.num_links + l;
    const int L = g.num_links;
    {   // I_bar stays float64 in the reference (np.array of python floats)
        double a[3][3];
        for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) a[i][k] = Ib[i][k];
        out[0 * L] = jacobi_min_eig<3>(a, 3);
    }
    {   // spatial inertia, float32 entries: [[I_bar, skew(h)], [skew(h)^T, m I]]
        double a[6][6];
        for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) a[i][k] = 0.0;
        for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) a[i][k] = f32r(Ib[i][k]);
        // skew(h) in the upper-right block, its transpose below: rows 0..2 x columns 3..5
        a[0][4] = f32r(-h[2]); a[0][5] = f32r(h[1]);
        a[1][3] = f32r(h[2]);  a[1][5] = f32r(-h[0]);
        a[2][3] = f32r(-h[1]); a[2][4] = f32r(h[0]);
        for (int i = 0; i < 3; ++i) for (int k = 3; k < 6; ++k) a[k][i] = a[i][k];
        for (int i = 0; i < 3; ++i) a[3 + i][3 + i] = f32r(m);
        out[1 * L] = jacobi_min_eig<6>(a, 6);
    }
    double J[4][4];
    {   // pseudo inertia, float32 entries
        const double tr = 0.5 * (Ib[0][0] + Ib[1][1] + Ib[2][2]);
        for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) J[i][k] = f32r(((i == k) ? tr : 0.0) - Ib[i][k]);
        for (int i = 0; i < 3; ++i) { J[i][3] = f32r(h[i]); J[3][i] = f32r(h[i]); }
        J[3][3] = f32r(m);
        double a[4][4];
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) a[i][k] = J[i][k];
        out[2 * L] = jacobi_min_eig<4>(a, 4);
    }
    {   // CoM matrix, float32 entries
        double a[4][4];
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) a[i][k] = 0.0;
        a[0][0] = f32r(m);
        for (int i = 0; i < 3; ++i) { a[0][1 + i] = f32r(h[i] - m * c[i]); a[1 + i][0] = a[0][1 + i]; a[1 + i][1 + i] = f32r(m * (s[i] * s[i])); }
        out[3 * L] = jacobi_min_eig<4>(a, 4);
    }
    {   // tr(J Q): float32 J times float32 Q, products and sums in float32 as numpy does for float32 operands
        float Q[4][4];
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) Q[i][k] = 0.0f;
        double qd[3], qc[3], cqc = 0.0;
        for (int i = 0; i < 3; ++i) { qd[i] = 1.0 / (s[i] * s[i]); qc[i] = qd[i] * c[i]; cqc += c[i] * qc[i]; }
        for (int i = 0; i < 3; ++i) { Q[i][i] = (float)qd[i]; Q[i][3] = (float)qc[i]; Q[3][i] = (float)qc[i]; }
        Q[3][3] = (float)(1.0 - cqc);
        float tr = 0.0f;
        for (int i = 0; i < 4; ++i) {
            float d = 0.0f;
            for (int k = 0; k < 4; ++k) d = __fadd_rn(d, __fmul_rn((float)J[i][k], Q[k][i]));
            tr = __fadd_rn(tr, d);
        }
        out[4 * L] = (double)tr;
    }
}

}  // namespace sysid
