// Stage 1+2 kernels: fused regressor -> projector -> Gram accumulation (never writes the stacked regressor),
// plus the small debug/compat kernels that DO write per-sample blocks (parity tests, per-sample API).
//
// Fused kernel, one persistent CTA (512 threads, 16 warps) per SM, ~215 KB shared memory:
//   F phase  warp 0, one lane per sample, 32 samples ("super-batch"): forward_sample() -> per-sample context
//   C phase  all threads, one (sample, column) item each: column_item() -> one 18 x 160 projected row block per sample
//            into the shared tile (4 samples = 72 rows at a time)
//   M phase  all warps: DMMA (mma.sync m8n8k4 f64) rank-72 update of the 160 x 160 lower-triangular Gram held in
//            registers (210 8x8 tiles over 16 warps, 13-14 tiles each, tables in gram_tiles.inc)
// The tau column rides along as column c of the row block, so [A b]^T [A b] yields G, r = A^T b and s = b^T b at once.
#pragma once
#include <cuda_runtime.h>
#include "kinematics.cuh"

namespace sysid {

#include "gram_tiles.inc"

constexpr int GRAM_THREADS = 512;
constexpr int GRAM_WARPS = GRAM_THREADS / 32;
constexpr int SB_SAMPLES = 32;                 // samples per F phase
constexpr int TILE_SAMPLES = 4;                // samples per C/M round
constexpr int TILE_ROWS = TILE_SAMPLES * MAXV; // 72 = 18 k-steps of 4
constexpr int TILE_LD = 164;                   // == 4 (mod 16): conflict-free DMMA fragment loads
constexpr int TILE_DOUBLES = TILE_ROWS * TILE_LD;
static_assert(TILE_ROWS % 4 == 0, "k-steps of 4 rows");
static_assert(SCR_DOUBLES * SCR_LANES <= TILE_DOUBLES, "F-phase scratch aliases the tile");
static_assert(sizeof(double) * (TILE_DOUBLES + SB_SAMPLES * CTX_STRIDE) + 1024 <= 232448, "shared memory budget");
constexpr size_t GRAM_SMEM_BYTES = sizeof(double) * (TILE_DOUBLES + SB_SAMPLES * CTX_STRIDE) + 64;
constexpr int PARTIAL_DOUBLES = GRAM_NTILES * 64 + 8;   // per CTA: tiles, then [wsum, flag0 count, flag1 count]

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// rank-(4*ksteps) update of this warp's tiles from `tile` (rows x TILE_LD doubles in shared memory)
template <int W>
__device__ __forceinline__ void mma_rows(const double* __restrict__ tile, int ksteps, int lane, double (&acc)[GRAM_MAX_NT][2]) {
    using T = WarpTiles<W>;
    const double* base = tile + (lane & 3) * TILE_LD + (lane >> 2);
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
        double frag[T::NG];
#pragma unroll
        for (int g = 0; g < T::NG; ++g) frag[g] = base[ks * 4 * TILE_LD + 8 * T::G(g)];
#pragma unroll
        for (int t = 0; t < T::NT; ++t) dmma884(acc[t][0], acc[t][1], frag[T::IA(t)], frag[T::IB(t)]);
    }
}

template <int W>
__device__ __forceinline__ void store_tiles(double* __restrict__ partial, int lane, const double (&acc)[GRAM_MAX_NT][2]) {
    using T = WarpTiles<W>;
#pragma unroll
    for (int t = 0; t < T::NT; ++t) {
        double2 v = make_double2(acc[t][0], acc[t][1]);
        *reinterpret_cast<double2*>(partial + T::ID(t) * 64 + (lane >> 2) * 8 + 2 * (lane & 3)) = v;
    }
}

#define SYSID_WARP_SWITCH(FN, ...)                                                              \
    switch (warp) {                                                                             \
        case 0: FN<0>(__VA_ARGS__); break;   case 1: FN<1>(__VA_ARGS__); break;                 \
        case 2: FN<2>(__VA_ARGS__); break;   case 3: FN<3>(__VA_ARGS__); break;                 \
        case 4: FN<4>(__VA_ARGS__); break;   case 5: FN<5>(__VA_ARGS__); break;                 \
        case 6: FN<6>(__VA_ARGS__); break;   case 7: FN<7>(__VA_ARGS__); break;                 \
        case 8: FN<8>(__VA_ARGS__); break;   case 9: FN<9>(__VA_ARGS__); break;                 \
        case 10: FN<10>(__VA_ARGS__); break; case 11: FN<11>(__VA_ARGS__); break;               \
        case 12: FN<12>(__VA_ARGS__); break; case 13: FN<13>(__VA_ARGS__); break;               \
        case 14: FN<14>(__VA_ARGS__); break; default: FN<15>(__VA_ARGS__); break;               \
    }

// C phase for TILE_SAMPLES samples starting at local sample s0 of the super-batch.  Items are visited in the
// depth-sorted column order M.colperm, so that the lanes of a warp walk chains of (nearly) equal length.
__device__ __noinline__ void fill_tile(const DevModel& M, const double* __restrict__ ctx, double* __restrict__ tile,
                                       int s0, int friction, int tid, int nthreads) {
    const int np = M.nparams, nd = M.nd;
    const int used = np + (friction ? 2 * nd : 0) + 1;      // columns that carry data (tau column last)
    for (int it = tid; it < TILE_SAMPLES * CW; it += nthreads) {
        const int sl = it / CW, pc = it - sl * CW;
        const int col = M.colperm[pc];
        if (col >= used) continue;
        const double* c = ctx + (s0 + sl) * CTX_STRIDE;
        const double wsq = c[CTX_W];
        double out[MAXV], pval[MAXCH];
        int prow[MAXCH];
        double* dst = tile + (sl * MAXV) * TILE_LD + col;
        if (wsq == 0.0) {
#pragma unroll
            for (int r = 0; r < MAXV; ++r) dst[r * TILE_LD] = 0.0;
            continue;
        }
        // without friction columns the tau column follows the body columns directly
        const int vcol = (!friction && col == np) ? np + 2 * nd : col;
        column_item<true>(M, c, vcol, out, prow, pval);
#pragma unroll
        for (int r = 0; r < MAXV; ++r) dst[r * TILE_LD] = out[r] * wsq;
#pragma unroll
        for (int e = 0; e < MAXCH; ++e) if (prow[e] >= 0) dst[prow[e] * TILE_LD] += pval[e] * wsq;
    }
}

struct GramArgs {
    SampleIO io;
    long long N;
    int friction;
    double* partial;      // [gridDim][PARTIAL_DOUBLES]
};

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_fused_kernel(const __grid_constant__ DevModel M, const GramArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* tile = smem;
    double* ctx = smem + TILE_DOUBLES;
    double* scratch = smem;                    // aliases the tile: only live during the F phase
    __shared__ double s_wsum;
    __shared__ int s_flag0, s_flag1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { s_wsum = 0.0; s_flag0 = 0; s_flag1 = 0; }
    double acc[GRAM_MAX_NT][2];
#pragma unroll
    for (int t = 0; t < GRAM_MAX_NT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    const long long nsb = (args.N + SB_SAMPLES - 1) / SB_SAMPLES;
#ifdef SYSID_PHASE_CLOCKS
    long long clkF = 0, clkC = 0, clkM = 0, clk0;
#define PHASE_TICK(acc) { const long long now_ = clock64(); acc += now_ - clk0; clk0 = now_; }
#else
#define PHASE_TICK(acc)
#endif
    __syncthreads();
#ifdef SYSID_PHASE_CLOCKS
    clk0 = clock64();
#endif
    for (long long sb = blockIdx.x; sb < nsb; sb += gridDim.x) {
        const long long base = sb * SB_SAMPLES;
        if (warp == 0) {
            const long long i = base + lane;
            double* c = ctx + lane * CTX_STRIDE;
            int flags = 0;
            double w = 0.0;
            if (i < args.N) {
                flags = forward_sample(M, args.io, i, scratch + lane, c);
                w = c[CTX_W]; w *= w;
            } else {
                c[CTX_W] = 0.0;
            }
            const unsigned f0 = __ballot_sync(0xffffffffu, flags & 1), f1 = __ballot_sync(0xffffffffu, flags & 2);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            if (lane == 0) { s_wsum += w; s_flag0 += __popc(f0); s_flag1 += __popc(f1); }
        }
        __syncthreads();
        PHASE_TICK(clkF)
        const int nsub = (int)min((long long)(SB_SAMPLES / TILE_SAMPLES), (args.N - base + TILE_SAMPLES - 1) / TILE_SAMPLES);
        for (int sub = 0; sub < nsub; ++sub) {
            fill_tile(M, ctx, tile, sub * TILE_SAMPLES, args.friction, tid, GRAM_THREADS);
            __syncthreads();
            PHASE_TICK(clkC)
            SYSID_WARP_SWITCH(mma_rows, tile, TILE_ROWS / 4, lane, acc)
            __syncthreads();
            PHASE_TICK(clkM)
        }
    }
    double* partial = args.partial + (size_t)blockIdx.x * PARTIAL_DOUBLES;
    SYSID_WARP_SWITCH(store_tiles, partial, lane, acc)
    if (tid == 0) {
#ifdef SYSID_PHASE_CLOCKS
        partial[GRAM_NTILES * 64 + 3] = (double)clkF; partial[GRAM_NTILES * 64 + 4] = (double)clkC; partial[GRAM_NTILES * 64 + 5] = (double)clkM;
#endif
        partial[GRAM_NTILES * 64 + 0] = s_wsum;
        partial[GRAM_NTILES * 64 + 1] = (double)s_flag0;
        partial[GRAM_NTILES * 64 + 2] = (double)s_flag1;
    }
}

// Gram of an already stacked matrix: rows x c (row-major) and b (rows); same M phase, tile filled by plain loads.
struct StackArgs {
    const double* A; const double* b; long long rows; int c; double* partial;
};

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_stack_kernel(const StackArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* tile = smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double acc[GRAM_MAX_NT][2];
#pragma unroll
    for (int t = 0; t < GRAM_MAX_NT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    const int c = args.c;
    const long long nchunks = (args.rows + TILE_ROWS - 1) / TILE_ROWS;
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long r0 = ch * TILE_ROWS;
        for (int it = tid; it < TILE_ROWS * CW; it += GRAM_THREADS) {
            const int r = it / CW, col = it - r * CW;
            const long long gr = r0 + r;
            double v = 0.0;
            if (gr < args.rows) {
                if (col < c) v = args.A[gr * c + col];
                else if (col == c) v = args.b[gr];
            }
            tile[r * TILE_LD + col] = v;
        }
        __syncthreads();
        SYSID_WARP_SWITCH(mma_rows, tile, TILE_ROWS / 4, lane, acc)
        __syncthreads();
    }
    double* partial = args.partial + (size_t)blockIdx.x * PARTIAL_DOUBLES;
    SYSID_WARP_SWITCH(store_tiles, partial, lane, acc)
    if (tid == 0) {
        partial[GRAM_NTILES * 64 + 0] = 0.0; partial[GRAM_NTILES * 64 + 1] = 0.0; partial[GRAM_NTILES * 64 + 2] = 0.0;
    }
}

// Deterministic reduction of the per-CTA partial Grams into stats = [G (c x c) | r (c) | s | n] (ADDS into stats).
// Element (i, j), i >= j, of the (c+1) x (c+1) augmented Gram lives in tile tri(i/8, j/8).
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nparts, int c, double n_rows_per_weight,
                                   double n_add_fixed, double* __restrict__ stats, long long* __restrict__ info) {
    const int ca = c + 1;
    const int total = ca * (ca + 1) / 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) {
        // unrank e -> (i, j), i >= j
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        const int ti = i >> 3, tj = j >> 3;
        const int off = (ti * (ti + 1) / 2 + tj) * 64 + (i & 7) * 8 + (j & 7);
        double sum = 0.0;
        for (int p = 0; p < nparts; ++p) sum += partial[(size_t)p * PARTIAL_DOUBLES + off];
        if (i < c) {
            stats[(size_t)i * c + j] += sum;
            if (i != j) stats[(size_t)j * c + i] += sum;
        } else if (j < c) {
            stats[(size_t)c * c + j] += sum;
        } else {
            stats[(size_t)c * c + c] += sum;
        }
    }
    if (e == 0) {
        double wsum = 0.0, f0 = 0.0, f1 = 0.0;
        for (int p = 0; p < nparts; ++p) {
            wsum += partial[(size_t)p * PARTIAL_DOUBLES + GRAM_NTILES * 64 + 0];
            f0 += partial[(size_t)p * PARTIAL_DOUBLES + GRAM_NTILES * 64 + 1];
            f1 += partial[(size_t)p * PARTIAL_DOUBLES + GRAM_NTILES * 64 + 2];
        }
        stats[(size_t)c * c + c + 1] += n_rows_per_weight * wsum + n_add_fixed;
        if (info) { info[0] += (long long)f0; info[1] += (long long)f1; }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Debug / compat kernels: 32 samples per CTA, per-sample blocks written to global memory.
// MODE 0: raw regressor Y (N x nv x nparams).  MODE 1: projected A (N x nv x ncols), b (N x nv), optional P.
// ------------------------------------------------------------------------------------------------------------
constexpr int DBG_THREADS = 256;
constexpr size_t DBG_SMEM_BYTES = sizeof(double) * (SCR_DOUBLES * SCR_LANES + SB_SAMPLES * CTX_STRIDE);

struct BatchArgs {
    SampleIO io; long long N; int friction;
    double* Y; double* A; double* b; double* P;
};

template <int MODE>
__global__ void __launch_bounds__(DBG_THREADS, 1)
sample_batch_kernel(const __grid_constant__ DevModel M, const BatchArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* scratch = smem;
    double* ctx = smem + SCR_DOUBLES * SCR_LANES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long base = (long long)blockIdx.x * SB_SAMPLES;
    if (warp == 0) {
        const long long i = base + lane;
        if (i < args.N) forward_sample(M, args.io, i, scratch + lane, ctx + lane * CTX_STRIDE);
    }
    __syncthreads();
    const int nv = M.nv, np = M.nparams, nd = M.nd;
    const int ncols = np + ((MODE == 1 && args.friction) ? 2 * nd : 0);
    for (int it = tid; it < SB_SAMPLES * CW; it += DBG_THREADS) {
        const int sl = it / CW, col = it - sl * CW;
        const long long i = base + sl;
        if (i >= args.N) continue;
        const double* c = ctx + sl * CTX_STRIDE;
        double out[MAXV], pval[MAXCH];
        int prow[MAXCH];
        if (MODE == 0) {
            if (col >= np) continue;
            column_item<false>(M, c, col, out, prow, pval);
#pragma unroll
            for (int e = 0; e < MAXCH; ++e) if (prow[e] >= 0) { for (int r = 0; r < MAXV; ++r) if (r == prow[e]) out[r] += pval[e]; }
            for (int r = 0; r < nv; ++r) args.Y[((size_t)i * nv + r) * np + col] = out[r];
        } else {
            const bool is_b = (col == ncols);
            if (col > ncols) continue;
            column_item<true>(M, c, is_b ? np + 2 * nd : col, out, prow, pval);
#pragma unroll
            for (int e = 0; e < MAXCH; ++e) if (prow[e] >= 0) { for (int r = 0; r < MAXV; ++r) if (r == prow[e]) out[r] += pval[e]; }
            if (is_b) { for (int r = 0; r < nv; ++r) args.b[(size_t)i * nv + r] = out[r]; }
            else { for (int r = 0; r < nv; ++r) args.A[((size_t)i * nv + r) * ncols + col] = out[r]; }
        }
    }
    if (MODE == 1 && args.P) {
        for (int it = tid; it < SB_SAMPLES * nv * nv; it += DBG_THREADS) {
            const int sl = it / (nv * nv), rc = it - sl * nv * nv;
            const long long i = base + sl;
            if (i >= args.N) continue;
            const int r = rc / nv, cc = rc - r * nv;
            const double* cx = ctx + sl * CTX_STRIDE;
            const int m3 = (int)cx[CTX_M3];
            double pv = (r == cc) ? 1.0 : 0.0;
            for (int k = 0; k < m3; ++k) pv -= cx[CTX_WM + k * MAXV + r] * cx[CTX_WM + k * MAXV + cc];
            args.P[(size_t)i * nv * nv + rc] = pv;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// tau-prediction error pass (reference print_tau_prediction_rmse): e_i = (P Y phi - P S^T tau)[6:]
// Same F/C phases; instead of the M phase each row of the tile is dotted with [phi; 0; 0; -1].
// partial per CTA: [sum_i ||e_i||^2, per-joint sum of squares (MAXD), count]
// ------------------------------------------------------------------------------------------------------------
constexpr int RMSE_PARTIAL = MAXD + 2;

struct RmseArgs {
    SampleIO io; long long N; const double* phi; double* partial;
};

__global__ void __launch_bounds__(GRAM_THREADS, 1)
rmse_kernel(const __grid_constant__ DevModel M, const RmseArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* tile = smem;
    double* ctx = smem + TILE_DOUBLES;
    double* scratch = smem;
    __shared__ double s_x[CW];
    __shared__ double s_acc[RMSE_PARTIAL];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int np = M.nparams, nd = M.nd;
    if (tid < CW) s_x[tid] = (tid < np) ? args.phi[tid] : ((tid == np + 2 * nd) ? -1.0 : 0.0);
    if (tid < RMSE_PARTIAL) s_acc[tid] = 0.0;
    const long long nsb = (args.N + SB_SAMPLES - 1) / SB_SAMPLES;
    __syncthreads();
    for (long long sb = blockIdx.x; sb < nsb; sb += gridDim.x) {
        const long long base = sb * SB_SAMPLES;
        if (warp == 0) {
            const long long i = base + lane;
            double* c = ctx + lane * CTX_STRIDE;
            if (i < args.N) forward_sample(M, args.io, i, scratch + lane, c);
            else c[CTX_W] = 0.0;
        }
        __syncthreads();
        const int nsub = (int)min((long long)(SB_SAMPLES / TILE_SAMPLES), (args.N - base + TILE_SAMPLES - 1) / TILE_SAMPLES);
        for (int sub = 0; sub < nsub; ++sub) {
            fill_tile(M, ctx, tile, sub * TILE_SAMPLES, 1, tid, GRAM_THREADS);
            __syncthreads();
            for (int row = warp; row < TILE_ROWS; row += GRAM_WARPS) {
                const int rr = row % MAXV;
                if (rr < 6 || rr >= M.nv) continue;
                const long long i = base + sub * TILE_SAMPLES + row / MAXV;
                if (i >= args.N) continue;
                double d = 0.0;
                for (int col = lane; col < CW; col += 32) {
                    const double xv = s_x[col];
                    if (xv != 0.0) d += tile[row * TILE_LD + col] * xv;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (lane == 0) { atomicAdd(&s_acc[0], d * d); atomicAdd(&s_acc[1 + (rr - 6)], d * d); }
            }
            __syncthreads();
        }
    }
    if (tid < RMSE_PARTIAL - 1) args.partial[(size_t)blockIdx.x * RMSE_PARTIAL + tid] = s_acc[tid];
}

__global__ void rmse_finalize_kernel(const double* __restrict__ partial, int nparts, int nd, long long N, double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k > nd) return;
    double sum = 0.0;
    for (int p = 0; p < nparts; ++p) sum += partial[(size_t)p * RMSE_PARTIAL + k];
    out[k] = (k == 0) ? sum / (double)N : sqrt(sum / (double)N);
}

}  // namespace sysid
