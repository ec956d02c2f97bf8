// Stage 1+2 kernels: fused regressor -> projector -> Gram accumulation (never writes the stacked regressor),
// plus the small debug/compat kernels that DO write per-sample blocks (parity tests, per-sample API).
//
// Fused kernel: one persistent, warp-specialised CTA (512 threads) per SM, ~194 KB shared memory.
//   producers  warps 12..15 (one per SM sub-partition): the F phases of phases.cuh for a super-batch of 16 samples
//              (per-sample context in shared memory), then per round of 2 samples the tile fill: one
//              (sample, body, row) item per thread -> the 18 x 160 projected row block of each sample (36 rows)
//              into one of two tile buffers
//   consumers  warps 0..11 (three per sub-partition): DMMA (mma.sync m8n8k4 f64) rank-36 update of the 160 x 160
//              lower-triangular Gram held in registers (210 8x8 tiles, 17-18 per warp, tables in gram_tiles.inc)
// Producers and consumers hand tile buffers over through named barriers (full / empty per buffer), so the
// latency-bound kinematics overlaps the FP64-pipe-bound contraction.
// The tau column rides along as column c of the row block, so [A b]^T [A b] yields G, r = A^T b and s = b^T b at once.
#pragma once
#include <cuda_runtime.h>
#include "phases.cuh"

namespace sysid {

#include "gram_tiles.inc"

constexpr int TILE_LD = 164;                   // == 4 (mod 16): conflict-free DMMA fragment loads
constexpr int PARTIAL_DOUBLES = GRAM_NTILES * 64 + 8;   // per CTA: tiles, then [wsum, flag0 count, flag1 count, 5 x phase clocks]

// fused kernel geometry
constexpr int GRAM_THREADS = 512;
constexpr int CONS_WARPS = 12, PROD_WARPS = 4;
constexpr int NCONS = CONS_WARPS * 32, NPROD = PROD_WARPS * 32;
static_assert(NCONS + NPROD == GRAM_THREADS, "warp roles");
constexpr int FSB = 16;                        // samples per super-batch (F phases)
constexpr int FTS = 2;                         // samples per tile round
constexpr int FROWS = FTS * MAXV;              // 36 = 9 k-steps of 4
constexpr int FTILE = FROWS * TILE_LD;
constexpr int NBUF = 2;
static_assert(FROWS % 4 == 0 && FSB % FTS == 0, "k-steps of 4 rows");
constexpr int FUSED_SMEM_DOUBLES = NBUF * FTILE + FSB * CX_STRIDE + FSB * SC_STRIDE;
constexpr size_t GRAM_SMEM_BYTES = sizeof(double) * FUSED_SMEM_DOUBLES;
static_assert(GRAM_SMEM_BYTES + 1024 <= 232448, "shared memory budget");

// stacked-matrix / rmse kernels: all 16 warps, 72-row tile
constexpr int TILE_SAMPLES = 4;
constexpr int TILE_ROWS = TILE_SAMPLES * MAXV;
constexpr int TILE_DOUBLES = TILE_ROWS * TILE_LD;
constexpr int STACK_WARPS = GRAM_THREADS / 32;

enum { BAR_PROD = 1, BAR_FULL = 2, BAR_EMPTY = 2 + NBUF };
static_assert(BAR_EMPTY + NBUF <= 16, "named barriers");

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// rank-(4*ksteps) update of this warp's tiles from `tile` (rows x TILE_LD doubles in shared memory)
template <int NW, int W, int MAXNT>
__device__ __forceinline__ void mma_rows(const double* __restrict__ tile, int ksteps, int lane, double (&acc)[MAXNT][2]) {
    using T = WarpTiles<NW, W>;
    const double* base = tile + (lane & 3) * TILE_LD + (lane >> 2);
#pragma unroll 3
    for (int ks = 0; ks < ksteps; ++ks) {
        double frag[T::NG];
#pragma unroll
        for (int g = 0; g < T::NG; ++g) frag[g] = base[ks * 4 * TILE_LD + 8 * T::G(g)];
#pragma unroll
        for (int t = 0; t < T::NT; ++t) dmma884(acc[t][0], acc[t][1], frag[T::IA(t)], frag[T::IB(t)]);
    }
}

template <int NW, int W, int MAXNT>
__device__ __forceinline__ void store_tiles(double* __restrict__ partial, int lane, const double (&acc)[MAXNT][2]) {
    using T = WarpTiles<NW, W>;
#pragma unroll
    for (int t = 0; t < T::NT; ++t) {
        double2 v = make_double2(acc[t][0], acc[t][1]);
        *reinterpret_cast<double2*>(partial + T::ID(t) * 64 + (lane >> 2) * 8 + 2 * (lane & 3)) = v;
    }
}

#define SYSID_WARP_SWITCH12(FN, ...)                                                                            \
    switch (warp) {                                                                                             \
        case 0: FN<12, 0, CONS_MAXNT>(__VA_ARGS__); break;   case 1: FN<12, 1, CONS_MAXNT>(__VA_ARGS__); break;   \
        case 2: FN<12, 2, CONS_MAXNT>(__VA_ARGS__); break;   case 3: FN<12, 3, CONS_MAXNT>(__VA_ARGS__); break;   \
        case 4: FN<12, 4, CONS_MAXNT>(__VA_ARGS__); break;   case 5: FN<12, 5, CONS_MAXNT>(__VA_ARGS__); break;   \
        case 6: FN<12, 6, CONS_MAXNT>(__VA_ARGS__); break;   case 7: FN<12, 7, CONS_MAXNT>(__VA_ARGS__); break;   \
        case 8: FN<12, 8, CONS_MAXNT>(__VA_ARGS__); break;   case 9: FN<12, 9, CONS_MAXNT>(__VA_ARGS__); break;   \
        case 10: FN<12, 10, CONS_MAXNT>(__VA_ARGS__); break; default: FN<12, 11, CONS_MAXNT>(__VA_ARGS__); break; \
    }
#define SYSID_WARP_SWITCH16(FN, ...)                                                                              \
    switch (warp) {                                                                                               \
        case 0: FN<16, 0, STACK_MAXNT>(__VA_ARGS__); break;   case 1: FN<16, 1, STACK_MAXNT>(__VA_ARGS__); break;   \
        case 2: FN<16, 2, STACK_MAXNT>(__VA_ARGS__); break;   case 3: FN<16, 3, STACK_MAXNT>(__VA_ARGS__); break;   \
        case 4: FN<16, 4, STACK_MAXNT>(__VA_ARGS__); break;   case 5: FN<16, 5, STACK_MAXNT>(__VA_ARGS__); break;   \
        case 6: FN<16, 6, STACK_MAXNT>(__VA_ARGS__); break;   case 7: FN<16, 7, STACK_MAXNT>(__VA_ARGS__); break;   \
        case 8: FN<16, 8, STACK_MAXNT>(__VA_ARGS__); break;   case 9: FN<16, 9, STACK_MAXNT>(__VA_ARGS__); break;   \
        case 10: FN<16, 10, STACK_MAXNT>(__VA_ARGS__); break; case 11: FN<16, 11, STACK_MAXNT>(__VA_ARGS__); break; \
        case 12: FN<16, 12, STACK_MAXNT>(__VA_ARGS__); break; case 13: FN<16, 13, STACK_MAXNT>(__VA_ARGS__); break; \
        case 14: FN<16, 14, STACK_MAXNT>(__VA_ARGS__); break; default: FN<16, 15, STACK_MAXNT>(__VA_ARGS__); break; \
    }
constexpr int CONS_MAXNT = WarpTiles<12, -1>::MAX_NT;
constexpr int STACK_MAXNT = WarpTiles<16, -1>::MAX_NT;

// All F phases of one super-batch, executed by a group of NT threads (index t) separated by SYNC().
#define SYSID_F_PHASES(SB, NT, SYNC)                                                                                     \
    for (int it = t; it < SB * M.nfch; it += NT) phase_chains<SB>(M, args.io, base, args.N, ctx, scr, s_bad, it);        \
    SYNC();                                                                                                              \
    for (int it = t; it < SB * MAXEE; it += NT) phase_feet<SB>(M, args.io, base, args.N, ctx, scr, s_bad, it);           \
    SYNC();                                                                                                              \
    for (int it = t; it < SB * (MAXEE * (MAXEE + 1) / 2); it += NT) phase_sblocks<SB>(M, base, args.N, ctx, scr, it);    \
    SYNC();                                                                                                              \
    for (int it = t; it < SB; it += NT) phase_chol<SB>(base, args.N, ctx, scr, s_bad, it);                               \
    SYNC();                                                                                                              \
    for (int it = t; it < SB * MAXV; it += NT) phase_wcols<SB>(M, base, args.N, ctx, scr, it);                           \
    SYNC();                                                                                                              \
    phase_proj<SB, NT>(args.io, base, args.N, ctx, scr, s_bad, t, wsum, nflag0, nflag1);                                 \
    SYNC();

struct GramArgs {
    SampleIO io;
    long long N;
    int friction;
    double* partial;      // [gridDim][PARTIAL_DOUBLES]
};

#ifdef SYSID_PHASE_CLOCKS
#define PHASE_TICK(acc) { const long long now_ = clock64(); acc += now_ - clk0; clk0 = now_; }
#else
#define PHASE_TICK(acc)
#endif

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_fused_kernel(const __grid_constant__ DevModel M, const GramArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* tiles = smem;
    double* ctx = smem + NBUF * FTILE;
    double* scr = ctx + FSB * CX_STRIDE;
    __shared__ double s_wsum;
    __shared__ int s_flag0, s_flag1;
    __shared__ int s_bad[FSB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { s_wsum = 0.0; s_flag0 = 0; s_flag1 = 0; }
    if (tid < FSB) s_bad[tid] = 0;
    const long long nsb = (args.N + FSB - 1) / FSB;
    // tile rounds this CTA runs: every super-batch is full (FSB / FTS rounds) except possibly the globally last one
    long long total_rounds = 0;
    if ((long long)blockIdx.x < nsb) {
        const long long mine = (nsb - 1 - blockIdx.x) / gridDim.x + 1;
        total_rounds = mine * (FSB / FTS);
        if ((nsb - 1) % gridDim.x == blockIdx.x) {
            const long long last = (args.N - (nsb - 1) * FSB + FTS - 1) / FTS;
            total_rounds -= (FSB / FTS) - last;
        }
    }
#ifdef SYSID_PHASE_CLOCKS
    long long clkA = 0, clkB = 0, clkC = 0, clk0;
#endif
    __syncthreads();
#ifdef SYSID_PHASE_CLOCKS
    clk0 = clock64();
#endif
    double* partial = args.partial + (size_t)blockIdx.x * PARTIAL_DOUBLES;
    if (warp >= CONS_WARPS) {
        // ------------------------------------------------------------------ producers
        const int t = tid - NCONS;
        double wsum = 0.0;
        int nflag0 = 0, nflag1 = 0;
        long long rnd = 0;
        auto psync = [] { named_sync(BAR_PROD, NPROD); };
        for (long long sb = blockIdx.x; sb < nsb; sb += gridDim.x) {
            const long long base = sb * FSB;
            SYSID_F_PHASES(FSB, NPROD, psync)
            if (t < FSB) s_bad[t] = 0;
            PHASE_TICK(clkA)
            const int nsub = (int)min((long long)(FSB / FTS), (args.N - base + FTS - 1) / FTS);
            for (int sub = 0; sub < nsub; ++sub, ++rnd) {
                const int b = (int)(rnd % NBUF);
                if (rnd >= NBUF) named_sync(BAR_EMPTY + b, GRAM_THREADS);
                PHASE_TICK(clkC)
                phase_fill<FTS, TILE_LD, NPROD>(M, ctx, tiles + b * FTILE, sub * FTS, args.friction, t);
                __threadfence_block();
                named_arrive(BAR_FULL + b, GRAM_THREADS);
                PHASE_TICK(clkB)
            }
            psync();      // the next super-batch overwrites the context
        }
        if (wsum != 0.0) atomicAdd(&s_wsum, wsum);
        if (nflag0) atomicAdd(&s_flag0, nflag0);
        if (nflag1) atomicAdd(&s_flag1, nflag1);
    } else {
        // ------------------------------------------------------------------ consumers
        double acc[CONS_MAXNT][2];
#pragma unroll
        for (int t = 0; t < CONS_MAXNT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        for (long long rnd = 0; rnd < total_rounds; ++rnd) {
            const int b = (int)(rnd % NBUF);
            named_sync(BAR_FULL + b, GRAM_THREADS);
            PHASE_TICK(clkA)
            const double* tile = tiles + b * FTILE;
            SYSID_WARP_SWITCH12(mma_rows, tile, FROWS / 4, lane, acc)
            if (rnd + NBUF < total_rounds) named_arrive(BAR_EMPTY + b, GRAM_THREADS);
            PHASE_TICK(clkB)
        }
        SYSID_WARP_SWITCH12(store_tiles, partial, lane, acc)
    }
    __syncthreads();
    if (tid == 0) {
        partial[GRAM_NTILES * 64 + 0] = s_wsum;
        partial[GRAM_NTILES * 64 + 1] = (double)s_flag0;
        partial[GRAM_NTILES * 64 + 2] = (double)s_flag1;
    }
#ifdef SYSID_PHASE_CLOCKS
    if (tid == NCONS) { partial[GRAM_NTILES * 64 + 3] = (double)clkA; partial[GRAM_NTILES * 64 + 4] = (double)clkB; partial[GRAM_NTILES * 64 + 5] = (double)clkC; }
    if (tid == 0) { partial[GRAM_NTILES * 64 + 6] = (double)clkA; partial[GRAM_NTILES * 64 + 7] = (double)clkB; }
#endif
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
    double acc[STACK_MAXNT][2];
#pragma unroll
    for (int t = 0; t < STACK_MAXNT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
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
        SYSID_WARP_SWITCH16(mma_rows, tile, TILE_ROWS / 4, lane, acc)
        __syncthreads();
    }
    double* partial = args.partial + (size_t)blockIdx.x * PARTIAL_DOUBLES;
    SYSID_WARP_SWITCH16(store_tiles, partial, lane, acc)
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
constexpr int SB_SAMPLES = 32;   // samples per CTA of the per-sample kernels (one lane each in forward_sample)
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
constexpr size_t RMSE_SMEM_BYTES = sizeof(double) * (TILE_DOUBLES + FSB * CX_STRIDE + FSB * SC_STRIDE);

struct RmseArgs {
    SampleIO io; long long N; const double* phi; double* partial;
};

__global__ void __launch_bounds__(GRAM_THREADS, 1)
rmse_kernel(const __grid_constant__ DevModel M, const RmseArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* tile = smem;
    double* ctx = smem + TILE_DOUBLES;
    double* scr = ctx + FSB * CX_STRIDE;
    __shared__ double s_x[CW];
    __shared__ double s_acc[RMSE_PARTIAL];
    __shared__ int s_bad[FSB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = tid;
    const int np = M.nparams, nd = M.nd;
    if (tid < CW) s_x[tid] = (tid < np) ? args.phi[tid] : ((tid == np + 2 * nd) ? -1.0 : 0.0);
    if (tid < RMSE_PARTIAL) s_acc[tid] = 0.0;
    if (tid < FSB) s_bad[tid] = 0;
    double wsum = 0.0;
    int nflag0 = 0, nflag1 = 0;
    const long long nsb = (args.N + FSB - 1) / FSB;
    __syncthreads();
    for (long long sb = blockIdx.x; sb < nsb; sb += gridDim.x) {
        const long long base = sb * FSB;
        SYSID_F_PHASES(FSB, GRAM_THREADS, __syncthreads)
        if (t < FSB) s_bad[t] = 0;
        const int nsub = (int)min((long long)(FSB / TILE_SAMPLES), (args.N - base + TILE_SAMPLES - 1) / TILE_SAMPLES);
        for (int sub = 0; sub < nsub; ++sub) {
            phase_fill<TILE_SAMPLES, TILE_LD, GRAM_THREADS>(M, ctx, tile, sub * TILE_SAMPLES, 1, t);
            __syncthreads();
            for (int row = warp; row < TILE_ROWS; row += STACK_WARPS) {
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
    (void)wsum; (void)nflag0; (void)nflag1;
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
