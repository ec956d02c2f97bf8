// Stage 1+2 kernels: fused regressor -> projector -> Gram accumulation (never writes the stacked regressor),
// plus the small debug/compat kernels that DO write per-sample blocks (parity tests, per-sample API).
//
// Fused kernel: one persistent CTA (512 threads, 16 warps) per SM, ~218 KB shared memory, phases separated by
// CTA barriers:
//   F phases  (phases.cuh) for a super-batch of FSB samples: stage -> sincos -> chains -> feet -> qbuild -> qcols;
//             the per-sample context (null-space basis Q of J_c, Pluecker axes, poses, body motions) stays in shared
//             memory; the Gram accumulators are parked in the CTA's partial-Gram slot (L2) meanwhile
//   per round of FTS samples:
//     fill    lane per (sample, basis vector, chain group) -> the rows Q^T [Y | friction | tau] of each sample, packed
//             (18 - rank J_c rows per sample, at most 18 FTS = 72 in all, zero-padded to a multiple of 4)
//     M       all warps: DMMA (mma.sync m8n8k4 f64), one k-step per 4 packed rows, on the 160 x 160 lower-triangular Gram
//             held in registers (210 8x8 tiles, 13-14 per warp, tables in gram_tiles.inc)
// Why phased and not warp-specialised: on B200 a DFMA warp that shares an SM sub-partition with saturating DMMA warps
// gets one issue slot per ~80-110 clk (tools/fp64_mix.cu, profiles/fp64_mix_r01.json), so producers starve exactly when
// consumers are busy; a specialised variant of this kernel measured 24-30 Msamples/s (profiles/phase_clocks_r01_*.txt).
// The tau column rides along as column c of the row block, so [A b]^T [A b] yields G, r = A^T b and s = b^T b at once.
#pragma once
#include <cuda_runtime.h>
#include "phases.cuh"
#include "tmem_park.cuh"

namespace sysid {

#include "gram_tiles.inc"

constexpr int TILE_LD = 164;                   // == 4 (mod 16): conflict-free DMMA fragment loads
constexpr int PARTIAL_DOUBLES = GRAM_NTILES * 64 + 16;  // per CTA: tiles, then [wsum, flag0 count, flag1 count, phase clocks]

constexpr int GRAM_THREADS = 512;
constexpr int GRAM_WARPS = GRAM_THREADS / 32;
#ifndef SYSID_MMA_UNROLL
#define SYSID_MMA_UNROLL 2
#endif
#ifndef SYSID_FTS
#define SYSID_FTS 4
#endif
#ifndef SYSID_FSB
#define SYSID_FSB 24
#endif
constexpr int MMA_UNROLL = SYSID_MMA_UNROLL;
constexpr int FTS = SYSID_FTS;                 // samples per tile round (at most: a round also stops at FROWS rows)
constexpr int FSB = SYSID_FSB;                 // samples per super-batch (F phases)
#ifndef SYSID_FILL_DMMA
constexpr int FROWS = FTS * MAXV;              // default: the scalar fill (phase_fill_q), every round takes FTS samples
constexpr int FUSED_W = 0;
#else
constexpr int FROWS = 64;                      // rows of the tile: four stance samples (<= 15 rows each) or three in flight (18)
constexpr int FUSED_W = FTS * MAXB * 44 + MAXV * MAXV;   // base-frame wrench matrices of the round's samples + an identity (proj_phase.cuh)
#endif
constexpr int FTILE = FROWS * TILE_LD;
static_assert(FROWS % 4 == 0 && FSB % FTS == 0 && FROWS >= MAXV, "k-steps of 4 rows; a round holds at least one sample");
// the F-phase scratch aliases the tile (never live together)
constexpr int FUSED_FSCR = FSB * (SC_STRIDE + IN_CHANNELS);     // scratch, then the staged inputs
constexpr int FUSED_FRONT = (FTILE > FUSED_FSCR) ? FTILE : FUSED_FSCR;
constexpr size_t GRAM_SMEM_BYTES = sizeof(double) * (FUSED_FRONT + FUSED_W + FSB * CX_STRIDE);
static_assert(GRAM_SMEM_BYTES + 1024 <= 232448, "shared memory budget");

// stacked-matrix / rmse kernels: 72-row tile
constexpr int TILE_SAMPLES = 4;
constexpr int TILE_ROWS = TILE_SAMPLES * MAXV;
constexpr int TILE_DOUBLES = TILE_ROWS * TILE_LD;
constexpr int RSB = 24;                        // rmse kernel: samples per super-batch

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

}  // namespace sysid
#include "proj_phase.cuh"
namespace sysid {
static_assert(WB == 44, "FUSED_W");

// rank-(4*ksteps) update of this warp's tiles from `tile` (rows x TILE_LD doubles in shared memory)
template <int NW, int W, int MAXNT>
__device__ __forceinline__ void mma_rows(const double* __restrict__ tile, int ksteps, int lane, double (&acc)[MAXNT][2]) {
    using T = WarpTiles<NW, W>;
    const double* base = tile + (lane & 3) * TILE_LD + (lane >> 2);
#pragma unroll MMA_UNROLL
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

template <int NW, int W, int MAXNT>
__device__ __forceinline__ void load_tiles(const double* __restrict__ partial, int lane, double (&acc)[MAXNT][2]) {
    using T = WarpTiles<NW, W>;
#pragma unroll
    for (int t = 0; t < T::NT; ++t) {
        const double2 v = *reinterpret_cast<const double2*>(partial + T::ID(t) * 64 + (lane >> 2) * 8 + 2 * (lane & 3));
        acc[t][0] = v.x; acc[t][1] = v.y;
    }
}

// warp-uniform dispatch to the per-warp tile tables
template <int NW, int MAXNT, int W = 0>
__device__ __forceinline__ void mma_dispatch(int w, const double* __restrict__ tile, int ksteps, int lane, double (&acc)[MAXNT][2]) {
    if constexpr (W < NW) {
        if (w == W) mma_rows<NW, W, MAXNT>(tile, ksteps, lane, acc);
        else mma_dispatch<NW, MAXNT, W + 1>(w, tile, ksteps, lane, acc);
    }
}
template <int NW, int MAXNT, int W = 0>
__device__ __forceinline__ void load_dispatch(int w, const double* __restrict__ partial, int lane, double (&acc)[MAXNT][2]) {
    if constexpr (W < NW) {
        if (w == W) load_tiles<NW, W, MAXNT>(partial, lane, acc);
        else load_dispatch<NW, MAXNT, W + 1>(w, partial, lane, acc);
    }
}
template <int NW, int MAXNT, int W = 0>
__device__ __forceinline__ void store_dispatch(int w, double* __restrict__ partial, int lane, const double (&acc)[MAXNT][2]) {
    if constexpr (W < NW) {
        if (w == W) store_tiles<NW, W, MAXNT>(partial, lane, acc);
        else store_dispatch<NW, MAXNT, W + 1>(w, partial, lane, acc);
    }
}
constexpr int GRAM_MAXNT = WarpTiles<16, -1>::MAX_NT;

// All F phases of one super-batch, executed by a group of NT threads (index t) separated by SYNC().
#define SYSID_F_PHASES(SB, NT, SYNC, CONTACT, AFTER_STAGE)                                                               \
    phase_stage<SB, NT>(M, args.io, base, Nlim, inp, t);                                                               \
    AFTER_STAGE                                                                                                          \
    SYNC();                                                                                                              \
    for (int it = t; it < SB * MAXD; it += NT) phase_sincos<SB>(M, base, Nlim, inp, ctx, scr, s_bad, it);              \
    SYNC();                                                                                                              \
    F_TICK(0)                                                                                                            \
    for (int it = t; it < 2 * ((SB * M.nfch + 31) & ~31); it += NT) phase_chains<SB>(M, base, Nlim, inp, ctx, scr, it); \
    SYNC();                                                                                                              \
    F_TICK(1)                                                                                                            \
    for (int it = t; it < SB * MAXEE * (MAXCH + 1); it += NT) phase_feet<SB>(M, base, Nlim, inp, ctx, scr, it);        \
    SYNC();                                                                                                              \
    F_TICK(2)                                                                                                            \
    CONTACT(SB, NT, SYNC)

// contact part, rmse kernel: S = J J^T -> Cholesky -> W = L^-1 J -> packed projector P = I - W^T W
#define SYSID_CONTACT_PROJ(SB, NT, SYNC)                                                                                 \
    for (int it = t; it < SB * (MAXEE * (MAXEE + 1) / 2); it += NT) phase_sblocks<SB>(M, base, Nlim, ctx, scr, it);    \
    SYNC();                                                                                                              \
    for (int it = t; it < SB; it += NT) phase_chol<SB>(base, Nlim, ctx, scr, s_bad, it);                               \
    SYNC();                                                                                                              \
    F_TICK(3)                                                                                                            \
    for (int it = t; it < SB * MAXV; it += NT) phase_wcols<SB>(M, base, Nlim, ctx, scr, it);                           \
    SYNC();                                                                                                              \
    F_TICK(4)                                                                                                            \
    phase_proj<SB, NT>(base, Nlim, inp, ctx, scr, s_bad, t, s_stat);                                                   \
    SYNC();
// contact part, Gram kernel: Householder QR of J_c^T -> orthonormal basis Q of null(J_c)
#define SYSID_CONTACT_QBASIS(SB, NT, SYNC)                                                                               \
    for (int it = t; it < ((16 * SB + 31) & ~31); it += NT) phase_qbuild<SB>(M, base, Nlim, ctx, scr, s_bad, it);      \
    SYNC();                                                                                                              \
    F_TICK(3)                                                                                                            \
    phase_finish<SB>(base, Nlim, inp, ctx, s_bad, t, s_stat);                                                          \
    for (int it = t; it < SB * NQMAX; it += NT) phase_qcols<SB>(base, Nlim, ctx, scr, it);                             \
    SYNC();

struct GramArgs {
    SampleIO io;
    long long N;
    int friction;
    double* partial;      // [gridDim][PARTIAL_DOUBLES]; segmented mode: [segments][PARTIAL_DOUBLES]
    long long seg_len;    // 0: one Gram of the whole launch (super-batches dealt round-robin to the CTAs).  > 0: one Gram per SEGMENT of
                          // seg_len consecutive samples (block bootstrap): segments dealt round-robin, each worked through by one CTA
    int accumulate;       // gram_struct_kernel, whole-launch mode: start from the CTA's partial Gram in `partial` instead of zero (host
                          // streaming: the chunks of a log add up in the partials and are reduced once, not once per chunk)
};

#ifdef SYSID_PHASE_CLOCKS
#define PHASE_TICK(acc) { const long long now_ = clock64(); acc += now_ - clk0; clk0 = now_; }
#define F_TICK(k) PHASE_TICK(clkSub[k])
#else
#define PHASE_TICK(acc)
#define F_TICK(k)
#endif

// Tile fill policy: default = the scalar fill (phase_fill_q).  -DSYSID_FILL_DMMA builds the tensor-pipe fill of
// proj_phase.cuh (wbuild + proj): parity-green and a third of the scalar warp instructions per row, but measured SLOWER inside
// this phased kernel (47.7 vs 52.7 Msamples/s on 262 144 G1 samples, profiles/gram_fused_r02_dmma_fill.txt): every phase between
// two CTA barriers is latency-bound at 16 warps per SM, and the M phase already runs at 98 % of the DMMA pipe, so trading scalar
// work for DMMA work only pays once the fill OVERLAPS the contraction.
// Accumulator parking policy: default = the accumulators live in tensor memory and visit registers only for the M phases
// (every other phase gets the whole register file: zero spills).  -DSYSID_PARK_F_ONLY: parked only across the F phases;
// -DSYSID_PARK_L2: the pre-TMEM path through the partial-Gram slot in L2.  Measured on the 1M-sample G1 log:
// 46.9 (L2) -> 50.4 (F only) -> 53.4 Msamples/s (default).
#if !defined(SYSID_PARK_L2) && !defined(SYSID_PARK_F_ONLY)
#define SYSID_PARK_FILL 1
#endif

// SEG = false: one Gram of the whole launch; SEG = true: one Gram per segment of args.seg_len samples (block bootstrap).  A template
// parameter, not a runtime flag: the bookkeeping of the segmented mode costs the whole-launch kernel registers it does not have
// (128 per thread, 0 spills; with a runtime flag ptxas spilled 28 bytes and the 1 M-sample launch took 2.5 % longer).
template <bool SEG>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_fused_kernel(const __grid_constant__ DevModel M, const GramArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* tile = smem;
    double* scr = smem;                        // aliases the tile: only live during the F phases
    double* inp = smem + FSB * SC_STRIDE;
    double* Wsm = smem + FUSED_FRONT;          // wrench matrices of the current round (empty in the scalar-fill build)
    double* ctx = smem + FUSED_FRONT + FUSED_W;
    __shared__ double s_stat[3];               // sum of weights, rank-loss count, skipped count
    __shared__ int s_bad[FSB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = tid;
    if (tid < 3) s_stat[tid] = 0.0;
    if (tid < FSB) s_bad[tid] = 0;
#ifdef SYSID_FILL_DMMA
    double* ident18 = Wsm + FTS * MAXB * 44;   // basis of a sample in flight (Q = I is not stored per sample)
    for (int e = tid; e < MAXV * MAXV; e += GRAM_THREADS) ident18[e] = (e / MAXV == e % MAXV) ? 1.0 : 0.0;
#endif
#ifndef SYSID_PARK_L2
    __shared__ uint32_t s_tmem;
    if (warp == 0) tmem_alloc(&s_tmem, TMEM_PARK_COLS);
    tmem_fence_before_sync();
#endif
    double acc[GRAM_MAXNT][2];
    constexpr bool segmented = SEG;
    const long long nseg = segmented ? (args.N + args.seg_len - 1) / args.seg_len : 1;
#if !defined(SYSID_PARK_FILL)
    if (segmented) return;                     // diagnostic parking policies: whole-launch mode only (the host refuses earlier)
#endif
#ifdef SYSID_PHASE_CLOCKS
    long long clkF = 0, clkC = 0, clkM = 0, clk0, clkSub[6] = {0, 0, 0, 0, 0, 0};
#endif
    __syncthreads();
#ifndef SYSID_PARK_L2
    tmem_fence_after_sync();
    // this warp's parking columns: lane quarter warp % 4 (the only TMEM lanes a warp can reach), column block warp / 4
    const uint32_t tpark = s_tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
#endif
#ifdef SYSID_PHASE_CLOCKS
    clk0 = clock64();
#endif
    for (long long seg = segmented ? blockIdx.x : 0; seg < nseg; seg += segmented ? gridDim.x : 1) {
    const long long seg_base = segmented ? seg * args.seg_len : 0;
    const long long Nlim = segmented ? min(args.N, seg_base + args.seg_len) : args.N;
    const long long nsb = (Nlim - seg_base + FSB - 1) / FSB;
    double* partial = args.partial + (size_t)(segmented ? seg : (long long)blockIdx.x) * PARTIAL_DOUBLES;
    bool first_sb = true;
    for (long long sb = segmented ? 0 : blockIdx.x; sb < nsb; sb += segmented ? 1 : gridDim.x) {
        const long long base = seg_base + sb * FSB;
        // The Gram accumulators (56 registers per thread) are parked in TENSOR MEMORY (tcgen05.st / tcgen05.ld, 115 KB of
        // the SM's otherwise idle 256 KB) whenever no M phase is running, so that the F phases and the tile fill get the
        // whole register file instead of spilling around them; see the policy note above the kernel.
#if defined(SYSID_PARK_L2)
#define SYSID_PARK_ACC if (!first_sb) store_dispatch<GRAM_WARPS, GRAM_MAXNT>(warp, partial, lane, acc);
#elif defined(SYSID_PARK_FILL)
#define SYSID_PARK_ACC                                   // already parked: the accumulators only visit registers for the M phases
#else
#define SYSID_PARK_ACC if (!first_sb) tmem_park<GRAM_MAXNT>(tpark, acc);
#endif
        SYSID_F_PHASES(FSB, GRAM_THREADS, __syncthreads, SYSID_CONTACT_QBASIS, SYSID_PARK_ACC)
#if defined(SYSID_PARK_FILL)
        if (first_sb) {
#pragma unroll
            for (int k = 0; k < GRAM_MAXNT; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }
            tmem_park<GRAM_MAXNT>(tpark, acc);
        }
#else
        if (!first_sb) {
#ifdef SYSID_PARK_L2
            load_dispatch<GRAM_WARPS, GRAM_MAXNT>(warp, partial, lane, acc);
#else
            tmem_unpark<GRAM_MAXNT>(tpark, acc);
#endif
        } else {
#pragma unroll
            for (int k = 0; k < GRAM_MAXNT; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }
        }
#endif
        if (t < FSB) s_bad[t] = 0;
        PHASE_TICK(clkF)
        first_sb = false;
        prefetch_inputs<FSB, GRAM_THREADS>(M, args.io, seg_base + (sb + (segmented ? 1 : (long long)gridDim.x)) * FSB, Nlim, t);     // lands in L2 during the rounds below
#ifndef SYSID_FILL_DMMA
        const int nsub = (int)min((long long)(FSB / FTS), (Nlim - base + FTS - 1) / FTS);
        for (int sub = 0; sub < nsub; ++sub) {
            const int ksteps = phase_fill_q<FTS, TILE_LD, GRAM_THREADS>(M, ctx, tile, sub * FTS, args.friction, t);
            __syncthreads();
            PHASE_TICK(clkC)
#if defined(SYSID_PARK_FILL)
            tmem_unpark<GRAM_MAXNT>(tpark, acc);
#endif
            mma_dispatch<GRAM_WARPS, GRAM_MAXNT>(warp, tile, ksteps, lane, acc);
#if defined(SYSID_PARK_FILL)
            tmem_park<GRAM_MAXNT>(tpark, acc);
#endif
            __syncthreads();
            PHASE_TICK(clkM)
        }
#else
        const int nsamp = (int)min((long long)FSB, Nlim - base);
        for (int s0 = 0; s0 < nsamp;) {
            // the round: up to FTS samples, as long as their rows (18 - rank J_c each, 0 for a skipped sample) fit the tile
            int off[FTS + 1], cnt = 0;
            off[0] = 0;
#pragma unroll
            for (int u = 0; u < FTS; ++u) {
                int nq = 0;
                bool fits = false;
                if (cnt == u && s0 + u < FSB) {
                    const double* cu = ctx + (s0 + u) * CX_STRIDE;
                    nq = (cu[CX_W] == 0.0) ? 0 : (int)cu[CX_NQ];
                    fits = off[u] + nq <= FROWS;
                }
                if (fits) { off[u + 1] = off[u] + nq; cnt = u + 1; } else off[u + 1] = off[u];
            }
            const int rows = off[FTS], ksteps = (rows + 3) >> 2;
#ifdef SYSID_PHASE_CLOCKS
            clkSub[4] += 1000000LL + ksteps;         // diagnostic: rounds (x 1e6) and k-steps of this CTA
#endif
            for (int e = t; e < (4 * ksteps - rows) * CW; e += GRAM_THREADS) tile[(rows + e / CW) * TILE_LD + (e % CW)] = 0.0;   // pad rows of the last k-step
            phase_wbuild<FTS>(M, ctx, Wsm, s0, cnt, warp, lane);
            __syncthreads();
            PHASE_TICK(clkSub[5])
            phase_proj_mma<FTS, TILE_LD>(M, ctx, Wsm, ident18, tile, s0, cnt, off, args.friction, warp, lane);
            __syncthreads();
            PHASE_TICK(clkC)
            if (ksteps > 0) {
#if defined(SYSID_PARK_FILL)
                tmem_unpark<GRAM_MAXNT>(tpark, acc);
#endif
                mma_dispatch<GRAM_WARPS, GRAM_MAXNT>(warp, tile, ksteps, lane, acc);
#if defined(SYSID_PARK_FILL)
                tmem_park<GRAM_MAXNT>(tpark, acc);
#endif
            }
            __syncthreads();
            PHASE_TICK(clkM)
            s0 += cnt;
        }
#endif
    }
    // flush: this CTA's partial Gram (whole-launch mode: once, after its last super-batch) or the segment's Gram
#if defined(SYSID_PARK_FILL)
    if (!first_sb) tmem_unpark<GRAM_MAXNT>(tpark, acc);
    else {
#pragma unroll
        for (int k = 0; k < GRAM_MAXNT; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }
    }
#endif
    store_dispatch<GRAM_WARPS, GRAM_MAXNT>(warp, partial, lane, acc);
    __syncthreads();                           // s_stat is complete (phase_finish of the last super-batch) and may be reset below
    if (tid == 0) {
        partial[GRAM_NTILES * 64 + 0] = s_stat[0];
        partial[GRAM_NTILES * 64 + 1] = s_stat[1];
        partial[GRAM_NTILES * 64 + 2] = s_stat[2];
        if (segmented) { s_stat[0] = 0.0; s_stat[1] = 0.0; s_stat[2] = 0.0; }
    }
    __syncthreads();
    }   // segments
#ifndef SYSID_PARK_L2
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_tmem, TMEM_PARK_COLS);
#endif
    double* partial = args.partial + (size_t)blockIdx.x * PARTIAL_DOUBLES;
    if (tid == 0 && !segmented) {
#ifdef SYSID_PHASE_CLOCKS
        partial[GRAM_NTILES * 64 + 3] = (double)clkF; partial[GRAM_NTILES * 64 + 4] = (double)clkC; partial[GRAM_NTILES * 64 + 5] = (double)clkM;
        for (int k = 0; k < 6; ++k) partial[GRAM_NTILES * 64 + 6 + k] = (double)clkSub[k];
#endif
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
    double acc[GRAM_MAXNT][2];
#pragma unroll
    for (int t = 0; t < GRAM_MAXNT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
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
        mma_dispatch<16, GRAM_MAXNT>(warp, tile, TILE_ROWS / 4, lane, acc);
        __syncthreads();
    }
    double* partial = args.partial + (size_t)blockIdx.x * PARTIAL_DOUBLES;
    store_dispatch<16, GRAM_MAXNT>(warp, partial, lane, acc);
    if (tid == 0) {
        partial[GRAM_NTILES * 64 + 0] = 0.0; partial[GRAM_NTILES * 64 + 1] = 0.0; partial[GRAM_NTILES * 64 + 2] = 0.0;
    }
}

// Deterministic reduction of the per-CTA partial Grams into stats = [G (c x c) | r (c) | s | n] (ADDS into stats).
// Element (a, b) of the augmented Gram in TILE coordinates lives in tile tri(a/8, b/8) of the lower triangle (both orders exist in
// a diagonal tile); ColMap takes a column of the statistics (0 .. c, c = the torque column) to its tile column -- the identity for
// gram_fused_kernel / gram_stack_kernel, the class layout for gram_struct_kernel.
// gridDim.y > 1 (block bootstrap): statistics block y is the single partial y -> stats + y * stats_stride.
struct ColMap { uint8_t p[CW]; };
constexpr int REDUCE_LANES = 8;          // lanes per output element of gram_reduce_kernel (a power of two <= 32)
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nparts, int c, double n_rows_per_weight,
                                   double n_add_fixed, double* __restrict__ stats, long long* __restrict__ info,
                                   const ColMap cm, long long stats_stride = 0) {
    partial += (size_t)blockIdx.y * nparts * PARTIAL_DOUBLES;
    stats += (size_t)blockIdx.y * stats_stride;
    const int ca = c + 1;
    const int total = ca * (ca + 1) / 2;
    // REDUCE_LANES consecutive lanes share one element: lane `sub` sums the partials sub, sub + REDUCE_LANES, ... and a fixed shuffle tree
    // combines them (one thread per element walked 148 partials one L2 latency after the other: 70 us; the order stays fixed, so the
    // result stays bit-reproducible)
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = gt / REDUCE_LANES, sub = gt % REDUCE_LANES;
    const bool live = e < total;
    double sum = 0.0;
    int i = 0, j = 0;
    if (live) {
        // unrank e -> (i, j), i >= j
        i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        j = e - i * (i + 1) / 2;
        const int pi = cm.p[i], pj = cm.p[j];
        const int a = (pi >> 3) >= (pj >> 3) ? pi : pj, b = (pi >> 3) >= (pj >> 3) ? pj : pi;
        const int ti = a >> 3, tj = b >> 3;
        const int off = (ti * (ti + 1) / 2 + tj) * 64 + (a & 7) * 8 + (b & 7);
        for (int p = sub; p < nparts; p += REDUCE_LANES) sum += partial[(size_t)p * PARTIAL_DOUBLES + off];
    }
#pragma unroll
    for (int o = REDUCE_LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (live && sub == 0) {
        if (i < c) {
            stats[(size_t)i * c + j] += sum;
            if (i != j) stats[(size_t)j * c + i] += sum;
        } else if (j < c) {
            stats[(size_t)c * c + j] += sum;
        } else {
            stats[(size_t)c * c + c] += sum;
        }
    }
    if (gt == 0) {
        double wsum = 0.0, f0 = 0.0, f1 = 0.0;
        for (int p = 0; p < nparts; ++p) {
            wsum += partial[(size_t)p * PARTIAL_DOUBLES + GRAM_NTILES * 64 + 0];
            f0 += partial[(size_t)p * PARTIAL_DOUBLES + GRAM_NTILES * 64 + 1];
            f1 += partial[(size_t)p * PARTIAL_DOUBLES + GRAM_NTILES * 64 + 2];
        }
        stats[(size_t)c * c + c + 1] += n_rows_per_weight * wsum + n_add_fixed;
        if (info) {
            if (gridDim.y > 1) { atomicAdd((unsigned long long*)&info[0], (unsigned long long)f0); atomicAdd((unsigned long long*)&info[1], (unsigned long long)f1); }
            else { info[0] += (long long)f0; info[1] += (long long)f1; }
        }
    }
}

// out (B x E) = Wt (B x K) S (K x E): the statistics of B bootstrap resamples from the per-block statistics S and the
// multiplicities Wt (every statistic is additive over blocks, n included).  A warp owns an 8 x 64 strip of `out` (eight DMMA
// tiles), A and B fragments straight from global memory (S is a few tens of MB: L2-resident).  E must be a multiple of 8.
constexpr int COMBINE_WARPS = 4;
__global__ void __launch_bounds__(32 * COMBINE_WARPS)
combine_stats_kernel(const double* __restrict__ Wt, long long B, long long K, const double* __restrict__ S, long long E,
                     double* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, kk = lane & 3;
    const long long e0 = ((long long)blockIdx.x * COMBINE_WARPS + warp) * 64, b0 = (long long)blockIdx.y * 8;
    if (e0 >= E) return;
    const int ntile = (int)min(8LL, (E - e0) / 8);
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    const bool rowok = b0 + r < B;
    const double* wrow = Wt + (rowok ? b0 + r : 0) * K;
    for (long long k = 0; k < K; k += 4) {
        const bool kok = k + kk < K;
        const double a = (rowok && kok) ? wrow[k + kk] : 0.0;
        const double* srow = S + (kok ? k + kk : 0) * E + e0 + r;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (t < ntile) {
                const double b = kok ? srow[8 * t] : 0.0;
                dmma884(acc[t][0], acc[t][1], a, b);
            }
        }
    }
    if (rowok) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            if (t < ntile) *reinterpret_cast<double2*>(out + (b0 + r) * E + e0 + 8 * t + 2 * kk) = make_double2(acc[t][0], acc[t][1]);
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
// Same F phases and explicit projector as before, but Y phi is formed the way recursive Newton-Euler forms it instead of
// row by row: every body's wrench (f; n) = bodyRegressor(omega, alpha, acc) phi_i in closed form (the transpose of
// body_row: f = m acc + alpha x h + omega x (omega x h), n = h x acc + I alpha + omega x (I omega), with phi_i read in the
// regressor's column order), carried to the base frame (F = R f, N = R n + p x F), summed over the subtree of every joint and
// projected on the joint's Pluecker axis; the six base rows are the total wrench.  ~1.5 k flops per sample instead of the
// 28 k of eighteen projected tile rows, and no tile.  (69 -> see profiles/ for the measured rate.)
// partial per CTA: [sum_i ||e_i||^2, per-joint sum of squares (MAXD), count]
// ------------------------------------------------------------------------------------------------------------
constexpr int RMSE_PARTIAL = MAXD + 2;
constexpr int RMSE_FRONT = RSB * (SC_STRIDE + IN_CHANNELS);          // F-phase scratch + staged inputs; wrenches afterwards
static_assert(RSB * (MAXB * 6 + MAXV) <= RMSE_FRONT, "wrenches and the row vector fit the scratch region");
constexpr size_t RMSE_SMEM_BYTES = sizeof(double) * (RMSE_FRONT + RSB * CX_STRIDE);

struct RmseArgs {
    SampleIO io; long long N; const double* phi; double* partial;
};

__global__ void __launch_bounds__(GRAM_THREADS, 1)
rmse_kernel(const __grid_constant__ DevModel M, const RmseArgs args) {
    extern __shared__ __align__(16) double smem[];
    double* ctx = smem + RMSE_FRONT;
    double* scr = smem;
    double* inp = smem + RSB * SC_STRIDE;
    double* Wb = smem;                         // [RSB][MAXB][6] base-frame wrenches (after the F phases)
    double* G = smem + RSB * MAXB * 6;         // [RSB][MAXV]    Y phi - S^T tau
    __shared__ double s_x[CW];
    __shared__ double s_acc[RMSE_PARTIAL];
    __shared__ int s_bad[RSB];
    const int tid = threadIdx.x, t = tid;
    const int np = M.nparams, nd = M.nd, nb = M.nb, nv = M.nv;
    if (tid < CW) s_x[tid] = (tid < np) ? args.phi[tid] : 0.0;
    if (tid < RMSE_PARTIAL) s_acc[tid] = 0.0;
    if (tid < RSB) s_bad[tid] = 0;
    __shared__ double s_stat[3];
#ifdef SYSID_PHASE_CLOCKS
    long long clk0 = 0, clkSub[6] = {0, 0, 0, 0, 0, 0};
#endif
    const long long nsb = (args.N + RSB - 1) / RSB;
    __syncthreads();
    const long long Nlim = args.N;
    for (long long sb = blockIdx.x; sb < nsb; sb += gridDim.x) {
        const long long base = sb * RSB;
        SYSID_F_PHASES(RSB, GRAM_THREADS, __syncthreads, SYSID_CONTACT_PROJ, )
        if (t < RSB) s_bad[t] = 0;
        // wrench of every body in the base frame: thread per (sample, body)
        for (int it = t; it < RSB * nb; it += GRAM_THREADS) {
            const int s = it / nb, bi = it - s * nb;           // body bi <-> joint bi + 1
            if (base + s >= args.N) continue;
            const double* c = ctx + s * CX_STRIDE;
            const double* b9 = c + CX_B9 + B9S * bi;
            const double* ph = s_x + 10 * bi;
            const double w0 = b9[0], w1 = b9[1], w2 = b9[2], al0 = b9[3], al1 = b9[4], al2 = b9[5], ac0 = b9[6], ac1 = b9[7], ac2 = b9[8];
            const double m = ph[0], h0 = ph[1], h1 = ph[2], h2 = ph[3];
            const double Ixx = ph[4], Ixy = ph[5], Iyy = ph[6], Ixz = ph[7], Iyz = ph[8], Izz = ph[9];     // the regressor's column order
            // f = m acc + alpha x h + omega x (omega x h)
            const double wh0 = w1 * h2 - w2 * h1, wh1 = w2 * h0 - w0 * h2, wh2 = w0 * h1 - w1 * h0;
            const double f0 = m * ac0 + (al1 * h2 - al2 * h1) + (w1 * wh2 - w2 * wh1);
            const double f1 = m * ac1 + (al2 * h0 - al0 * h2) + (w2 * wh0 - w0 * wh2);
            const double f2 = m * ac2 + (al0 * h1 - al1 * h0) + (w0 * wh1 - w1 * wh0);
            // n = h x acc + I alpha + omega x (I omega)
            const double Iw0 = Ixx * w0 + Ixy * w1 + Ixz * w2, Iw1 = Ixy * w0 + Iyy * w1 + Iyz * w2, Iw2 = Ixz * w0 + Iyz * w1 + Izz * w2;
            const double n0 = (h1 * ac2 - h2 * ac1) + (Ixx * al0 + Ixy * al1 + Ixz * al2) + (w1 * Iw2 - w2 * Iw1);
            const double n1 = (h2 * ac0 - h0 * ac2) + (Ixy * al0 + Iyy * al1 + Iyz * al2) + (w2 * Iw0 - w0 * Iw2);
            const double n2 = (h0 * ac1 - h1 * ac0) + (Ixz * al0 + Iyz * al1 + Izz * al2) + (w0 * Iw1 - w1 * Iw0);
            double* out = Wb + (s * MAXB + bi) * 6;
            if (bi == 0) { out[0] = f0; out[1] = f1; out[2] = f2; out[3] = n0; out[4] = n1; out[5] = n2; }
            else {
                const double* X = c + CX_X + 12 * (bi - 1);     // R (row-major) and p of the body's joint relative to the base
                const double F0 = X[0] * f0 + X[1] * f1 + X[2] * f2, F1 = X[3] * f0 + X[4] * f1 + X[5] * f2, F2 = X[6] * f0 + X[7] * f1 + X[8] * f2;
                const double p0 = X[9], p1 = X[10], p2 = X[11];
                out[0] = F0; out[1] = F1; out[2] = F2;
                out[3] = X[0] * n0 + X[1] * n1 + X[2] * n2 + (p1 * F2 - p2 * F1);
                out[4] = X[3] * n0 + X[4] * n1 + X[5] * n2 + (p2 * F0 - p0 * F2);
                out[5] = X[6] * n0 + X[7] * n1 + X[8] * n2 + (p0 * F1 - p1 * F0);
            }
        }
        __syncthreads();
        // (Y phi - S^T tau)[c]: thread per (sample, row)
        for (int it = t; it < RSB * nv; it += GRAM_THREADS) {
            const int s = it / nv, ci = it - s * nv;
            if (base + s >= args.N) continue;
            const double* c = ctx + s * CX_STRIDE;
            const double* ws = Wb + s * MAXB * 6;
            double g;
            if (ci < 6) {
                g = 0.0;
                for (int bi = 0; bi < nb; ++bi) g += ws[bi * 6 + ci];
            } else {
                const int j = ci - 4;                            // joint of the row (row 6 <-> joint 2)
                const unsigned mask = M.submask[j];
                double S[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                for (int bi = 1; bi < nb; ++bi)
                    if ((mask >> (bi + 1)) & 1u) {
#pragma unroll
                        for (int k = 0; k < 6; ++k) S[k] += ws[bi * 6 + k];
                    }
                const double* a = c + CX_A + 6 * (j - 2);
                g = a[0] * S[0] + a[1] * S[1] + a[2] * S[2] + a[3] * S[3] + a[4] * S[4] + a[5] * S[5] - c[CX_TAU + (ci - 6)];
            }
            G[s * MAXV + ci] = g;
        }
        __syncthreads();
        // e = (P g)[6 + k], thread per (sample, joint)
        // squared errors go to the (dead) wrench region; thread k then adds its joint's column sample by sample, so the
        // per-CTA sums have ONE summation order (no atomics: the pass is bit-reproducible, like the Gram)
        for (int it = t; it < RSB * nd; it += GRAM_THREADS) {
            const int s = it / nd, k = it - s * nd, r = 6 + k;
            double e = 0.0;
            if (base + s < args.N) {
                const double* c = ctx + s * CX_STRIDE;
                const double* P = c + CX_P;
                const double* gs = G + s * MAXV;
                for (int cc = 0; cc < nv; ++cc) e = fma(P[pk(r, cc)], gs[cc], e);
                e *= c[CX_W];
            }
            Wb[s * MAXD + k] = e * e;
        }
        __syncthreads();
        if (tid < nd) {
            double a = s_acc[1 + tid];
            for (int s = 0; s < RSB; ++s) a += Wb[s * MAXD + tid];
            s_acc[1 + tid] = a;
        }
        // the next super-batch's `stage` writes only the staged inputs and is followed by a barrier before anything touches Wb
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        for (int k = 0; k < nd; ++k) a += s_acc[1 + k];
        s_acc[0] = a;
    }
    __syncthreads();
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
