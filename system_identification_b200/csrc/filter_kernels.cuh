// Pre-processing step immediately before the hot path: the reference's read_data filters the
// logged joint velocities / accelerations / torques along the time axis with scipy
// (reference demo/solo_identification.py:15-32):
//     signal.filtfilt(b, a, x, axis=1)           zero-phase IIR, padtype 'odd', padlen 3 max(len a, len b), method 'pad'
//     savgol_filter(x, window_length, polyorder)  deriv 0, mode 'interp'
// Both are restated here for channel-major device arrays (channels x N, leading dimension ld), fp64.
//
// filtfilt: scipy's lfilter is the direct-form-II-transposed recursion, sequential in time.  It is parallelised over
// (channel, chunk of FILT_CHUNK samples) in three kernels per direction:
//   A  every chunk runs the recursion from a ZERO state and keeps only its final state f_k
//   S  one thread per channel chains the chunks:  s_{k+1} = Phi^FILT_CHUNK s_k + f_k,  s_0 = zi * x_ext[0]
//      (Phi = the recursion's zero-input state transition; its power is formed on the host)
//   B  every chunk re-runs the recursion from its true initial state s_k and writes the outputs
// Inside a chunk the arithmetic is exactly lfilter's; only the chunk-initial states carry ~1e-16 relative rounding of
// their own.  The odd extension is generated on the fly; the backward direction reads the forward result reversed and
// writes its output reversed and cropped.
#pragma once
#include <cuda_runtime.h>

namespace sysid {

constexpr int FILT_MAXS = 8;          // state dimension = max(len a, len b) - 1 (coefficients zero-padded: same arithmetic)
constexpr int FILT_CHUNK = 256;

struct FiltCoef {
    double b[FILT_MAXS + 1];
    double a[FILT_MAXS + 1];          // a[0] == 1 (normalised on the host)
    double zi[FILT_MAXS];
    double phiL[FILT_MAXS][FILT_MAXS];   // Phi^FILT_CHUNK
    double phiM[FILT_MAXS][FILT_MAXS];   // Phi^(FILT_CHUNK * seg): one segment of the chaining kernel
};

struct FiltArgs {
    const double* x;        // forward: the signal (channels x N, ld);  backward: the forward result (channels x Next, ld = Next)
    double* y;              // forward: channels x Next (ld = Next);     backward: the final output (channels x N, ld)
    double* fstate;         // [channels][nchunks][FILT_MAXS] chunk-final states from zero state
    double* sstate;         // [channels][nchunks][FILT_MAXS] true chunk-initial states
    long long N, ld, Next;
    int channels, padlen, nchunks, backward;
    int seg;                // chunks per thread of the chaining kernel: ceil(nchunks / FILT_SCAN_THREADS)
    int pad_float32;        // the log was float32 (np.loadtxt(dtype=float32), reference quirk Q8): scipy then forms the odd
                            // extension in float32 before lfilter widens it -- reproduce that rounding of the pad samples
};

// sample i (0 <= i < Next) of the stream this direction filters
__device__ __forceinline__ double filt_input(const FiltArgs& g, int ch, long long i) {
    if (g.backward) return g.x[(size_t)ch * g.Next + (g.Next - 1 - i)];
    const double* x = g.x + (size_t)ch * g.ld;
    const long long j = i - g.padlen;
    if (j >= 0 && j < g.N) return x[j];
    const double e = (j < 0) ? 2.0 * x[0] - x[-j]                   // odd extension, left:  2 x[0] - x[padlen - i]
                             : 2.0 * x[g.N - 1] - x[2 * (g.N - 1) - j];   // right: 2 x[N-1] - x[N-2-(j-N)]
    return g.pad_float32 ? (double)__double2float_rn(e) : e;
}

__device__ __forceinline__ double df2t_step(const FiltCoef& c, double z[FILT_MAXS], double xv) {
    const double yv = c.b[0] * xv + z[0];
#pragma unroll
    for (int k = 0; k < FILT_MAXS - 1; ++k) z[k] = c.b[k + 1] * xv + z[k + 1] - c.a[k + 1] * yv;
    z[FILT_MAXS - 1] = c.b[FILT_MAXS] * xv - c.a[FILT_MAXS] * yv;
    return yv;
}

// MODE 0 = kernel A (zero state -> final state), MODE 1 = kernel B (true state -> outputs)
// A thread owns one (channel, chunk) and its recursion is sequential, so neighbouring lanes are FILT_CHUNK samples (2 KB)
// apart: read directly, every warp load touches 32 different cache lines and the 64 resident warps of an SM thrash L1
// (9.0 ms for 48 channels x 1 M samples).  The block therefore moves its 128 chunks through shared memory in sub-tiles of
// 32 samples: a warp loads 32 CONSECUTIVE samples of one chunk per instruction (coalesced, also for the reversed and
// odd-extended streams: the index map is applied per element), the owners run 32 steps out of a pitch-33 tile, and
// kernel B writes the outputs back the same way.
constexpr int FILT_THREADS = 128;
constexpr int FILT_SUB = 32;
constexpr int FILT_PITCH = FILT_SUB + 1;
static_assert(FILT_CHUNK % FILT_SUB == 0, "whole sub-tiles");

template <int MODE>
__global__ void __launch_bounds__(FILT_THREADS) filt_chunk_kernel(const __grid_constant__ FiltCoef c, const FiltArgs g) {
    __shared__ double tile[FILT_THREADS * FILT_PITCH];
    __shared__ int s_ch[FILT_THREADS];
    __shared__ long long s_i0[FILT_THREADS];
    const int tid = threadIdx.x;
    const long long total = (long long)g.channels * g.nchunks;
    const long long t = (long long)blockIdx.x * FILT_THREADS + tid;
    const bool live = t < total;
    const int ch = live ? (int)(t / g.nchunks) : -1, k = live ? (int)(t - (long long)ch * g.nchunks) : 0;
    const long long i0 = (long long)k * FILT_CHUNK;
    const long long i1 = (i0 + FILT_CHUNK < g.Next) ? i0 + FILT_CHUNK : g.Next;
    s_ch[tid] = ch; s_i0[tid] = i0;
    double z[FILT_MAXS];
    double* st = live ? ((MODE == 0) ? g.fstate : g.sstate) + ((size_t)ch * g.nchunks + k) * FILT_MAXS : nullptr;
#pragma unroll
    for (int e = 0; e < FILT_MAXS; ++e) z[e] = (MODE == 0 || !live) ? 0.0 : st[e];
    __syncthreads();
    for (int sub = 0; sub < FILT_CHUNK / FILT_SUB; ++sub) {
        // cooperative, coalesced load of samples [32 sub, 32 sub + 32) of every chunk of the block
        for (int e = tid; e < FILT_THREADS * FILT_SUB; e += FILT_THREADS) {
            const int r = e / FILT_SUB, j = e - r * FILT_SUB;
            const int chr = s_ch[r];
            const long long i = s_i0[r] + sub * FILT_SUB + j;
            if (chr >= 0 && i < g.Next) tile[r * FILT_PITCH + j] = filt_input(g, chr, i);
        }
        __syncthreads();
        if (live) {
            const long long ib = i0 + sub * FILT_SUB;
#pragma unroll 4
            for (int j = 0; j < FILT_SUB; ++j) {
                if (ib + j < i1) {
                    const double yv = df2t_step(c, z, tile[tid * FILT_PITCH + j]);
                    if (MODE == 1) tile[tid * FILT_PITCH + j] = yv;
                }
            }
        }
        if (MODE == 1) {
            __syncthreads();
            for (int e = tid; e < FILT_THREADS * FILT_SUB; e += FILT_THREADS) {
                const int r = e / FILT_SUB, j = e - r * FILT_SUB;
                const int chr = s_ch[r];
                const long long i = s_i0[r] + sub * FILT_SUB + j;
                if (chr >= 0 && i < g.Next && i < s_i0[r] + FILT_CHUNK) {
                    const double yv = tile[r * FILT_PITCH + j];
                    if (!g.backward) g.y[(size_t)chr * g.Next + i] = yv;
                    else {
                        const long long jj = g.Next - 1 - i - g.padlen;       // un-reverse and crop the extension
                        if (jj >= 0 && jj < g.N) g.y[(size_t)chr * g.ld + jj] = yv;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (MODE == 0 && live) {
#pragma unroll
        for (int e = 0; e < FILT_MAXS; ++e) st[e] = z[e];
    }
}

// Kernel S: one block per channel chains the chunks, s_{k+1} = Phi^L s_k + f_k, as a two-level scan of these affine maps:
// every thread folds its segment of `seg` consecutive chunks from a zero state, thread 0 chains the FILT_SCAN_THREADS
// segment results with Phi^(L seg), and every thread re-runs its segment from its true initial state, writing s_k.
// (History of this kernel, 48 channels x 3 907 chunks: one thread per channel reading f_k from global memory, an L2 round
// trip per chunk: ~2 ms per direction; the same chain out of shared memory: 0.51 ms; two levels: see DESIGN 4.4.)
constexpr int FILT_SCAN_THREADS = 256;

__device__ __forceinline__ void filt_affine(const double (&phi)[FILT_MAXS][FILT_MAXS], double (&s)[FILT_MAXS], const double* __restrict__ f) {
    double nx[FILT_MAXS];
#pragma unroll
    for (int e = 0; e < FILT_MAXS; ++e) nx[e] = f[e];
#pragma unroll
    for (int e = 0; e < FILT_MAXS; ++e)
#pragma unroll
        for (int q = 0; q < FILT_MAXS; ++q) nx[e] = fma(phi[e][q], s[q], nx[e]);
#pragma unroll
    for (int e = 0; e < FILT_MAXS; ++e) s[e] = nx[e];
}

__global__ void __launch_bounds__(FILT_SCAN_THREADS) filt_scan_kernel(const __grid_constant__ FiltCoef c, const FiltArgs g) {
    __shared__ double segs[FILT_SCAN_THREADS][FILT_MAXS + 1];
    const int ch = blockIdx.x, tid = threadIdx.x;
    if (ch >= g.channels) return;
    const double* fs = g.fstate + (size_t)ch * g.nchunks * FILT_MAXS;
    double* ss = g.sstate + (size_t)ch * g.nchunks * FILT_MAXS;
    const int k0 = tid * g.seg, k1 = (k0 + g.seg < g.nchunks) ? k0 + g.seg : g.nchunks;
    double s[FILT_MAXS];
    // level 1: the segment from a zero state
#pragma unroll
    for (int e = 0; e < FILT_MAXS; ++e) s[e] = 0.0;
    for (int k = k0; k < k1; ++k) filt_affine(c.phiL, s, fs + (size_t)k * FILT_MAXS);
#pragma unroll
    for (int e = 0; e < FILT_MAXS; ++e) segs[tid][e] = s[e];
    __syncthreads();
    // level 2: chain the segments; segs[j] becomes the true state at the start of segment j
    if (tid == 0) {
        const double x0 = filt_input(g, ch, 0);
#pragma unroll
        for (int e = 0; e < FILT_MAXS; ++e) s[e] = c.zi[e] * x0;
        for (int j = 0; j < FILT_SCAN_THREADS; ++j) {
            double f[FILT_MAXS];
#pragma unroll
            for (int e = 0; e < FILT_MAXS; ++e) { f[e] = segs[j][e]; segs[j][e] = s[e]; }
            filt_affine(c.phiM, s, f);
        }
    }
    __syncthreads();
    // level 3: the segment again, from its true initial state
#pragma unroll
    for (int e = 0; e < FILT_MAXS; ++e) s[e] = segs[tid][e];
    for (int k = k0; k < k1; ++k) {
#pragma unroll
        for (int e = 0; e < FILT_MAXS; ++e) ss[(size_t)k * FILT_MAXS + e] = s[e];
        filt_affine(c.phiL, s, fs + (size_t)k * FILT_MAXS);
    }
}

// ---------------------------------------------------------------------------------------------- Savitzky-Golay
constexpr int SG_MAXW = 63;            // window_length limit (odd)
struct SavgolArgs {
    const double* x; double* y;
    const double* coef;                // [W] interior FIR taps, then [half][W] left-edge rows, then [half][W] right-edge rows
    long long N, ld;
    int channels, W;
};

__global__ void __launch_bounds__(256) savgol_kernel(const SavgolArgs g) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)g.channels * g.N) return;
    const int ch = (int)(t / g.N);
    const long long i = t - (long long)ch * g.N;
    const double* x = g.x + (size_t)ch * g.ld;
    const int W = g.W, half = W / 2;
    double s = 0.0;
    if (i < half) {                                   // polynomial fitted to the first W samples, evaluated at i
        const double* row = g.coef + W + (size_t)i * W;
        for (int k = 0; k < W; ++k) s = fma(row[k], x[k], s);
    } else if (i >= g.N - half) {                     // ... to the last W samples
        const double* row = g.coef + W + (size_t)half * W + (size_t)(i - (g.N - half)) * W;
        for (int k = 0; k < W; ++k) s = fma(row[k], x[g.N - W + k], s);
    } else {
        for (int k = 0; k < W; ++k) s = fma(g.coef[k], x[i - half + k], s);
    }
    g.y[(size_t)ch * g.ld + i] = s;
}

}  // namespace sysid
