// Log ingest on the device: the tab-separated "%.6f" text of the reference's .dat files and the
// vectorised form of its CSV post-processing loops.
//
//   dat_count_kernel / dat_offsets_kernel / dat_parse_kernel   np.loadtxt(path, delimiter='\t', dtype=np.float32)
//                                                              reference spot_identification.py:9-14, demo/solo_identification.py:9-14
//   fd_rate_kernel        the row loop of calculate_low_motor_ddq: delta * 1000 / delta_tick with its three branches
//                         reference g1-data/low_ddq_contact_tick.py:46-70, low_ddq_tick.py:19-33, low_ddq.py:19-33
//   contact_label_kernel  np.where(tau >= 10, 1, np.where(tau > -5, 2, 0))   reference g1-data/low_ddq_contact_tick.py:72-81
//   round_dat_kernel      np.savetxt(fmt='%.6f') followed by np.loadtxt(dtype=np.float32): the value a number has after
//                         its trip through a .dat file   reference g1-data/csv2dat.py:50-55 + spot_identification.py:10-14
//
// All of it is byte / element streaming work: HBM-bound, no tensor cores.  The text is staged through shared memory
// with coalesced word loads; a thread then walks its own 32-byte slice from a padded (conflict-free) layout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sysid {

constexpr int DAT_SLICE = 32;                                   // bytes per thread
constexpr int DAT_THREADS = 256;
constexpr int DAT_BLOCK_BYTES = DAT_SLICE * DAT_THREADS;        // 8 KB of text per block
constexpr int DAT_HALO = 64;                                    // a field that starts in the block may run this far past it
constexpr int DAT_MAXFIELD = DAT_HALO - 1;
// shared layout: byte p of the block window (p = -4 .. DAT_BLOCK_BYTES + DAT_HALO - 1) lives in word (p + 4) / 4; every
// 8 words (one slice) are followed by one pad word, so lanes walking their slices in step touch 32 distinct banks.
constexpr int DAT_WORDS = (DAT_BLOCK_BYTES + DAT_HALO + 4) / 4;
constexpr int DAT_SWORDS = DAT_WORDS + DAT_WORDS / 8 + 1;

struct DatHeader {            // first 64 bytes of the workspace
    long long ndelims;        // tabs + newlines in the text
    long long nlines;         // newlines (+1 if the last byte is not a newline)
    long long nfields;        // ndelims (+1 if the last byte is not a newline)
    long long bad_fields;     // fields that are empty, malformed, too long, or outside the exact-conversion range
    long long first_bad;      // smallest index of such a field (LLONG_MAX if none)
    long long ragged;         // rows whose field count differs from cols
    long long reserved[2];
};

__device__ __forceinline__ int dat_sidx(int w) { return w + (w >> 3); }

// Stage the block's window into shared memory.  Bytes outside [0, nbytes) read as '\n' (a delimiter).
__device__ __forceinline__ void dat_stage(const unsigned char* __restrict__ text, long long nbytes, long long b0, unsigned* sm, int tid) {
    // word w covers bytes b0 - 4 + 4 w .. + 3 (the text pointer is 4-byte aligned, b0 a multiple of 8192): one coalesced
    // 4-byte load per thread per pass
    for (int w = tid; w < DAT_WORDS; w += DAT_THREADS) {
        const long long p = b0 - 4 + 4LL * w;
        unsigned v;
        if (p >= 0 && p + 4 <= nbytes) v = *reinterpret_cast<const unsigned*>(text + p);
        else {
            v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long pk = p + k;
                const unsigned c = (pk >= 0 && pk < nbytes) ? text[pk] : (unsigned)'\n';
                v |= c << (8 * k);
            }
        }
        sm[dat_sidx(w)] = v;
    }
}

__device__ __forceinline__ unsigned dat_byte(const unsigned* sm, int p) {      // p relative to the block start, >= -4
    const int q = p + 4;
    return (sm[dat_sidx(q >> 2)] >> (8 * (q & 3))) & 0xffu;
}

__device__ __forceinline__ int dat_slice_delims(const unsigned* sm, int tid, unsigned delim, int* newlines) {
    int n = 0, nl = 0;
#pragma unroll
    for (int w = 0; w < DAT_SLICE / 4; ++w) {
        const unsigned v = sm[dat_sidx(1 + tid * (DAT_SLICE / 4) + w)];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned c = (v >> (8 * k)) & 0xffu;
            nl += (c == '\n');
            n += (c == '\n') | (c == delim);
        }
    }
    *newlines = nl;
    return n;
}

// pass 1: delimiters and newlines per block
__global__ void __launch_bounds__(DAT_THREADS)
dat_count_kernel(const unsigned char* __restrict__ text, long long nbytes, unsigned delim, long long* __restrict__ blk, DatHeader* hdr) {
    __shared__ unsigned sm[DAT_SWORDS];
    __shared__ int s_n[DAT_THREADS / 32], s_l[DAT_THREADS / 32];
    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * DAT_BLOCK_BYTES;
    dat_stage(text, nbytes, b0, sm, tid);
    __syncthreads();
    int nl, n = dat_slice_delims(sm, tid, delim, &nl);
    // bytes past the end of the text were staged as '\n': do not count them
    const long long s0 = b0 + (long long)tid * DAT_SLICE;
    if (s0 + DAT_SLICE > nbytes) {
        n = 0; nl = 0;
        for (long long p = s0; p < nbytes; ++p) { const unsigned c = text[p]; nl += (c == '\n'); n += (c == '\n') | (c == delim); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); nl += __shfl_xor_sync(0xffffffffu, nl, o); }
    if ((tid & 31) == 0) { s_n[tid >> 5] = n; s_l[tid >> 5] = nl; }
    __syncthreads();
    if (tid == 0) {
        int tn = 0, tl = 0;
        for (int k = 0; k < DAT_THREADS / 32; ++k) { tn += s_n[k]; tl += s_l[k]; }
        blk[blockIdx.x] = tn;
        atomicAdd(reinterpret_cast<unsigned long long*>(&hdr->nlines), (unsigned long long)tl);     // integer: order-independent
    }
}

// pass 2: exclusive scan of the block counts (one block; the list is short: nbytes / 8192 entries)
__global__ void __launch_bounds__(1024)
dat_offsets_kernel(long long* __restrict__ blk, long long nblk, const unsigned char* __restrict__ text, long long nbytes, DatHeader* hdr) {
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (long long base = 0; base < nblk; base += 1024) {
        const long long i = base + tid;
        const long long v = (i < nblk) ? blk[i] : 0;
        long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_w[lane] = w;
        }
        __syncthreads();
        const long long excl = s_carry + (warp ? s_w[warp - 1] : 0) + x - v;
        if (i < nblk) blk[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        const long long open = (nbytes > 0 && text[nbytes - 1] != '\n') ? 1 : 0;     // last line without a newline
        hdr->ndelims = s_carry;
        hdr->nfields = s_carry + open;
        hdr->nlines += open;
        hdr->bad_fields = 0; hdr->first_bad = 0x7fffffffffffffffLL; hdr->ragged = 0;
    }
}

__device__ __constant__ double DAT_P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                              1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// M / D correctly rounded (ties to even) for 64-bit integers: long digit strings ("%.6f" above 9.007e9, 17-digit CSV
// fields) are M / 10^k = (M / 5^k) 2^-k with 5^k < 2^63 for k <= 27; the power of two is exact.  N = M 2^s with s chosen so that the integer quotient has at least 56 bits; the remainder is
// the sticky bit.
__device__ __constant__ unsigned long long DAT_U5[28] = {1ULL, 5ULL, 25ULL, 125ULL, 625ULL, 3125ULL, 15625ULL, 78125ULL, 390625ULL, 1953125ULL, 9765625ULL, 48828125ULL, 244140625ULL, 1220703125ULL, 6103515625ULL, 30517578125ULL, 152587890625ULL, 762939453125ULL, 3814697265625ULL, 19073486328125ULL, 95367431640625ULL, 476837158203125ULL, 2384185791015625ULL, 11920928955078125ULL, 59604644775390625ULL, 298023223876953125ULL, 1490116119384765625ULL, 7450580596923828125ULL};

__device__ double dat_ratio(unsigned long long M, unsigned long long D) {
    const int bm = 64 - __clzll((long long)M), bd = 64 - __clzll((long long)D);
    int s = 56 + bd - bm;
    if (s < 0) s = 0;
    const unsigned __int128 N = (unsigned __int128)M << s;
    const unsigned long long Q = (unsigned long long)(N / D);
    const bool sticky = (N % D) != 0;
    const int nb = 64 - __clzll((long long)Q), r = nb - 53;        // nb >= 56
    unsigned long long mant = Q >> r;
    const unsigned long long rem = Q & ((1ULL << r) - 1ULL), half = 1ULL << (r - 1);
    if (rem > half || (rem == half && (sticky || (mant & 1ULL)))) ++mant;
    return scalbn((double)mant, r - s);
}

// One decimal field -> double, exactly as strtod would round it.  Clinger's exact path: the digits form an integer
// M <= 2^53 and the decimal exponent e satisfies |e| <= 22, so M and 10^|e| are both exact doubles and ONE correctly
// rounded IEEE multiplication or division gives the correctly rounded value ("%.6f" text below 9.007e9).  Longer digit
// strings and smaller exponents (M < 2^64, -27 <= e <= 0) go through the integer division of dat_ratio.  Anything else (20+ significant
// digits, large exponents, garbage) returns false: counted and reported, never approximated.
// (the field is read in place from the staged text through an accessor: s[i] = byte i of the field)
struct DatSmemField {
    const unsigned* sm; int p0;
    __device__ __forceinline__ unsigned operator[](int i) const { return dat_byte(sm, p0 + i); }
};

template <typename S>
__device__ __forceinline__ bool dat_lower_eq(const S& s, int a, int n, const char* word, int wl) {
    if (n != wl) return false;
    for (int k = 0; k < wl; ++k) if ((s[a + k] | 0x20u) != (unsigned)(unsigned char)word[k]) return false;
    return true;
}

template <typename S>
__device__ bool dat_convert(const S& s, int n, double* out) {
    int a = 0, b = n;
    while (a < b && (s[a] == ' ' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\r')) --b;
    if (a == b) return false;
    bool neg = false;
    if (s[a] == '+' || s[a] == '-') { neg = (s[a] == '-'); ++a; }
    if (a == b) return false;
    if ((s[a] | 0x20) == 'n' || (s[a] | 0x20) == 'i') {
        if (dat_lower_eq(s, a, b - a, "nan", 3)) { *out = neg ? -__longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double(0x7ff8000000000000LL); return true; }
        if (dat_lower_eq(s, a, b - a, "inf", 3) || dat_lower_eq(s, a, b - a, "infinity", 8)) {
            *out = neg ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL); return true;
        }
        return false;
    }
    unsigned long long mant = 0;
    int e10 = 0, ndig = 0;
    bool any = false, dot = false;
    for (; a < b; ++a) {
        const unsigned c = s[a];
        if (c >= '0' && c <= '9') {
            any = true;
            if (mant < 1000000000000000000ULL) { mant = mant * 10 + (c - '0'); if (dot) --e10; ndig += (mant != 0); }
            else { if (c != '0') return false; if (!dot) ++e10; }       // a 20th significant digit: only zeros are exact
        } else if (c == '.' && !dot) dot = true;
        else break;
    }
    if (!any) return false;
    if (a < b) {
        if ((s[a] | 0x20) != 'e') return false;
        ++a;
        bool eneg = false;
        if (a < b && (s[a] == '+' || s[a] == '-')) { eneg = (s[a] == '-'); ++a; }
        if (a == b) return false;
        int ev = 0;
        for (; a < b; ++a) {
            const unsigned c = s[a];
            if (c < '0' || c > '9') return false;
            if (ev < 100000) ev = ev * 10 + (int)(c - '0');
        }
        e10 += eneg ? -ev : ev;
    }
    double v;
    if (mant == 0) v = 0.0;
    else if (mant <= (1ULL << 53) && e10 >= -22 && e10 <= 22) {
        const double m = (double)mant;                         // exact
        if (e10 == 0) v = m;
        else if (e10 < 0) v = __ddiv_rn(m, DAT_P10[-e10]);     // both exact: one rounding
        else v = __dmul_rn(m, DAT_P10[e10]);
    } else if (e10 == 0) v = __ull2double_rn(mant);            // integer -> double is correctly rounded
    else if (e10 < 0 && e10 >= -27) v = scalbn(dat_ratio(mant, DAT_U5[-e10]), e10);
    else return false;
    *out = neg ? -v : v;
    return true;
}

struct DatParseArgs {
    const unsigned char* text; long long nbytes; unsigned delim;
    const long long* blkoff; DatHeader* hdr;
    long long rows, cols;
    double* out; long long ld;
    int round_f32;
    int transpose;       // out is cols x rows (element (row, col) at out[col * ld + row]): a row-major CSV lands channel-major
    int empty_nan;       // an empty field is NaN (pandas.read_csv) instead of an error (np.loadtxt)
};

// bit p of the result: byte p of the thread's slice is a delimiter (the field delimiter or a newline)
__device__ __forceinline__ unsigned dat_slice_delim_mask(const unsigned* sm, int tid, unsigned delim) {
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < DAT_SLICE / 4; ++w) {
        const unsigned v = sm[dat_sidx(1 + tid * (DAT_SLICE / 4) + w)];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned c = (v >> (8 * k)) & 0xffu;
            m |= (unsigned)((c == '\n') | (c == delim)) << (4 * w + k);
        }
    }
    return m;
}

// pass 3: every thread converts the fields that START in its 32-byte slice.  The starts are collected in a bit mask
// first and the fields are then converted one ordinal at a time, so the lanes of a warp run the conversion loop together
// (their j-th fields) instead of one after the other at 32 different byte offsets.
__global__ void __launch_bounds__(DAT_THREADS)
dat_parse_kernel(const DatParseArgs g) {
    __shared__ unsigned sm[DAT_SWORDS];
    __shared__ int s_w[DAT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long b0 = (long long)blockIdx.x * DAT_BLOCK_BYTES;
    dat_stage(g.text, g.nbytes, b0, sm, tid);
    __syncthreads();
    const unsigned dmask = dat_slice_delim_mask(sm, tid, g.delim);
    const int n = __popc(dmask);
    // exclusive scan of the slice counts over the block
    int x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    int wbase = 0;
    for (int k = 0; k < warp; ++k) wbase += s_w[k];
    const long long idx0 = g.blkoff[blockIdx.x] + wbase + x - n;  // index of the field the slice's first byte belongs to
    const int p0 = tid * DAT_SLICE;
    // a field starts at the first byte of the text and after every delimiter; bytes past the end of the text start nothing
    const unsigned prev = dat_byte(sm, p0 - 1);
    const unsigned first = (b0 + p0 == 0 || prev == '\n' || prev == g.delim) ? 1u : 0u;
    unsigned smask = (dmask << 1) | first;
    const long long left = g.nbytes - (b0 + p0);
    if (left < DAT_SLICE) smask &= (left <= 0) ? 0u : ((1u << (int)left) - 1u);
    long long nbad = 0, firstbad = 0x7fffffffffffffffLL, nragged = 0;
    while (smask) {
        const int o = __ffs(smask) - 1;
        smask &= smask - 1;
        const int p = p0 + o;
        const long long idx = idx0 + __popc(dmask & ((1u << o) - 1u));
        // the field ends at the next delimiter or at the end of the text (it may run into the halo)
        int len = 0;
        bool closed = false;
        unsigned term = '\n';
        for (int e = p; e < DAT_BLOCK_BYTES + DAT_HALO; ++e) {
            const unsigned ce = dat_byte(sm, e);
            if (ce == '\n' || ce == g.delim || b0 + e >= g.nbytes) { closed = true; term = (b0 + e >= g.nbytes) ? (unsigned)'\n' : ce; break; }
            ++len;
        }
        double v = __longlong_as_double(0x7ff8000000000000LL);
        const bool ok = closed && len <= DAT_MAXFIELD && ((len == 0 && g.empty_nan) || dat_convert(DatSmemField{sm, p}, len, &v));
        const long long row = idx / g.cols, col = idx - row * g.cols;
        if (!ok) { ++nbad; if (idx < firstbad) firstbad = idx; v = __longlong_as_double(0x7ff8000000000000LL); }
        if (row < g.rows) g.out[g.transpose ? col * g.ld + row : row * g.ld + col] = g.round_f32 ? (double)__double2float_rn(v) : v;
        // a row ends exactly where a newline is: the terminator of the last column, and of no other
        if (closed && ((term == '\n') != (col == g.cols - 1))) ++nragged;
    }
    // block totals -> header (integer atomics: order-independent)
    if (nbad | nragged) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&g.hdr->bad_fields), (unsigned long long)nbad);
        atomicAdd(reinterpret_cast<unsigned long long*>(&g.hdr->ragged), (unsigned long long)nragged);
        if (nbad) atomicMin(&g.hdr->first_bad, firstbad);
    }
}

// ------------------------------------------------------------------------------------------------ CSV post-processing
// y[ch][0] = NaN;  for i >= 1, with dt = tick[i] - tick[i-1] and dx = x[ch][i] - x[ch][i-1] (the reference's branches, in order):
//   dt > 0 -> dx * scale / dt  (Python evaluates left to right: (dx * scale) / dt);  dx == 0 -> 0;  otherwise NaN.
struct FdArgs { const double* tick; const double* x; double* y; long long N, ldx, ldy; int channels; double scale; };

__global__ void fd_rate_kernel(const FdArgs g) {
    const long long total = (long long)g.channels * g.N;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ch = e / g.N, i = e - ch * g.N;
        double r = nan;
        if (i > 0) {
            const double dt = __dsub_rn(g.tick[i], g.tick[i - 1]);
            const double dx = __dsub_rn(g.x[ch * g.ldx + i], g.x[ch * g.ldx + i - 1]);
            if (dt > 0.0) r = __ddiv_rn(__dmul_rn(dx, g.scale), dt);
            else if (dx == 0.0) r = 0.0;
        }
        g.y[ch * g.ldy + i] = r;
    }
}

// out = tau >= hi ? 1 : (tau > lo ? 2 : 0)   (NaN compares false twice -> 0, as np.where does)
__global__ void contact_label_kernel(const double* __restrict__ tau, double* __restrict__ out, long long N, double hi, double lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double t = tau[i];
        out[i] = (t >= hi) ? 1.0 : ((t > lo) ? 2.0 : 0.0);
    }
}

// The value x has after np.savetxt(fmt='%.6f') and np.loadtxt (optionally dtype=float32), without the text:
// printf rounds the EXACT binary value to 6 decimals, ties to even; reading the digits back is M / 10^6 correctly rounded.
// x 10^6 = p + e exactly (p the rounded product, e its FMA residual), so the nearest integer of the exact product and the
// direction of a near-tie are both decided without error.  |x| >= 2^33: ulp(x) > 10^-6, the text reads back as x itself.
__device__ __forceinline__ double round_dat_value(double x, int to_f32) {
    double r = x;
    if (fabs(x) < 8589934592.0) {            // also false for NaN / inf, which print as nan / inf and read back unchanged
        const double p = __dmul_rn(x, 1e6);
        const double e = __fma_rn(x, 1e6, -p);
        double y = rint(p);                   // ties to even
        const double d = __dsub_rn(p, y);     // exact
        if (d == 0.5 && e > 0.0) y += 1.0;    // p sits on a tie but the exact product is above it
        else if (d == -0.5 && e < 0.0) y -= 1.0;   // ... below it (every other case: rint's answer stands)
        r = __ddiv_rn(y, 1e6);
        if (y == 0.0) r = copysign(0.0, x);   // "-0.000000" reads back as -0.0
    }
    return to_f32 ? (double)__double2float_rn(r) : r;
}

struct RoundArgs { const double* x; double* y; long long N, ldx, ldy; int channels; int to_f32; };

__global__ void round_dat_kernel(const RoundArgs g) {
    const long long total = (long long)g.channels * g.N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ch = e / g.N, i = e - ch * g.N;
        g.y[ch * g.ldy + i] = round_dat_value(g.x[ch * g.ldx + i], g.to_f32);
    }
}

}  // namespace sysid
