// C-ABI entry points of the log ingest: see include/sysid_b200.h and ingest_kernels.cuh.
#include <cstdarg>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../include/sysid_b200.h"
#include "ingest_kernels.cuh"

namespace sysid { int set_error(int code, const char* message); }      // sysid_api.cu (thread-local message)

using namespace sysid;

namespace {

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return set_error(code, buf);
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) return fail(SYSID_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

long long dat_blocks(int64_t nbytes) { return (nbytes + DAT_BLOCK_BYTES - 1) / DAT_BLOCK_BYTES; }

unsigned stream_grid(long long items) {
    long long b = (items + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148LL * 16) b = 148LL * 16;           // grid-stride: 16 blocks of 256 per SM keep the loads in flight
    return (unsigned)b;
}

}  // namespace

extern "C" {

size_t sysid_dat_workspace_bytes(int64_t nbytes) {
    if (nbytes < 0) return 0;
    return sizeof(DatHeader) + sizeof(long long) * (size_t)(dat_blocks(nbytes) + 1);
}

int sysid_dat_scan(const void* text, int64_t nbytes, int32_t delimiter, void* workspace, size_t workspace_bytes,
                   int64_t* dims_host, void* stream) {
    if (!text || !workspace || !dims_host) return fail(SYSID_ERR_INVALID, "null argument");
    if (nbytes <= 0) return fail(SYSID_ERR_INVALID, "empty text");
    if (((uintptr_t)text & 3u) != 0) return fail(SYSID_ERR_INVALID, "text must be 4-byte aligned");
    if (delimiter <= 0 || delimiter > 127 || delimiter == '\n') return fail(SYSID_ERR_INVALID, "bad delimiter");
    if (workspace_bytes < sysid_dat_workspace_bytes(nbytes)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    DatHeader* hdr = (DatHeader*)workspace;
    long long* blk = (long long*)((char*)workspace + sizeof(DatHeader));
    const long long nblk = dat_blocks(nbytes);
    CUDA_TRY(cudaMemsetAsync(hdr, 0, sizeof(DatHeader), st));
    dat_count_kernel<<<(unsigned)nblk, DAT_THREADS, 0, st>>>((const unsigned char*)text, nbytes, (unsigned)delimiter, blk, hdr);
    CUDA_TRY(cudaGetLastError());
    dat_offsets_kernel<<<1, 1024, 0, st>>>(blk, nblk, (const unsigned char*)text, nbytes, hdr);
    CUDA_TRY(cudaGetLastError());
    DatHeader h;
    CUDA_TRY(cudaMemcpyAsync(&h, hdr, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));          // the caller sizes the output from the answer
    dims_host[0] = h.nlines;
    dims_host[1] = (h.nlines > 0) ? h.nfields / h.nlines : 0;
    if (h.nlines <= 0 || h.nfields % h.nlines != 0)
        return fail(SYSID_ERR_INVALID, "the number of columns changed between rows (%lld fields in %lld rows)", h.nfields, h.nlines);
    return SYSID_OK;
}

int sysid_dat_parse(const void* text, int64_t nbytes, int32_t delimiter, void* workspace, size_t workspace_bytes,
                    int64_t rows, int64_t cols, double* out, int64_t ld, int32_t round_float32, int64_t* info_host, void* stream) {
    return sysid_dat_parse_ex(text, nbytes, delimiter, workspace, workspace_bytes, rows, cols, out, ld,
                              round_float32 ? SYSID_DAT_ROUND_FLOAT32 : 0, info_host, stream);
}

int sysid_dat_parse_ex(const void* text, int64_t nbytes, int32_t delimiter, void* workspace, size_t workspace_bytes,
                       int64_t rows, int64_t cols, double* out, int64_t ld, int32_t flags, int64_t* info_host, void* stream) {
    const int32_t round_float32 = flags & SYSID_DAT_ROUND_FLOAT32;
    const bool transpose = (flags & SYSID_DAT_TRANSPOSE) != 0;
    if (!text || !workspace || !out) return fail(SYSID_ERR_INVALID, "null argument");
    if (nbytes <= 0 || rows <= 0 || cols <= 0 || ld < (transpose ? rows : cols)) return fail(SYSID_ERR_INVALID, "bad nbytes/rows/cols/ld");
    if (workspace_bytes < sysid_dat_workspace_bytes(nbytes)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    DatHeader* hdr = (DatHeader*)workspace;
    DatParseArgs g;
    g.text = (const unsigned char*)text; g.nbytes = nbytes; g.delim = (unsigned)delimiter;
    g.blkoff = (const long long*)((char*)workspace + sizeof(DatHeader)); g.hdr = hdr;
    g.rows = rows; g.cols = cols; g.out = out; g.ld = ld; g.round_f32 = round_float32 ? 1 : 0;
    g.transpose = transpose ? 1 : 0; g.empty_nan = (flags & SYSID_DAT_EMPTY_IS_NAN) ? 1 : 0;
    dat_parse_kernel<<<(unsigned)dat_blocks(nbytes), DAT_THREADS, 0, st>>>(g);
    CUDA_TRY(cudaGetLastError());
    if (info_host) {
        DatHeader h;
        CUDA_TRY(cudaMemcpyAsync(&h, hdr, sizeof(h), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        info_host[0] = h.bad_fields; info_host[1] = h.first_bad; info_host[2] = h.ragged; info_host[3] = h.nfields;
        if (h.nfields != rows * cols) return fail(SYSID_ERR_INVALID, "text holds %lld fields, expected %lld x %lld", h.nfields, (long long)rows, (long long)cols);
        if (h.ragged) return fail(SYSID_ERR_INVALID, "the number of columns changed between rows (%lld row ends out of place)", h.ragged);
        if (h.bad_fields) {
            const long long r = h.first_bad / cols, c = h.first_bad % cols;
            return fail(SYSID_ERR_INVALID, "could not convert %lld field(s) to float exactly; first at row %lld, column %lld", h.bad_fields, r, c);
        }
    }
    return SYSID_OK;
}

int sysid_fd_rate(const double* tick, const double* x, double* y, int32_t channels, int64_t N, int64_t ld_x, int64_t ld_y,
                  double scale, void* stream) {
    if (!tick || !x || !y) return fail(SYSID_ERR_INVALID, "null argument");
    if (channels < 0 || N < 0 || ld_x < N || ld_y < N) return fail(SYSID_ERR_INVALID, "bad channels/N/ld");
    if (x == y) return fail(SYSID_ERR_INVALID, "fd_rate cannot run in place");
    if (channels == 0 || N == 0) return SYSID_OK;
    FdArgs g{tick, x, y, N, ld_x, ld_y, channels, scale};
    fd_rate_kernel<<<stream_grid((long long)channels * N), 256, 0, (cudaStream_t)stream>>>(g);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

int sysid_contact_from_tau(const double* tau, double* out, int64_t N, double hi, double lo, void* stream) {
    if (!tau || !out) return fail(SYSID_ERR_INVALID, "null argument");
    if (N < 0) return fail(SYSID_ERR_INVALID, "bad N");
    if (N == 0) return SYSID_OK;
    contact_label_kernel<<<stream_grid(N), 256, 0, (cudaStream_t)stream>>>(tau, out, N, hi, lo);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

int sysid_round_dat(const double* x, double* y, int32_t channels, int64_t N, int64_t ld_x, int64_t ld_y, int32_t to_float32,
                    void* stream) {
    if (!x || !y) return fail(SYSID_ERR_INVALID, "null argument");
    if (channels < 0 || N < 0 || ld_x < N || ld_y < N) return fail(SYSID_ERR_INVALID, "bad channels/N/ld");
    if (channels == 0 || N == 0) return SYSID_OK;
    RoundArgs g{x, y, N, ld_x, ld_y, channels, to_float32};
    round_dat_kernel<<<stream_grid((long long)channels * N), 256, 0, (cudaStream_t)stream>>>(g);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

}  // extern "C"
