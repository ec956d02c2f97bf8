// C-ABI entry points of SURVEY 8f row f4: sysid_tsqr (behind Solver.solve_llsq_svd) and sysid_physical_consistency.
#include <cstdarg>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../include/sysid_b200.h"
#include "tsqr_kernels.cuh"

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

constexpr int TSQR_GRID0 = 148;      // one CTA per SM for the leaf level
constexpr int TSQR_FANIN = 12;       // triangles merged per CTA at the inner levels

}  // namespace

extern "C" {

size_t sysid_tsqr_workspace_bytes(int32_t c) {
    if (c < 1 || c + 1 > TSQR_MAXC) return 0;
    const size_t ca = (size_t)c + 1;
    // level 0 writes up to 148 triangles, level 1 up to 13, ...: two ping-pong areas of the larger size are plenty
    return sizeof(double) * ca * ca * (TSQR_GRID0 + (TSQR_GRID0 + TSQR_FANIN - 1) / TSQR_FANIN + 1);
}

int sysid_tsqr(const double* A, const double* b, int64_t rows, int32_t c, double* R_out, void* workspace, size_t workspace_bytes,
               void* stream) {
    if (!A || !R_out || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (rows < 1) return fail(SYSID_ERR_INVALID, "bad rows");
    if (c < 1 || c + 1 > TSQR_MAXC) return fail(SYSID_ERR_UNSUPPORTED, "c + 1 must be <= %d", TSQR_MAXC);
    if (workspace_bytes < sysid_tsqr_workspace_bytes(c)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int ca = c + 1;
    const size_t smem = tsqr_smem_bytes(ca);
    CUDA_TRY(cudaFuncSetAttribute(tsqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double* area0 = (double*)workspace;
    double* area1 = area0 + (size_t)ca * ca * TSQR_GRID0;
    // level 0: the stack itself
    long long nblk = (rows + TSQR_ROWS - 1) / TSQR_ROWS;
    int grid = (int)(nblk < TSQR_GRID0 ? nblk : TSQR_GRID0);
    TsqrArgs g;
    g.A = A; g.lda = c; g.ncolsA = c; g.b = b; g.rows = rows; g.ca = ca;
    g.out = (grid == 1) ? R_out : area0;
    tsqr_kernel<<<grid, TSQR_THREADS, smem, st>>>(g);
    CUDA_TRY(cudaGetLastError());
    // inner levels: the triangles of the previous level, stacked, are the next matrix (already augmented)
    double* src = area0;
    double* dst = area1;
    while (grid > 1) {
        const long long r = (long long)grid * ca;
        int next = (grid + TSQR_FANIN - 1) / TSQR_FANIN;
        // a CTA takes whole blocks of 64 rows: never more CTAs than blocks
        const long long nb = (r + TSQR_ROWS - 1) / TSQR_ROWS;
        if (next > nb) next = (int)nb;
        TsqrArgs h;
        h.A = src; h.lda = ca; h.ncolsA = ca; h.b = nullptr; h.rows = r; h.ca = ca;
        h.out = (next == 1) ? R_out : dst;
        tsqr_kernel<<<next, TSQR_THREADS, smem, st>>>(h);
        CUDA_TRY(cudaGetLastError());
        double* tmp = src; src = dst; dst = tmp;
        grid = next;
    }
    return SYSID_OK;
}

int sysid_physical_consistency(const double* phi, int64_t phi_stride, int32_t batch, int32_t num_links, const double* semi_axes,
                               const double* centers, double* out, void* stream) {
    if (!phi || !semi_axes || !centers || !out) return fail(SYSID_ERR_INVALID, "null argument");
    if (batch < 0 || num_links < 1 || phi_stride < 10LL * num_links) return fail(SYSID_ERR_INVALID, "bad batch/num_links/stride");
    if (batch == 0) return SYSID_OK;
    ConsistencyArgs g{phi, phi_stride, batch, num_links, semi_axes, centers, out};
    const long long items = (long long)batch * num_links;
    consistency_kernel<<<(unsigned)((items + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

}  // extern "C"
