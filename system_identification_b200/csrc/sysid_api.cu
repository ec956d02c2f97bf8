// C-ABI of the B200-native inertial-identification hot path (see include/sysid_b200.h).
// Host side of the library: model validation/flattening into the kernel-parameter image, launches, error reporting.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sysid_b200.h"
#include "gram_kernels.cuh"
#include "gram_struct.cuh"
#include "sdp_kernels.cuh"
#include "filter_kernels.cuh"
#include "bigmodel.cuh"

using namespace sysid;

// Internal streams / events of the host-streaming entry, created once per model handle (creating and destroying them per call
// costs host time in front of a 19 ms stream).  One call at a time may borrow them (try-lock); a concurrent call on the same
// handle falls back to per-call resources.
struct HostStreamRes {
    cudaStream_t s_copy = nullptr, s_solve = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // copied[2], consumed[2], start, snap, solved, snap2
    bool ok = false;
    bool create() {
        if (cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&s_solve, cudaStreamNonBlocking) != cudaSuccess) return false;
        for (int i = 0; i < 8; ++i) if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return false;
        ok = true;
        return true;
    }
    void destroy() {
        for (int i = 0; i < 8; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
        if (s_copy) cudaStreamDestroy(s_copy);       // pending work completes first; the runtime releases the stream afterwards
        if (s_solve) cudaStreamDestroy(s_solve);
        *this = HostStreamRes();
    }
};

struct sysid_model {
    DevModel dev;                  // fused path (trees inside the compiled envelope); dims are valid for every model
    bool big = false;              // tree outside the envelope (G1-29dof): stages 1-2 through bigmodel.cuh
    big::BigModel bm;
    int sm_count;
    mutable std::mutex host_mu;
    mutable HostStreamRes host_res;
};

namespace {
thread_local char g_err[512] = "";
}

// the other translation units of the library (ingest_api.cu) record their error text through this
namespace sysid {
int set_error(int code, const char* message) {
    snprintf(g_err, sizeof(g_err), "%s", message);
    return code;
}
}  // namespace sysid

namespace {
__global__ void add_rankloss_kernel(const int* __restrict__ count, long long* __restrict__ info) { info[0] += (long long)count[0]; }
}  // namespace

namespace {

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) return fail(SYSID_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

int device_sm_count(int* out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return SYSID_OK;
}

template <typename K>
int opt_in_smem(K kernel, size_t bytes) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return SYSID_OK;
}

SampleIO make_io(const double* q, const double* dq, const double* ddq, const double* tau, const double* cnt,
                 const double* w, int64_t ld) {
    SampleIO io;
    io.q = q; io.dq = dq; io.ddq = ddq; io.tau = tau; io.cnt = cnt; io.weights = w; io.ld = ld;
    return io;
}

// Does the structured-basis Gram kernel (gram_struct.cuh) serve this tree?  Every child of the root must head a simple chain
// with consecutive joint numbers, at most one contact frame per chain (three or more joints from the root), at most four chains,
// and the chains must split into two column classes of at most six joints.  Fills the st_* plan of the model image.
void plan_struct_basis(DevModel& M) {
    M.st_ok = 0;
    if (M.nfch < 1 || M.nfch > ST_MAXLEG) return;
    int total = 0;
    for (int c = 0; c < M.nfch; ++c) {
        if (M.fch_own[c] != 0 || M.parent[M.fch[c][0]] != 1) return;
        for (int e = 1; e < M.fch_len[c]; ++e) if (M.fch[c][e] != M.fch[c][0] + e || M.parent[M.fch[c][e]] != M.fch[c][e - 1]) return;
        total += M.fch_len[c];
        M.st_cfoot[c] = -1;
    }
    if (total != M.nd) return;
    int leg_of[MAXJ];
    for (int j = 0; j < MAXJ; ++j) leg_of[j] = -1;
    for (int c = 0; c < M.nfch; ++c) for (int e = 0; e < M.fch_len[c]; ++e) leg_of[M.fch[c][e]] = c;
    for (int k = 0; k < M.n_ee; ++k) {
        if (M.ee_joint[k] == 1) continue;                          // a frame on the root body: base part only
        const int c = leg_of[M.ee_joint[k]];
        if (c < 0 || M.st_cfoot[c] >= 0 || M.chain_len[k] < 3) return;
        for (int e = 0; e < M.chain_len[k]; ++e) if (M.chain[k][e] != M.fch[c][0] + M.chain_len[k] - 1 - e) return;
        M.st_cfoot[c] = (int8_t)k;
    }
    // two classes of at most six joints: the most balanced split
    int best = -1, best_diff = 99;
    for (int mask = 0; mask < (1 << M.nfch); ++mask) {
        int nA = 0, nB = 0;
        for (int c = 0; c < M.nfch; ++c) ((mask >> c) & 1 ? nB : nA) += M.fch_len[c];
        if (nA > 6 || nB > 6) continue;
        const int diff = nA > nB ? nA - nB : nB - nA;
        if (diff < best_diff) { best_diff = diff; best = mask; }
    }
    if (best < 0) return;
    int nslot[2] = {0, 0};
    for (int X = 0; X < 2; ++X) for (int e = 0; e < 6; ++e) M.st_slotjoint[X][e] = -1;
    for (int c = 0; c < M.nfch; ++c) {
        const int X = (best >> c) & 1;
        M.st_ccls[c] = (int8_t)X;
        for (int e = 0; e < M.fch_len[c]; ++e) {
            const int j = M.fch[c][e], slot = nslot[X]++;
            M.st_slotjoint[X][slot] = (int8_t)j;
            M.st_jrec[j] = (uint32_t)c | ((uint32_t)e << 2) | ((uint32_t)X << 5) | ((uint32_t)slot << 6) |
                           ((uint32_t)(72 * X + 10 * slot) << 9) | ((uint32_t)(72 * X + 60 + slot) << 17);
        }
    }
    M.st_nslot[0] = (int8_t)nslot[0]; M.st_nslot[1] = (int8_t)nslot[1];
    { int nf = 0, ml = 0; for (int c = 0; c < M.nfch; ++c) { nf += M.st_cfoot[c] >= 0; if (M.fch_len[c] > ml) ml = M.fch_len[c]; } M.st_nred = (int8_t)(6 + 3 * nf); M.st_maxlen = (int8_t)ml; }
    // warp tasks of the tile fill, heaviest first (cost ~ bodies emitted + joints walked)
    {
        uint32_t code[64];
        int cost[64], n = 0;
        // bodies per task: runs of two share the walk down a six-joint leg (G1: 73.6 Msamples/s against 71.0 for one body per task
        // and 72.2 for runs of three); a three-joint leg is one task (Solo: 67.8 against 66.7 / 67.6 for one / two bodies per task).
        // SYSID_ST_GROUP overrides, for measurements.
        static const int group_env = [] { const char* e = std::getenv("SYSID_ST_GROUP"); return e ? std::atoi(e) : 0; }();
        const int group = group_env > 0 ? group_env : (M.st_maxlen > 3 ? ST_GROUP : 3);
        for (int c = 0; c < M.nfch; ++c)
            for (int e0 = 0; e0 < M.fch_len[c]; e0 += group) {
                const int cnt = (M.fch_len[c] - e0 < group) ? M.fch_len[c] - e0 : group;
                const int j = M.fch[c][e0], X = M.st_ccls[c], slot = (int)((M.st_jrec[j] >> 6) & 7);
                code[n] = 0u | ((uint32_t)j << 4) | ((uint32_t)cnt << 8); cost[n] = 20 * cnt + 3 * (e0 + cnt); ++n;
                code[n] = (uint32_t)(2 + X) | ((uint32_t)j << 4) | ((uint32_t)cnt << 8) | ((uint32_t)slot << 12) | ((slot == 0 ? 1u : 0u) << 16);
                cost[n] = 14 * cnt + 3 * (e0 + cnt) + (slot == 0 ? 3 : 0); ++n;
            }
        code[n] = 1u; cost[n] = 19; ++n;
        if (n > 32 || n > 2 * GRAM_WARPS) return;
        for (int a = 0; a < n; ++a)
            for (int b = a + 1; b < n; ++b)
                if (cost[b] > cost[a]) { int tc = cost[a]; cost[a] = cost[b]; cost[b] = tc; const uint32_t tk = code[a]; code[a] = code[b]; code[b] = tk; }
        M.st_ntask = (int8_t)n;
        for (int a = 0; a < n; ++a) M.st_task[a] = code[a];
    }
    M.st_ok = 1;
}

// column of the statistics (0 .. c, c = the torque column) -> column of the Gram tiles
ColMap make_colmap(const DevModel& M, int friction, bool structured) {
    ColMap cm;
    for (int i = 0; i < CW; ++i) cm.p[i] = (uint8_t)i;
    if (!structured) return cm;
    const int np = M.nparams, nd = M.nd;
    for (int col = 0; col < np; ++col) {
        const int j = col / 10 + 1;
        cm.p[col] = (uint8_t)((j == 1 ? ST_ROOTCOL : (int)((M.st_jrec[j] >> 9) & 255)) + col % 10);
    }
    int c = np;
    if (friction) {
        for (int k = 0; k < nd; ++k) { cm.p[np + k] = (uint8_t)((M.st_jrec[k + 2] >> 17) & 255); cm.p[np + nd + k] = (uint8_t)(((M.st_jrec[k + 2] >> 17) & 255) + 6); }
        c = np + 2 * nd;
    }
    cm.p[c] = (uint8_t)ST_TAUCOL;
    return cm;
}

// SYSID_GRAM_LEGACY=1 forces the unstructured kernel (A/B timing, parity of the two kernels against each other)
bool use_struct(const sysid_model* model) {
    static const bool legacy = [] { const char* e = std::getenv("SYSID_GRAM_LEGACY"); return e && e[0] == '1'; }();
    return model->dev.st_ok && !legacy;
}

}  // namespace

extern "C" {

int sysid_abi_version(void) { return SYSID_ABI_VERSION; }

const char* sysid_last_error(void) { return g_err; }

void sysid_get_limits(sysid_limits* out) {
    if (!out) return;
    out->max_bodies = MAXB; out->max_nv = MAXV; out->max_ee = MAXEE; out->max_depth = MAXCH; out->max_cols_padded = CW;
}

int sysid_model_create(const sysid_tree_desc* d, sysid_model** out) {
    if (!d || !out) return fail(SYSID_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!d->parent || !d->jtype || !d->axis || !d->place_R || !d->place_p) return fail(SYSID_ERR_INVALID, "null tree array");
    if (d->njoints < 2) return fail(SYSID_ERR_INVALID, "tree needs at least the universe and a root joint");
    if (d->njoints > big::BJ) return fail(SYSID_ERR_UNSUPPORTED, "njoints %d exceeds this build's envelope (%d)", d->njoints, big::BJ);
    if (d->n_ee < 0 || d->n_ee > MAXEE) return fail(SYSID_ERR_UNSUPPORTED, "n_ee %d exceeds this build's envelope (%d)", d->n_ee, MAXEE);
    if (d->n_ee > 0 && (!d->ee_joint || !d->ee_offset)) return fail(SYSID_ERR_INVALID, "null end-effector array");
    if (d->jtype[1] != SYSID_JT_FREEFLYER || d->parent[1] != 0)
        return fail(SYSID_ERR_UNSUPPORTED, "joint 1 must be the free-flyer root (floating_base=True is the accelerated path)");
    sysid_model* m = new (std::nothrow) sysid_model;
    if (!m) return fail(SYSID_ERR_INVALID, "out of host memory");
    std::memset(&m->dev, 0, sizeof(DevModel));
    std::memset(&m->bm, 0, sizeof(big::BigModel));
    DevModel& M = m->dev;
    {
        // does the tree fit the fused kernels' envelope?  (bodies, chain depth; everything else is checked below)
        int depth_[big::BJ] = {0}, maxdepth = 0;
        bool sane = true;
        for (int j = 2; j < d->njoints; ++j) {
            if (d->parent[j] < 1 || d->parent[j] >= j) { sane = false; break; }
            depth_[j] = depth_[d->parent[j]] + 1;
            if (depth_[j] > maxdepth) maxdepth = depth_[j];
        }
        if (sane && (d->njoints > MAXJ || maxdepth > MAXCH)) {
            big::BigModel& B = m->bm;
            const int nb = d->njoints - 1;
            if (10 * nb + 2 * (nb - 1) + 1 > big::BCW || 6 + nb - 1 > big::BV) { delete m; return fail(SYSID_ERR_UNSUPPORTED, "tree with %d joints exceeds the large-model path (%d columns)", d->njoints, big::BCW); }
            B.njoints = d->njoints; B.nb = nb; B.n_ee = d->n_ee; B.nv = 6 + nb - 1; B.nq = 7 + nb - 1; B.nd = nb - 1; B.nparams = 10 * nb;
            for (int j = 0; j < d->njoints; ++j) {
                B.parent[j] = d->parent[j]; B.jtype[j] = d->jtype[j];
                B.idx_v[j] = (j <= 1) ? 0 : 6 + (j - 2); B.idx_q[j] = (j <= 1) ? 0 : 7 + (j - 2);
                for (int k = 0; k < 3; ++k) { B.axis[j][k] = d->axis[3 * j + k]; B.pp[j][k] = d->place_p[3 * j + k]; }
                for (int k = 0; k < 9; ++k) B.pR[j][k] = d->place_R[9 * j + k];
                if (j >= 2 && (d->jtype[j] < SYSID_JT_RX || d->jtype[j] > SYSID_JT_RU)) { delete m; return fail(SYSID_ERR_UNSUPPORTED, "joint %d: only revolute joints below the free-flyer root", j); }
            }
            for (int k = 0; k < 3; ++k) B.gravity[k] = d->gravity[k];
            for (int k = 0; k < d->n_ee; ++k) {
                if (d->ee_joint[k] < 1 || d->ee_joint[k] >= d->njoints) { delete m; return fail(SYSID_ERR_INVALID, "end effector %d: joint %d out of range", k, d->ee_joint[k]); }
                B.ee_joint[k] = d->ee_joint[k];
                for (int e = 0; e < 3; ++e) B.ee_off[k][e] = d->ee_offset[3 * k + e];
            }
            m->big = true;
            M.njoints = B.njoints; M.nb = B.nb; M.n_ee = B.n_ee; M.nv = B.nv; M.nq = B.nq; M.nd = B.nd; M.nparams = B.nparams;   // dims only
            int rc = device_sm_count(&m->sm_count);
            if (rc != SYSID_OK) { delete m; return rc; }
            *out = m;
            return SYSID_OK;
        }
    }
    M.njoints = d->njoints; M.nb = d->njoints - 1; M.n_ee = d->n_ee;
    M.nv = 6 + (M.nb - 1); M.nq = 7 + (M.nb - 1); M.nd = M.nb - 1; M.nparams = 10 * M.nb;
    for (int j = 0; j < d->njoints; ++j) {
        M.parent[j] = d->parent[j];
        M.jtype[j] = d->jtype[j];
        for (int k = 0; k < 3; ++k) { M.axis[j][k] = d->axis[3 * j + k]; M.pp[j][k] = d->place_p[3 * j + k]; }
        for (int k = 0; k < 9; ++k) M.pR[j][k] = d->place_R[9 * j + k];
        if (j >= 2) {
            if (d->jtype[j] < SYSID_JT_RX || d->jtype[j] > SYSID_JT_RU) { delete m; return fail(SYSID_ERR_UNSUPPORTED, "joint %d: only revolute joints below the free-flyer root", j); }
            if (d->parent[j] < 1 || d->parent[j] >= j) { delete m; return fail(SYSID_ERR_INVALID, "joint %d: parent %d breaks the depth-first numbering", j, d->parent[j]); }
            if (d->jtype[j] == SYSID_JT_RU) {
                const double nrm = std::sqrt(M.axis[j][0] * M.axis[j][0] + M.axis[j][1] * M.axis[j][1] + M.axis[j][2] * M.axis[j][2]);
                if (!(std::fabs(nrm - 1.0) < 1e-9)) { delete m; return fail(SYSID_ERR_INVALID, "joint %d: axis is not unit length", j); }
            }
        }
    }
    for (int k = 0; k < 3; ++k) M.gravity[k] = d->gravity[k];
    for (int k = 0; k < d->n_ee; ++k) {
        const int jf = d->ee_joint[k];
        if (jf < 1 || jf >= d->njoints) { delete m; return fail(SYSID_ERR_INVALID, "end effector %d: joint %d out of range", k, jf); }
        M.ee_joint[k] = jf;
        for (int e = 0; e < 3; ++e) M.ee_off[k][e] = d->ee_offset[3 * k + e];
        int len = 0;
        for (int c = jf; c > 1; c = M.parent[c]) {
            if (len >= MAXCH) { delete m; return fail(SYSID_ERR_UNSUPPORTED, "end effector %d: chain longer than %d joints", k, MAXCH); }
            M.chain[k][len++] = c;
        }
        M.chain_len[k] = len;
    }
    for (int a = 0; a < d->n_ee; ++a)
        for (int b = 0; b < d->n_ee; ++b) {
            int ns = 0;
            while (ns < M.chain_len[a] && ns < M.chain_len[b] &&
                   M.chain[a][M.chain_len[a] - 1 - ns] == M.chain[b][M.chain_len[b] - 1 - ns]) ++ns;
            M.nshared[a][b] = ns;
        }
    // chain depth of every body (the column walk is unrolled MAXCH deep) and the depth-sorted column order
    int depth[MAXJ] = {0};
    for (int j = 2; j < d->njoints; ++j) {
        depth[j] = depth[M.parent[j]] + 1;
        if (depth[j] > MAXCH) { delete m; return fail(SYSID_ERR_UNSUPPORTED, "joint %d sits %d joints below the root (limit %d)", j, depth[j], MAXCH); }
    }
    {
        int work[CW];
        for (int c = 0; c < CW; ++c) {
            if (c < M.nparams) work[c] = 2 + depth[c / 10 + 1];          // body column: base rows + one emit per ancestor
            else if (c < M.nparams + 2 * M.nd) work[c] = 1;              // friction column: one emit
            else if (c == M.nparams + 2 * M.nd) work[c] = 2 + MAXCH;     // tau column: 12 emits
            else work[c] = 0;                                            // padding
        }
        int n = 0;
        for (int wv = 2 + MAXCH; wv >= 0; --wv)
            for (int c = 0; c < CW; ++c) if (work[c] == wv) M.colperm[n++] = (uint8_t)c;
    }
    // leaf chains (root -> leaf), each joint owned by the first chain that contains it
    {
        bool has_child[MAXJ] = {false}, owned[MAXJ] = {false};
        for (int j = 2; j < d->njoints; ++j) has_child[M.parent[j]] = true;
        M.nfch = 0;
        for (int j = 2; j < d->njoints; ++j) {
            if (has_child[j]) continue;
            const int c = M.nfch++;
            int len = depth[j], e = len;
            for (int k = j; k > 1; k = M.parent[k]) M.fch[c][--e] = k;
            M.fch_len[c] = len;
            int own = 0;
            while (own < len && owned[M.fch[c][own]]) ++own;
            M.fch_own[c] = own;
            for (int k = own; k < len; ++k) owned[M.fch[c][k]] = true;
        }
    }
    // tile fill: (chains x split + 1) lane groups of 96 lanes; use up to two groups per chain while all fit one pass
    M.fill_split = ((M.nfch * 2 + 1) * 96 <= GRAM_THREADS) ? 2 : 1;
    for (int i = 1; i < d->njoints; ++i)
        for (int k = i; k >= 1; k = M.parent[k]) M.submask[k] |= 1u << i;       // i is in the subtree of each of its ancestors
    // projection phase: the bodies [root, owned joints of chain 0, of chain 1, ...] are cut into PROJ_PARTS contiguous runs of
    // (almost) equal length; a run that starts inside a chain first walks the chain's earlier joints without emitting them
    {
        int seq_joint[MAXB], seq_chain[MAXB], seq_pos[MAXB], nseq = 0;
        seq_joint[nseq] = 1; seq_chain[nseq] = -1; seq_pos[nseq] = 0; ++nseq;      // the root body
        for (int c = 0; c < M.nfch; ++c)
            for (int e = M.fch_own[c]; e < M.fch_len[c]; ++e) { seq_joint[nseq] = M.fch[c][e]; seq_chain[nseq] = c; seq_pos[nseq] = e; ++nseq; }
        int cost[PROJ_PARTS] = {0, 0, 0, 0};
        int lo = 0;
        for (int p = 0; p < PROJ_PARTS; ++p) {
            const int hi = lo + nseq / PROJ_PARTS + (p < nseq % PROJ_PARTS ? 1 : 0);
            int n = 0, cur_chain = -2;
            bool ok = true;
            for (int u = lo; u < hi && ok; ++u) {
                if (seq_chain[u] < 0) { M.proj_root[p] = 1; cost[p] += 4; continue; }
                const int c = seq_chain[u];
                if (c != cur_chain) {                     // a new walk: the chain's joints before this one are accumulated only
                    for (int e = 0; e < seq_pos[u] && ok; ++e) {
                        if (n >= PROJ_MAXITEMS) { ok = false; break; }
                        M.proj_item[p][n] = (uint32_t)M.fch[c][e] | ((e == 0 ? 2u : 0u) << 8); ++n; cost[p] += 1;
                    }
                    cur_chain = c;
                }
                if (n >= PROJ_MAXITEMS) { ok = false; break; }
                M.proj_item[p][n] = (uint32_t)seq_joint[u] | ((1u | (seq_pos[u] == 0 ? 2u : 0u)) << 8); ++n; cost[p] += 4;
            }
            if (!ok) { delete m; return fail(SYSID_ERR_UNSUPPORTED, "projection plan: more than %d chain steps in one part", PROJ_MAXITEMS); }
            M.proj_n[p] = (int8_t)n;
            lo = hi;
        }
        // friction / torque / zero-padding column groups from column nparams to CW: each to the currently cheapest part
        const int ntail = (CW - M.nparams + 7) / 8;
        if (ntail > 32) { delete m; return fail(SYSID_ERR_UNSUPPORTED, "projection plan: %d tail column groups", ntail); }
        for (int nt = 0; nt < ntail; ++nt) {
            int best = 0;
            for (int p = 1; p < PROJ_PARTS; ++p) if (cost[p] < cost[best]) best = p;
            M.proj_tail[best] |= 1u << nt; cost[best] += 3;
            // block (joints 4 ks .., columns 8 nt ..) of [diag(dq) | diag(sign dq) | tau] is non-zero where a diagonal or the
            // torque column crosses it
            for (int fr = 0; fr < 2; ++fr) {
                const int nd = M.nd, tcol = fr ? 2 * nd : 0, c0 = 8 * nt;
                for (int ks = 0; ks < 3 && 4 * ks < nd; ++ks) {
                    const int j0 = 4 * ks;
                    bool need = (tcol >= c0 && tcol < c0 + 8);
                    if (fr) need = need || (c0 < j0 + 4 && j0 < c0 + 8) || (c0 < nd + j0 + 4 && nd + j0 < c0 + 8);
                    if (need) M.proj_tailks[fr][nt] |= (uint8_t)(1u << ks);
                }
            }
        }
    }
    plan_struct_basis(M);
    int rc = device_sm_count(&m->sm_count);
    if (rc != SYSID_OK) { delete m; return rc; }
    *out = m;
    return SYSID_OK;
}

void sysid_model_destroy(sysid_model* model) {
    if (model && model->host_res.ok) model->host_res.destroy();
    delete model;
}

int sysid_model_dims(const sysid_model* model, sysid_dims* out) {
    if (!model || !out) return fail(SYSID_ERR_INVALID, "null argument");
    const DevModel& M = model->dev;
    out->nq = M.nq; out->nv = M.nv; out->nbodies = M.nb; out->ndof = M.nd; out->nparams = M.nparams;
    out->ncols = M.nparams + 2 * M.nd; out->n_ee = M.n_ee;
    return SYSID_OK;
}

// ---- large-model path (bigmodel.cuh): one chunk of <= BIG_CHUNK samples through kin -> rows -> tail -> contact -> zrows ----------
static int big_chunk_rows(const sysid_model* model, const SampleIO& io, long long base, int ns, int friction, bool contacts,
                          const big::BigWs& w, cudaStream_t st) {
    const big::BigModel& B = model->bm;
    CUDA_TRY(cudaMemsetAsync(w.Yt, 0, sizeof(double) * (size_t)ns * B.nv * big::BCW, st));
    big::big_kin_kernel<<<(ns + 127) / 128, 128, 0, st>>>(B, io, base, ns, w.kin);
    big::big_rows_kernel<<<(ns * B.nb + 127) / 128, 128, 0, st>>>(B, ns, w.kin, w.Yt);
    big::big_tail_kernel<<<(ns * B.nd + 127) / 128, 128, 0, st>>>(B, io, base, ns, friction, w.Yt);
    if (contacts) {
        big::big_contact_kernel<<<(ns + 63) / 64, 64, 0, st>>>(B, io, base, ns, w.kin, w.W, w.m3, w.rankloss);
        const long long tot = (long long)ns * big::BCW;
        big::big_zrows_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(B, big::big_row_masks(B, friction), ns, w.Yt, w.W, w.m3, w.Z);
    }
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

static size_t big_workspace_bytes(const sysid_model* model) { return big::big_workspace(model->bm, nullptr, true).bytes + 256; }

// per-sample outputs of the large-model path need scratch the small-model ABI has no argument for: stream-ordered allocation
static int big_batch(const sysid_model* model, const SampleIO& io, int64_t N, int friction, double* Y, double* A, double* b, double* P, cudaStream_t st) {
    void* buf = nullptr;
    CUDA_TRY(cudaMallocAsync(&buf, big::big_workspace(model->bm, nullptr).bytes + 256, st));
    const big::BigWs w = big::big_workspace(model->bm, buf);
    const big::BigModel& B = model->bm;
    const int ncols = B.nparams + (friction ? 2 * B.nd : 0);
    int rc = SYSID_OK;
    for (int64_t lo = 0; lo < N && rc == SYSID_OK; lo += big::BIG_CHUNK) {
        const int ns = (int)((N - lo < big::BIG_CHUNK) ? (N - lo) : big::BIG_CHUNK);
        rc = big_chunk_rows(model, io, lo, ns, friction, A || b || P, w, st);
        if (rc != SYSID_OK) break;
        const long long tot = (long long)ns * B.nv * big::BCW;
        big::big_emit_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(B, ns, ncols, ncols, w.Yt, w.W, w.Z, w.m3,
            Y ? Y + (size_t)lo * B.nv * B.nparams : nullptr, A ? A + (size_t)lo * B.nv * ncols : nullptr,
            b ? b + (size_t)lo * B.nv : nullptr, P ? P + (size_t)lo * B.nv * B.nv : nullptr);
        if (cudaGetLastError() != cudaSuccess) rc = fail(SYSID_ERR_CUDA, "large-model emit kernel failed");
    }
    cudaFreeAsync(buf, st);
    return rc;
}

int sysid_regressor_batch(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                          int64_t N, int64_t ld, double* Y_out, void* stream) {
    if (!model || !q || !dq || !ddq || !Y_out) return fail(SYSID_ERR_INVALID, "null argument");
    if (N < 0 || ld < N) return fail(SYSID_ERR_INVALID, "bad N/ld");
    if (N == 0) return SYSID_OK;
    if (model->big) return big_batch(model, make_io(q, dq, ddq, nullptr, nullptr, nullptr, ld), N, 0, Y_out, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    const DevModel& M = model->dev;
    CUDA_TRY(cudaMemsetAsync(Y_out, 0, sizeof(double) * (size_t)N * M.nv * M.nparams, st));
    int rc = opt_in_smem(sample_batch_kernel<0>, DBG_SMEM_BYTES);
    if (rc) return rc;
    BatchArgs a{};
    a.io = make_io(q, dq, ddq, nullptr, nullptr, nullptr, ld);
    a.N = N; a.friction = 0; a.Y = Y_out;
    // the raw regressor does not depend on contacts: run with n_ee = 0 in a local copy of the model image,
    // so the (null) contact channel is never read
    DevModel Mloc = M;
    Mloc.n_ee = 0;
    const unsigned grid = (unsigned)((N + SB_SAMPLES - 1) / SB_SAMPLES);
    sample_batch_kernel<0><<<grid, DBG_THREADS, DBG_SMEM_BYTES, st>>>(Mloc, a);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

int sysid_projected_batch(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                          const double* tau, const double* contact, int64_t N, int64_t ld, int32_t friction,
                          double* A_out, double* b_out, double* P_out, void* stream) {
    if (!model || !q || !dq || !ddq || !tau || !A_out || !b_out) return fail(SYSID_ERR_INVALID, "null argument");
    if (model->dev.n_ee > 0 && !contact) return fail(SYSID_ERR_INVALID, "null contact array");
    if (N < 0 || ld < N) return fail(SYSID_ERR_INVALID, "bad N/ld");
    if (N == 0) return SYSID_OK;
    if (model->big) return big_batch(model, make_io(q, dq, ddq, tau, contact, nullptr, ld), N, friction ? 1 : 0, nullptr, A_out, b_out, P_out, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = opt_in_smem(sample_batch_kernel<1>, DBG_SMEM_BYTES);
    if (rc) return rc;
    BatchArgs a{};
    a.io = make_io(q, dq, ddq, tau, contact, nullptr, ld);
    a.N = N; a.friction = friction ? 1 : 0; a.A = A_out; a.b = b_out; a.P = P_out;
    const unsigned grid = (unsigned)((N + SB_SAMPLES - 1) / SB_SAMPLES);
    sample_batch_kernel<1><<<grid, DBG_THREADS, DBG_SMEM_BYTES, st>>>(model->dev, a);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

size_t sysid_stats_len(const sysid_model* model, int32_t friction) {
    if (!model) return 0;
    const size_t c = (size_t)model->dev.nparams + (friction ? 2 * (size_t)model->dev.nd : 0);
    return c * c + c + 2;
}

size_t sysid_gram_workspace_bytes(const sysid_model* model) {
    if (!model) return 0;
    if (model->big) return big_workspace_bytes(model);
    return sizeof(double) * (size_t)model->sm_count * PARTIAL_DOUBLES;
}

// large-model statistics: chunks through HBM, two DMMA SYRK passes per chunk (Ytilde^T Ytilde - Z^T Z), deterministic reduction
static int big_gram(const sysid_model* model, const SampleIO& io, int64_t N, int friction, double* stats, int64_t* info,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (io.weights) return fail(SYSID_ERR_UNSUPPORTED, "per-sample weights are not available on the large-model path");
    if (workspace_bytes < big_workspace_bytes(model)) return fail(SYSID_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, big_workspace_bytes(model));
    const big::BigModel& B = model->bm;
    const big::BigWs w = big::big_workspace(B, (void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255), true);
    const big::BigWs walt = big::big_workspace_alt(w);
    const int c = B.nparams + (friction ? 2 * B.nd : 0);
    CUDA_TRY(cudaMemsetAsync(w.partial, 0, sizeof(double) * (size_t)big::BIG_NZ * big::SY_NBLK * big::SY_BLK * big::SY_BLK, st));
    CUDA_TRY(cudaMemsetAsync(w.rankloss, 0, sizeof(int) * 4, st));
    const big::RowMasks ymasks = big::big_row_masks(B, friction), zmasks = big::big_z_masks(B);
    CUDA_TRY(cudaFuncSetAttribute(big::big_syrk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big::SY_ROWS_SMEM));
    // Two streams, two sets of chunk buffers: the rows of chunk i + 1 (thread-per-sample kernels bound by their own latency, which leave
    // most of every SM idle) are built on `stream` while the two SYRK passes of chunk i run on an internal one; the SYRK launches stay
    // in order on that stream (one partial buffer), the reduction waits for the last of them.
    struct Side {
        cudaStream_t s = nullptr; cudaEvent_t rows[2] = {nullptr, nullptr}, syrk[2] = {nullptr, nullptr};
        bool create() {
            if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return false;
            for (int k = 0; k < 2; ++k)
                if (cudaEventCreateWithFlags(&rows[k], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&syrk[k], cudaEventDisableTiming) != cudaSuccess) return false;
            return true;
        }
        ~Side() {           // destroying a stream / an event with work pending is deferred by the runtime until that work is done
            for (int k = 0; k < 2; ++k) { if (rows[k]) cudaEventDestroy(rows[k]); if (syrk[k]) cudaEventDestroy(syrk[k]); }
            if (s) cudaStreamDestroy(s);
        }
    } side;
    if (!side.create()) return fail(SYSID_ERR_CUDA, "creating the internal stream of the large-model statistics failed");
    int64_t i = 0;
    for (int64_t lo = 0; lo < N; lo += big::BIG_CHUNK, ++i) {
        const int ns = (int)((N - lo < big::BIG_CHUNK) ? (N - lo) : big::BIG_CHUNK);
        const int k = (int)(i & 1);
        const big::BigWs& wk = k ? walt : w;
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(st, side.syrk[k], 0));          // the SYRK of chunk i - 2 has consumed this set
        int rc = big_chunk_rows(model, io, lo, ns, friction, true, wk, st);
        if (rc != SYSID_OK) return rc;
        CUDA_TRY(cudaEventRecord(side.rows[k], st));
        CUDA_TRY(cudaStreamWaitEvent(side.s, side.rows[k], 0));
        // Ytilde^T Ytilde class by class on the structural masks of its rows, then - Z^T Z (dense rows, 3 n_ee slots per sample)
        big::big_syrk_rows_kernel<<<dim3(big::SY_NBLK, big::BIG_NZ), big::SY_THREADS, big::SY_ROWS_SMEM, side.s>>>(wk.Yt, ns, B.nv, ymasks, 1.0, w.partial);
        big::big_syrk_rows_kernel<<<dim3(big::SY_NBLK, big::BIG_NZ), big::SY_THREADS, big::SY_ROWS_SMEM, side.s>>>(wk.Z, ns, big::BMR, zmasks, -1.0, w.partial);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(side.syrk[k], side.s));
    }
    if (i > 0) CUDA_TRY(cudaStreamWaitEvent(st, side.syrk[(i - 1) & 1], 0));
    const int total = (c + 1) * (c + 2) / 2;
    big::big_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(w.partial, big::BIG_NZ, c, (double)B.nv * (double)N, stats);
    CUDA_TRY(cudaGetLastError());
    if (info) {
        add_rankloss_kernel<<<1, 1, 0, st>>>(w.rankloss, (long long*)info);
        CUDA_TRY(cudaGetLastError());
    }
    return SYSID_OK;
}

size_t sysid_gram_from_stack_workspace_bytes(int32_t c) {
    (void)c;
    int sms = 0;
    if (device_sm_count(&sms) != SYSID_OK) return 0;
    return sizeof(double) * (size_t)sms * PARTIAL_DOUBLES;
}

static int gram_accumulate_impl(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                                const double* tau, const double* contact, int64_t N, int64_t ld, const double* weights,
                                int32_t friction, double* stats, int64_t* info, void* workspace, size_t workspace_bytes,
                                int reserve_sms, void* stream, bool keep_partials = false);

int sysid_gram_accumulate(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                          const double* tau, const double* contact, int64_t N, int64_t ld, const double* weights,
                          int32_t friction, double* stats, int64_t* info, void* workspace, size_t workspace_bytes,
                          void* stream) {
    return gram_accumulate_impl(model, q, dq, ddq, tau, contact, N, ld, weights, friction, stats, info, workspace, workspace_bytes, 0, stream);
}

// reserve_sms: leave that many SMs free (the persistent CTAs of this kernel take a whole SM each; a concurrent single-CTA
// LMI pre-solve on another stream needs one)
static int gram_accumulate_impl(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                                const double* tau, const double* contact, int64_t N, int64_t ld, const double* weights,
                                int32_t friction, double* stats, int64_t* info, void* workspace, size_t workspace_bytes,
                                int reserve_sms, void* stream, bool keep_partials) {
    // keep_partials (structured kernel only; the caller zeroed the partial Grams first): this launch ADDS to the partial Grams in
    // the workspace and nothing is reduced -- gram_reduce_partials does that once for all chunks of a streamed log
    if (!model || !q || !dq || !ddq || !tau || !stats || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (model->dev.n_ee > 0 && !contact) return fail(SYSID_ERR_INVALID, "null contact array");
    if (N < 0 || ld < N) return fail(SYSID_ERR_INVALID, "bad N/ld");
    if (N == 0) return SYSID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (model->big) return big_gram(model, make_io(q, dq, ddq, tau, contact, weights, ld), N, friction ? 1 : 0, stats, info, workspace, workspace_bytes, st);
    const DevModel& M = model->dev;
    const long long nsb = use_struct(model) ? (N + ST_SB - 1) / ST_SB : (N + FSB - 1) / FSB;
    static const int debug_reserve = [] { const char* e = std::getenv("SYSID_DEBUG_RESERVE_SMS"); return e ? std::atoi(e) : 0; }();   // diagnostic
    if (debug_reserve > reserve_sms) reserve_sms = debug_reserve;
    const int max_ctas = (model->sm_count - reserve_sms > 1) ? model->sm_count - reserve_sms : 1;
    const int grid = (int)(nsb < max_ctas ? nsb : max_ctas);
    if (workspace_bytes < sizeof(double) * (size_t)grid * PARTIAL_DOUBLES)
        return fail(SYSID_ERR_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, sizeof(double) * (size_t)grid * PARTIAL_DOUBLES);
    const bool structured = use_struct(model);
    static const bool debug_kernel = std::getenv("SYSID_DEBUG_KERNEL") != nullptr;
    if (debug_kernel) fprintf(stderr, "[sysid] gram kernel: %s (st_ok %d)\n", structured ? "structured" : "unstructured", (int)model->dev.st_ok);
    GramArgs a{};
    a.io = make_io(q, dq, ddq, tau, contact, weights, ld);
    a.N = N; a.friction = friction ? 1 : 0; a.partial = (double*)workspace; a.seg_len = 0; a.accumulate = 0;
    if (keep_partials && !structured) return fail(SYSID_ERR_INVALID, "keep_partials needs the structured kernel");
    if (structured) {
        int rc = opt_in_smem(gram_struct_kernel<false>, ST_SMEM_BYTES);
        if (rc) return rc;
        a.accumulate = keep_partials ? 1 : 0;
        gram_struct_kernel<false><<<grid, GRAM_THREADS, ST_SMEM_BYTES, st>>>(M, a);
        if (keep_partials) { CUDA_TRY(cudaGetLastError()); return SYSID_OK; }
    } else {
        int rc = opt_in_smem(gram_fused_kernel<false>, GRAM_SMEM_BYTES);
        if (rc) return rc;
        gram_fused_kernel<false><<<grid, GRAM_THREADS, GRAM_SMEM_BYTES, st>>>(M, a);
    }
    CUDA_TRY(cudaGetLastError());
    const int c = M.nparams + (friction ? 2 * M.nd : 0);
    const int total = (c + 1) * (c + 2) / 2;
    gram_reduce_kernel<<<(total * REDUCE_LANES + 255) / 256, 256, 0, st>>>((const double*)workspace, grid, c, (double)M.nv, 0.0, stats,
                                                             (long long*)info, make_colmap(M, friction ? 1 : 0, structured));
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

// Reduction of the partial Grams a streamed log accumulated with keep_partials (ADDS into stats, like every reduction here).
static int gram_reduce_partials(const sysid_model* model, int32_t friction, double* stats, int64_t* info, void* workspace, cudaStream_t st) {
    const DevModel& M = model->dev;
    const int c = M.nparams + (friction ? 2 * M.nd : 0);
    const int total = (c + 1) * (c + 2) / 2;
    gram_reduce_kernel<<<(total * REDUCE_LANES + 255) / 256, 256, 0, st>>>((const double*)workspace, model->sm_count, c, (double)M.nv, 0.0, stats,
                                                             (long long*)info, make_colmap(M, friction ? 1 : 0, true));
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

// ---- block bootstrap (BASELINE configs[4]): one statistics vector per block of consecutive samples, ONE launch -----------------
size_t sysid_gram_blocks_workspace_bytes(const sysid_model* model, int64_t N, int64_t block) {
    if (!model || N < 0 || block <= 0) return 0;
    const size_t nseg = (size_t)((N + block - 1) / block);
    return sizeof(double) * (nseg > 0 ? nseg : 1) * PARTIAL_DOUBLES;
}

int sysid_gram_blocks(const sysid_model* model, const double* q, const double* dq, const double* ddq, const double* tau,
                      const double* contact, int64_t N, int64_t ld, int64_t block, int32_t friction, double* stats_blocks,
                      int64_t stats_stride, int64_t* info, void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !q || !dq || !ddq || !tau || !stats_blocks || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (model->dev.n_ee > 0 && !contact) return fail(SYSID_ERR_INVALID, "null contact array");
    if (N < 0 || ld < N || block <= 0) return fail(SYSID_ERR_INVALID, "bad N/ld/block");
    if (model->big) return fail(SYSID_ERR_UNSUPPORTED, "per-block statistics are not available on the large-model path");
    if (N == 0) return SYSID_OK;
#if !defined(SYSID_PARK_FILL)
    return fail(SYSID_ERR_UNSUPPORTED, "segmented statistics need the default accumulator parking policy");
#endif
    const DevModel& M = model->dev;
    const size_t slen = sysid_stats_len(model, friction);
    if (stats_stride < (int64_t)slen) return fail(SYSID_ERR_INVALID, "stats_stride smaller than the statistics vector");
    const long long nseg = (N + block - 1) / block;
    if (workspace_bytes < sysid_gram_blocks_workspace_bytes(model, N, block)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const bool structured = use_struct(model);
    GramArgs a{};
    a.io = make_io(q, dq, ddq, tau, contact, nullptr, ld);
    a.N = N; a.friction = friction ? 1 : 0; a.partial = (double*)workspace; a.seg_len = block; a.accumulate = 0;
    const int grid = (int)(nseg < model->sm_count ? nseg : model->sm_count);
    if (structured) {
        int rc = opt_in_smem(gram_struct_kernel<true>, ST_SMEM_BYTES);
        if (rc) return rc;
        gram_struct_kernel<true><<<grid, GRAM_THREADS, ST_SMEM_BYTES, st>>>(M, a);
    } else {
        int rc = opt_in_smem(gram_fused_kernel<true>, GRAM_SMEM_BYTES);
        if (rc) return rc;
        gram_fused_kernel<true><<<grid, GRAM_THREADS, GRAM_SMEM_BYTES, st>>>(M, a);
    }
    CUDA_TRY(cudaGetLastError());
    const int c = M.nparams + (friction ? 2 * M.nd : 0);
    const int total = (c + 1) * (c + 2) / 2;
    CUDA_TRY(cudaMemsetAsync(stats_blocks, 0, sizeof(double) * (size_t)stats_stride * (size_t)nseg, st));
    for (long long y0 = 0; y0 < nseg; y0 += 65535) {
        const unsigned ny = (unsigned)((nseg - y0 < 65535) ? (nseg - y0) : 65535);
        gram_reduce_kernel<<<dim3((total * REDUCE_LANES + 255) / 256, ny), 256, 0, st>>>((const double*)workspace + (size_t)y0 * PARTIAL_DOUBLES, 1, c, (double)M.nv, 0.0,
                                                                        stats_blocks + (size_t)y0 * stats_stride, (long long*)info,
                                                                        make_colmap(M, friction ? 1 : 0, structured), (long long)stats_stride);
        CUDA_TRY(cudaGetLastError());
    }
    return SYSID_OK;
}

int sysid_combine_stats(const double* weights, int64_t B, int64_t K, const double* stats_blocks, int64_t slen, double* out, void* stream) {
    if (!weights || !stats_blocks || !out) return fail(SYSID_ERR_INVALID, "null argument");
    if (B < 0 || K <= 0 || slen <= 0 || slen % 8 != 0) return fail(SYSID_ERR_INVALID, "bad B/K/slen (slen must be a multiple of 8)");
    if (B == 0) return SYSID_OK;
    const unsigned gx = (unsigned)((slen + 64 * COMBINE_WARPS - 1) / (64 * COMBINE_WARPS));
    for (int64_t b0 = 0; b0 < B; b0 += 8 * 65535) {
        const unsigned gy = (unsigned)(((B - b0 < 8 * 65535 ? B - b0 : 8 * 65535) + 7) / 8);
        combine_stats_kernel<<<dim3(gx, gy), 32 * COMBINE_WARPS, 0, (cudaStream_t)stream>>>(weights + b0 * K, B - b0, K, stats_blocks, slen, out + b0 * slen);
        CUDA_TRY(cudaGetLastError());
    }
    return SYSID_OK;
}

size_t sysid_gram_host_workspace_bytes(const sysid_model* model, int64_t chunk) {
    if (!model || chunk <= 0) return 0;
    const DevModel& M = model->dev;
    const size_t per_sample = (size_t)M.nq + 2 * (size_t)M.nv + (size_t)M.nd + (size_t)M.n_ee + 1;
    // two fp64 staging buffers + two float32 landing buffers (float32 host arrays are widened on the device)
    return sysid_gram_workspace_bytes(model) + 2 * (sizeof(double) + sizeof(float)) * per_sample * (size_t)chunk + 512;
}

namespace {
__global__ void copy_f64_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dst[e] = src[e];
}
// float32 landing buffer -> fp64 staging buffer (exact widening, what numpy does when the reference mixes its float32 q /
// contact arrays into fp64 arithmetic): channels x n, both with leading dimension ld
__global__ void widen_f32_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t n, int64_t ld, int channels) {
    const int64_t total = (int64_t)channels * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ch = e / n, i = e - ch * n;
        dst[ch * ld + i] = (double)src[ch * ld + i];
    }
}
}  // namespace

int sysid_gram_accumulate_host_ex(const sysid_model* model, const void* const* arrays_host, const int32_t* dtypes,
                                  const int64_t* lds_host, int64_t N, const double* weights_host, int32_t friction,
                                  double* stats, int64_t* info, void* workspace, size_t workspace_bytes, int64_t chunk,
                                  void* stream) {
    return sysid_gram_accumulate_host_presolve(model, arrays_host, dtypes, lds_host, N, weights_host, friction, stats, info,
                                               workspace, workspace_bytes, chunk, nullptr, stream);
}

int sysid_gram_accumulate_host_presolve(const sysid_model* model, const void* const* arrays_host, const int32_t* dtypes,
                                        const int64_t* lds_host, int64_t N, const double* weights_host, int32_t friction,
                                        double* stats, int64_t* info, void* workspace, size_t workspace_bytes, int64_t chunk,
                                        const sysid_presolve* pre, void* stream) {
    if (!model || !arrays_host || !dtypes || !lds_host || !stats || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (pre && (!pre->desc || !pre->plan || !pre->sdp_workspace || !pre->stats_snapshot || !pre->x_scratch || !pre->info_scratch || !pre->warm_out))
        return fail(SYSID_ERR_INVALID, "null presolve argument");
    if (pre && pre->refine_at > 0 && !pre->stats_snapshot2) return fail(SYSID_ERR_INVALID, "refine_at needs stats_snapshot2");
    const DevModel& M = model->dev;
    const int channels[5] = {M.nq, M.nv, M.nv, M.nd, M.n_ee};
    for (int a = 0; a < 5; ++a) {
        if (channels[a] > 0 && !arrays_host[a]) return fail(SYSID_ERR_INVALID, a == 4 ? "null contact array" : "null argument");
        if (dtypes[a] != SYSID_DTYPE_F64 && dtypes[a] != SYSID_DTYPE_F32) return fail(SYSID_ERR_INVALID, "dtype must be SYSID_DTYPE_F64 or SYSID_DTYPE_F32");
        if (lds_host[a] < N) return fail(SYSID_ERR_INVALID, "bad N/ld/chunk");
    }
    if (N < 0 || chunk <= 0) return fail(SYSID_ERR_INVALID, "bad N/ld/chunk");
    if (N == 0) return SYSID_OK;
    if (workspace_bytes < sysid_gram_host_workspace_bytes(model, chunk)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t gram_ws = (sysid_gram_workspace_bytes(model) + 255) & ~(size_t)255;
    const size_t per_sample = (size_t)M.nq + 2 * (size_t)M.nv + (size_t)M.nd + (size_t)M.n_ee + 1;
    double* stage[2] = {(double*)((char*)workspace + gram_ws), (double*)((char*)workspace + gram_ws) + per_sample * (size_t)chunk};
    float* land[2] = {(float*)(stage[1] + per_sample * (size_t)chunk), (float*)(stage[1] + per_sample * (size_t)chunk) + per_sample * (size_t)chunk};
    // streams / events: the handle's own set when free, else a per-call set
    std::unique_lock<std::mutex> lk(model->host_mu, std::try_to_lock);
    HostStreamRes local;
    HostStreamRes* R = nullptr;
    if (lk.owns_lock()) {
        if (!model->host_res.ok && !model->host_res.create()) { model->host_res.destroy(); return fail(SYSID_ERR_CUDA, "creating the internal streams failed"); }
        R = &model->host_res;
    } else {
        if (!local.create()) { local.destroy(); return fail(SYSID_ERR_CUDA, "creating the internal streams failed"); }
        R = &local;
    }
    cudaStream_t cp = R->s_copy, sv = R->s_solve;
    cudaEvent_t* copied = &R->ev[0]; cudaEvent_t* consumed = &R->ev[2];
    cudaEvent_t start = R->ev[4], snap = R->ev[5], solved = R->ev[6], snap2 = R->ev[7];
    int rc = SYSID_OK;
    auto cleanup = [&]() { if (R == &local) local.destroy(); };
    // LMI pre-solve (optional): the statistics are additive, so the fit of the first pre->samples samples is a point (and a set
    // of multipliers) within the statistical noise of the final fit.  It runs as ONE thread block on an internal stream beside the
    // remaining chunks, which leave it one SM; its result (pre->warm_out) warm-starts the final sysid_sdp_solve_plan.
    const size_t warm_n = pre ? sdp_warm_doubles(pre->desc->num_links, pre->desc->ndof) : 0;
    int64_t first = chunk;
    bool presolve = false;
    if (pre) {
        first = (pre->samples > 0 && pre->samples < chunk) ? pre->samples : chunk;
        presolve = (N >= 2 * first);
    }
    bool presolve_running = false, refined = false;
#define HOST_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { rc = fail(SYSID_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); cleanup(); return rc; } } while (0)
    if (pre) HOST_TRY(cudaMemsetAsync(pre->warm_out, 0, sizeof(double) * warm_n, st));      // "no record" until the pre-solve has written one
    // structured kernel: the chunks add up in the per-CTA partial Grams and are reduced ONCE (a reduction per chunk reads 16 MB of
    // partials: 70 us x ~14 chunks of a 1 M-sample log); the pre-solve snapshots are reductions of the partials so far
    const bool keep = use_struct(model) && !model->big;
    if (keep) HOST_TRY(cudaMemsetAsync(workspace, 0, sizeof(double) * (size_t)model->sm_count * PARTIAL_DOUBLES, st));
    // the staging buffers may still be read by earlier work on `stream`
    HOST_TRY(cudaEventRecord(start, st));
    HOST_TRY(cudaStreamWaitEvent(cp, start, 0));
    long long k = 0;
    int64_t want = 0;
    for (int64_t lo = 0; lo < N; ++k) {
        const int b = (int)(k & 1);
        // chunk sizes ramp up from 16 384 samples by x1.5: the first kernel starts after 9 MB have crossed PCIe instead of 72 MB, and
        // with two staging buffers the copy of chunk k+2 (which may only start when chunk k is consumed) stays hidden behind the
        // kernel of chunk k+1 as long as it is < 1.77x longer (552 B/sample at ~53 GB/s against ~18.4 ns/sample of compute)
        if (k == 0) want = 16384; else want += want / 2;
        if (want > chunk) want = chunk;
        // whole waves: every persistent CTA of the launch gets the same number of super-batches (no tail of one super-batch, which
        // at 37 super-batches per CTA is 2.7 % of the launch)
        const int64_t sbs = use_struct(model) ? ST_SB : FSB;       // samples per super-batch of the kernel that will run
        const int64_t wave = sbs * (model->sm_count - (presolve_running ? 1 : 0));
        int64_t take = (want >= 8 * wave) ? want - want % wave : want - want % sbs;
        if (take <= 0) take = want;
        const int64_t n = (N - lo < take) ? (N - lo) : take;
        if (k >= 2) HOST_TRY(cudaStreamWaitEvent(cp, consumed[b], 0));
        double* dst[6];
        size_t off = 0;
        for (int a = 0; a < 5; ++a) { dst[a] = stage[b] + off * (size_t)chunk; off += (size_t)channels[a]; }
        dst[5] = stage[b] + off * (size_t)chunk;
        for (int a = 0; a < 5; ++a) {
            if (channels[a] == 0) continue;
            if (dtypes[a] == SYSID_DTYPE_F64) {
                HOST_TRY(cudaMemcpy2DAsync(dst[a], chunk * sizeof(double), (const double*)arrays_host[a] + lo, lds_host[a] * sizeof(double),
                                           n * sizeof(double), channels[a], cudaMemcpyHostToDevice, cp));
            } else {
                float* l = land[b] + (dst[a] - stage[b]);
                HOST_TRY(cudaMemcpy2DAsync(l, chunk * sizeof(float), (const float*)arrays_host[a] + lo, lds_host[a] * sizeof(float),
                                           n * sizeof(float), channels[a], cudaMemcpyHostToDevice, cp));
            }
        }
        if (weights_host) HOST_TRY(cudaMemcpyAsync(dst[5], weights_host + lo, n * sizeof(double), cudaMemcpyHostToDevice, cp));
        HOST_TRY(cudaEventRecord(copied[b], cp));
        HOST_TRY(cudaStreamWaitEvent(st, copied[b], 0));
        for (int a = 0; a < 5; ++a)
            if (channels[a] > 0 && dtypes[a] == SYSID_DTYPE_F32) {
                const int64_t total = (int64_t)channels[a] * n;
                const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
                widen_f32_kernel<<<blocks, 256, 0, st>>>(land[b] + (dst[a] - stage[b]), dst[a], n, chunk, channels[a]);
                HOST_TRY(cudaGetLastError());
            }
        rc = gram_accumulate_impl(model, dst[0], dst[1], dst[2], dst[3], M.n_ee > 0 ? dst[4] : nullptr, n, chunk,
                                  weights_host ? dst[5] : nullptr, friction, stats, info, workspace, gram_ws, presolve_running ? 1 : 0, st, keep);
        if (rc != SYSID_OK) { cleanup(); return rc; }
        HOST_TRY(cudaEventRecord(consumed[b], st));
        if (presolve && !presolve_running && lo + n >= first) {
            const size_t slen = sysid_stats_len(model, friction);
            // snapshot by a kernel, not cudaMemcpyAsync: a device-to-device copy would queue on a copy engine behind the 72 MB
            // host-to-device transfer of the next chunk and stall `stream` for more than a millisecond
            if (keep) {
                HOST_TRY(cudaMemsetAsync(pre->stats_snapshot, 0, sizeof(double) * slen, st));
                rc = gram_reduce_partials(model, friction, pre->stats_snapshot, nullptr, workspace, st);
                if (rc != SYSID_OK) { cleanup(); return rc; }
            } else {
                copy_f64_kernel<<<(unsigned)((slen + 255) / 256), 256, 0, st>>>(stats, pre->stats_snapshot, (int64_t)slen);
                HOST_TRY(cudaGetLastError());
            }
            HOST_TRY(cudaEventRecord(snap, st));
            HOST_TRY(cudaStreamWaitEvent(sv, snap, 0));
            char msg[256] = "";
            sysid_sdp_desc d1 = *pre->desc;
            if (pre->first_tol > 0.0) d1.tol = pre->first_tol;
            rc = sdp_solve_planned(d1, (const double*)pre->plan, pre->stats_snapshot, (int64_t)slen, 1, pre->x_scratch,
                                   pre->info_scratch, pre->sdp_workspace, pre->sdp_workspace_bytes, nullptr, pre->warm_out, sv, msg, sizeof(msg));
            if (rc != SYSID_OK) { fail(rc, "pre-solve: %s", msg); cleanup(); return rc; }
            HOST_TRY(cudaEventRecord(solved, sv));
            presolve_running = true;
        } else if (presolve_running && !refined && pre->refine_at > 0 && lo + n >= pre->refine_at && lo + n < N) {
            // second stage: the statistics so far, solved behind the first pre-solve (same internal stream) from its record; the
            // record it leaves is the start of the final solve.  Same scratch buffers: the stream orders the two solves.
            const size_t slen = sysid_stats_len(model, friction);
            if (keep) {
                HOST_TRY(cudaMemsetAsync(pre->stats_snapshot2, 0, sizeof(double) * slen, st));
                rc = gram_reduce_partials(model, friction, pre->stats_snapshot2, nullptr, workspace, st);
                if (rc != SYSID_OK) { cleanup(); return rc; }
            } else {
                copy_f64_kernel<<<(unsigned)((slen + 255) / 256), 256, 0, st>>>(stats, pre->stats_snapshot2, (int64_t)slen);
                HOST_TRY(cudaGetLastError());
            }
            HOST_TRY(cudaEventRecord(snap2, st));
            HOST_TRY(cudaStreamWaitEvent(sv, snap2, 0));
            char msg[256] = "";
            rc = sdp_solve_planned(*pre->desc, (const double*)pre->plan, pre->stats_snapshot2, (int64_t)slen, 1, pre->x_scratch,
                                   pre->info_scratch, pre->sdp_workspace, pre->sdp_workspace_bytes, pre->warm_out, pre->warm_out, sv, msg, sizeof(msg));
            if (rc != SYSID_OK) { fail(rc, "pre-solve (second stage): %s", msg); cleanup(); return rc; }
            HOST_TRY(cudaEventRecord(solved, sv));
            refined = true;
        }
        lo += n;
    }
    if (keep) { rc = gram_reduce_partials(model, friction, stats, info, workspace, st); if (rc != SYSID_OK) { cleanup(); return rc; } }
    if (presolve_running) HOST_TRY(cudaStreamWaitEvent(st, solved, 0));      // later work on `stream` sees the warm-start record
#undef HOST_TRY
    cleanup();
    return SYSID_OK;
}

int sysid_gram_accumulate_host(const sysid_model* model, const double* q_host, const double* dq_host, const double* ddq_host,
                               const double* tau_host, const double* contact_host, int64_t N, int64_t ld_host,
                               const double* weights_host, int32_t friction, double* stats, int64_t* info,
                               void* workspace, size_t workspace_bytes, int64_t chunk, void* stream) {
    if (!model || !q_host || !dq_host || !ddq_host || !tau_host) return fail(SYSID_ERR_INVALID, "null argument");
    const void* arrays[5] = {q_host, dq_host, ddq_host, tau_host, contact_host};
    const int32_t dt[5] = {SYSID_DTYPE_F64, SYSID_DTYPE_F64, SYSID_DTYPE_F64, SYSID_DTYPE_F64, SYSID_DTYPE_F64};
    const int64_t lds[5] = {ld_host, ld_host, ld_host, ld_host, ld_host};
    return sysid_gram_accumulate_host_ex(model, arrays, dt, lds, N, weights_host, friction, stats, info, workspace,
                                         workspace_bytes, chunk, stream);
}

int sysid_gram_from_stack(const double* A, const double* b, int64_t rows, int32_t c, double* stats,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (!A || !b || !stats || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (c < 1 || c + 1 > CW) return fail(SYSID_ERR_UNSUPPORTED, "c = %d outside [1, %d]", c, CW - 1);
    if (rows < 0) return fail(SYSID_ERR_INVALID, "bad rows");
    if (rows == 0) return SYSID_OK;
    int sms = 0;
    int rc = device_sm_count(&sms);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long nchunks = (rows + TILE_ROWS - 1) / TILE_ROWS;
    const int grid = (int)(nchunks < sms ? nchunks : sms);
    if (workspace_bytes < sizeof(double) * (size_t)grid * PARTIAL_DOUBLES) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    const size_t smem = sizeof(double) * TILE_DOUBLES;
    rc = opt_in_smem(gram_stack_kernel, smem);
    if (rc) return rc;
    StackArgs a{A, b, rows, c, (double*)workspace};
    gram_stack_kernel<<<grid, GRAM_THREADS, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    const int total = (c + 1) * (c + 2) / 2;
    gram_reduce_kernel<<<(total * REDUCE_LANES + 255) / 256, 256, 0, st>>>((const double*)workspace, grid, c, 0.0, (double)rows, stats, nullptr, make_colmap(DevModel{}, 0, false));
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

size_t sysid_predict_rmse_workspace_bytes(const sysid_model* model) {
    if (!model) return 0;
    if (model->big) return big_workspace_bytes(model);
    return sizeof(double) * (size_t)model->sm_count * RMSE_PARTIAL;
}

int sysid_predict_rmse(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                       const double* tau, const double* contact, int64_t N, int64_t ld, const double* phi,
                       double* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !q || !dq || !ddq || !tau || !phi || !out || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (model->dev.n_ee > 0 && !contact) return fail(SYSID_ERR_INVALID, "null contact array");
    if (N <= 0 || ld < N) return fail(SYSID_ERR_INVALID, "bad N/ld");
    cudaStream_t st = (cudaStream_t)stream;
    if (model->big) {
        if (workspace_bytes < big_workspace_bytes(model)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
        const big::BigModel& B = model->bm;
        const big::BigWs w = big::big_workspace(B, (void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255));
        const SampleIO io = make_io(q, dq, ddq, tau, contact, nullptr, ld);
        double* sums = w.e2 + (size_t)big::BIG_CHUNK * big::BV;
        CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * big::BV, st));
        const int ctau = B.nparams;                     // no friction columns in this pass: the torque column follows the body columns
        for (int64_t lo = 0; lo < N; lo += big::BIG_CHUNK) {
            const int ns = (int)((N - lo < big::BIG_CHUNK) ? (N - lo) : big::BIG_CHUNK);
            int rc = big_chunk_rows(model, io, lo, ns, 0, true, w, st);
            if (rc != SYSID_OK) return rc;
            big::big_err_kernel<<<(ns * B.nd + 127) / 128, 128, 0, st>>>(B, ns, ctau, w.Yt, w.W, w.Z, w.m3, phi, w.e2);
            big::big_err_sum_kernel<<<1, 64, 0, st>>>(ns, B.nd, w.e2, sums);
            CUDA_TRY(cudaGetLastError());
        }
        big::big_err_final_kernel<<<1, 32, 0, st>>>(B.nd, N, sums, out);
        CUDA_TRY(cudaGetLastError());
        return SYSID_OK;
    }
    const long long nsb = (N + RSB - 1) / RSB;
    const int grid = (int)(nsb < model->sm_count ? nsb : model->sm_count);
    if (workspace_bytes < sizeof(double) * (size_t)grid * RMSE_PARTIAL) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    int rc = opt_in_smem(rmse_kernel, RMSE_SMEM_BYTES);
    if (rc) return rc;
    RmseArgs a{};
    a.io = make_io(q, dq, ddq, tau, contact, nullptr, ld);
    a.N = N; a.phi = phi; a.partial = (double*)workspace;
    rmse_kernel<<<grid, GRAM_THREADS, RMSE_SMEM_BYTES, st>>>(model->dev, a);
    CUDA_TRY(cudaGetLastError());
    rmse_finalize_kernel<<<1, 32, 0, st>>>((const double*)workspace, grid, model->dev.nd, N, out);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

// ------------------------------------------------------------------------------------------------ pre-processing filters
size_t sysid_filtfilt_workspace_bytes(int32_t channels, int64_t N, int32_t ncoef) {
    if (channels <= 0 || N <= 0 || ncoef < 1) return 0;
    const int64_t next = N + 2 * 3 * (int64_t)ncoef;
    const int64_t nchunks = (next + FILT_CHUNK - 1) / FILT_CHUNK;
    return sizeof(double) * ((size_t)channels * (size_t)next + 2 * (size_t)channels * (size_t)nchunks * FILT_MAXS) + 256;
}

int sysid_filtfilt(const double* b_host, int32_t nb, const double* a_host, int32_t na, const double* x, double* y,
                   int32_t channels, int64_t N, int64_t ld, int32_t pad_float32, void* workspace, size_t workspace_bytes,
                   void* stream) {
    if (!b_host || !a_host || !x || !y || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (nb < 1 || na < 1 || nb > FILT_MAXS + 1 || na > FILT_MAXS + 1) return fail(SYSID_ERR_UNSUPPORTED, "filter longer than %d coefficients", FILT_MAXS + 1);
    if (a_host[0] == 0.0) return fail(SYSID_ERR_INVALID, "a[0] must be non-zero");
    if (channels < 0 || N < 0 || ld < N) return fail(SYSID_ERR_INVALID, "bad channels/N/ld");
    const int ntaps = nb > na ? nb : na, padlen = 3 * ntaps;
    if (N <= padlen) return fail(SYSID_ERR_INVALID, "The length of the input vector x must be greater than padlen, which is %d.", padlen);
    if (channels == 0) return SYSID_OK;
    if (workspace_bytes < sysid_filtfilt_workspace_bytes(channels, N, ntaps)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    FiltCoef c;
    std::memset(&c, 0, sizeof(c));
    for (int k = 0; k < nb; ++k) c.b[k] = b_host[k] / a_host[0];
    for (int k = 0; k < na; ++k) c.a[k] = a_host[k] / a_host[0];
    const int n = ntaps - 1;
    // steady state of the step response (scipy.signal.lfilter_zi): (I - companion(a)^T) zi = b[1:] - a[1:] b[0], solved in closed form
    if (n > 0) {
        double bsum = 0.0, colsum = 1.0;
        for (int k = 1; k <= n; ++k) { bsum += c.b[k] - c.a[k] * c.b[0]; colsum += c.a[k]; }
        c.zi[0] = bsum / colsum;
        double asum = 1.0, csum = 0.0;
        for (int k = 1; k < n; ++k) {
            asum += c.a[k];
            csum += c.b[k] - c.a[k] * c.b[0];
            c.zi[k] = asum * c.zi[0] - csum;
        }
    }
    {   // Phi^FILT_CHUNK by repeated squaring in extended precision (FILT_CHUNK is a power of two)
        static_assert((FILT_CHUNK & (FILT_CHUNK - 1)) == 0, "power of two");
        long double P[FILT_MAXS][FILT_MAXS] = {}, Q[FILT_MAXS][FILT_MAXS];
        for (int k = 0; k < FILT_MAXS; ++k) { P[k][0] = -(long double)c.a[k + 1]; if (k + 1 < FILT_MAXS) P[k][k + 1] = 1.0L; }
        for (int sq = 1; sq < FILT_CHUNK; sq <<= 1) {
            for (int i = 0; i < FILT_MAXS; ++i) for (int j = 0; j < FILT_MAXS; ++j) { long double t = 0; for (int k = 0; k < FILT_MAXS; ++k) t += P[i][k] * P[k][j]; Q[i][j] = t; }
            std::memcpy(P, Q, sizeof(P));
        }
        for (int i = 0; i < FILT_MAXS; ++i) for (int j = 0; j < FILT_MAXS; ++j) c.phiL[i][j] = (double)P[i][j];
        // Phi^(FILT_CHUNK seg) for the segments of the chaining kernel: binary powering of Phi^FILT_CHUNK
        const long long next = N + 2 * (long long)padlen;
        const int nchunks = (int)((next + FILT_CHUNK - 1) / FILT_CHUNK);
        int seg = (nchunks + FILT_SCAN_THREADS - 1) / FILT_SCAN_THREADS;
        long double R[FILT_MAXS][FILT_MAXS] = {};
        for (int i = 0; i < FILT_MAXS; ++i) R[i][i] = 1.0L;
        for (int e = seg; e > 0; e >>= 1) {
            if (e & 1) {
                for (int i = 0; i < FILT_MAXS; ++i) for (int j = 0; j < FILT_MAXS; ++j) { long double t = 0; for (int k = 0; k < FILT_MAXS; ++k) t += R[i][k] * P[k][j]; Q[i][j] = t; }
                std::memcpy(R, Q, sizeof(R));
            }
            for (int i = 0; i < FILT_MAXS; ++i) for (int j = 0; j < FILT_MAXS; ++j) { long double t = 0; for (int k = 0; k < FILT_MAXS; ++k) t += P[i][k] * P[k][j]; Q[i][j] = t; }
            std::memcpy(P, Q, sizeof(P));
        }
        for (int i = 0; i < FILT_MAXS; ++i) for (int j = 0; j < FILT_MAXS; ++j) c.phiM[i][j] = (double)R[i][j];
    }
    cudaStream_t st = (cudaStream_t)stream;
    FiltArgs g{};
    g.N = N; g.ld = ld; g.Next = N + 2 * (long long)padlen; g.channels = channels; g.padlen = padlen;
    g.pad_float32 = pad_float32 ? 1 : 0;
    g.nchunks = (int)((g.Next + FILT_CHUNK - 1) / FILT_CHUNK);
    g.seg = (g.nchunks + FILT_SCAN_THREADS - 1) / FILT_SCAN_THREADS;
    double* Y1 = (double*)workspace;
    g.fstate = Y1 + (size_t)channels * g.Next;
    g.sstate = g.fstate + (size_t)channels * g.nchunks * FILT_MAXS;
    const long long items = (long long)channels * g.nchunks;
    const unsigned grid = (unsigned)((items + FILT_THREADS - 1) / FILT_THREADS);
    for (int dir = 0; dir < 2; ++dir) {
        g.backward = dir;
        g.x = dir ? Y1 : x;
        g.y = dir ? y : Y1;
        filt_chunk_kernel<0><<<grid, FILT_THREADS, 0, st>>>(c, g);
        filt_scan_kernel<<<channels, FILT_SCAN_THREADS, 0, st>>>(c, g);
        filt_chunk_kernel<1><<<grid, FILT_THREADS, 0, st>>>(c, g);
        CUDA_TRY(cudaGetLastError());
    }
    return SYSID_OK;
}

namespace {
// rows of E = V_eval pinv(V_fit) for a degree-p polynomial fitted to W equally spaced samples: row t gives the fitted
// value at window position pos[t] as a linear combination of the W samples.  Abscissae are centred and scaled to
// [-1, 1] (the fitted polynomial does not depend on the parametrisation); QR by modified Gram-Schmidt, twice, in
// extended precision.
void polyfit_rows(int W, int p, const int* pos, int npos, double* rows /* npos x W */) {
    const int m = p + 1, half = W / 2;
    std::vector<long double> Q((size_t)W * m), R((size_t)m * m, 0.0L);
    for (int j = 0; j < W; ++j) { long double u = (long double)(j - half) / (long double)(half > 0 ? half : 1), pw = 1.0L; for (int k = 0; k < m; ++k) { Q[(size_t)j * m + k] = pw; pw *= u; } }
    for (int k = 0; k < m; ++k) {
        for (int pass = 0; pass < 2; ++pass)
            for (int i = 0; i < k; ++i) {
                long double d = 0; for (int j = 0; j < W; ++j) d += Q[(size_t)j * m + i] * Q[(size_t)j * m + k];
                for (int j = 0; j < W; ++j) Q[(size_t)j * m + k] -= d * Q[(size_t)j * m + i];
                R[(size_t)i * m + k] += d;
            }
        long double nr = 0; for (int j = 0; j < W; ++j) nr += Q[(size_t)j * m + k] * Q[(size_t)j * m + k];
        nr = sqrtl(nr);
        R[(size_t)k * m + k] = nr;
        for (int j = 0; j < W; ++j) Q[(size_t)j * m + k] /= nr;
    }
    for (int t = 0; t < npos; ++t) {
        // w solves R^T w = v(pos[t]);  row = Q w
        long double v[32], w[32];
        long double u = (long double)(pos[t] - half) / (long double)(half > 0 ? half : 1), pw = 1.0L;
        for (int k = 0; k < m; ++k) { v[k] = pw; pw *= u; }
        for (int k = 0; k < m; ++k) { long double sres = v[k]; for (int i = 0; i < k; ++i) sres -= R[(size_t)i * m + k] * w[i]; w[k] = sres / R[(size_t)k * m + k]; }
        for (int j = 0; j < W; ++j) { long double sres = 0; for (int k = 0; k < m; ++k) sres += Q[(size_t)j * m + k] * w[k]; rows[(size_t)t * W + j] = (double)sres; }
    }
}
}  // namespace

size_t sysid_savgol_workspace_bytes(int32_t window_length) {
    if (window_length < 1) return 0;
    return sizeof(double) * (size_t)window_length * (size_t)(window_length + 1) + 256;
}

int sysid_savgol(int32_t window_length, int32_t polyorder, const double* x, double* y, int32_t channels, int64_t N,
                 int64_t ld, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !y || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    if (x == y) return fail(SYSID_ERR_INVALID, "savgol cannot run in place");
    if (window_length < 1 || window_length % 2 == 0 || window_length > SG_MAXW) return fail(SYSID_ERR_UNSUPPORTED, "window_length must be odd and <= %d", SG_MAXW);
    if (polyorder < 0 || polyorder >= window_length || polyorder > 15) return fail(SYSID_ERR_INVALID, "polyorder must be less than window_length.");
    if (channels < 0 || N < 0 || ld < N) return fail(SYSID_ERR_INVALID, "bad channels/N/ld");
    if (N < window_length) return fail(SYSID_ERR_INVALID, "If mode is 'interp', window_length must be less than or equal to the size of x.");
    if (channels == 0 || N == 0) return SYSID_OK;
    if (workspace_bytes < sysid_savgol_workspace_bytes(window_length)) return fail(SYSID_ERR_WORKSPACE, "workspace too small");
    const int W = window_length, half = W / 2;
    std::vector<double> coef((size_t)W * (W + 1), 0.0);
    std::vector<int> pos;
    pos.push_back(half);                                          // interior taps: the fit evaluated at the window centre
    for (int i = 0; i < half; ++i) pos.push_back(i);              // left edge rows
    for (int i = 0; i < half; ++i) pos.push_back(W - half + i);   // right edge rows
    polyfit_rows(W, polyorder, pos.data(), (int)pos.size(), coef.data());
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(workspace, coef.data(), sizeof(double) * (size_t)W * (2 * half + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));     // coef dies with this frame
    SavgolArgs g{x, y, (const double*)workspace, N, ld, channels, W};
    const long long items = (long long)channels * N;
    savgol_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(g);
    CUDA_TRY(cudaGetLastError());
    return SYSID_OK;
}

// ------------------------------------------------------------------------------------------------ stage 3
size_t sysid_sdp_workspace_bytes(int32_t num_links, int32_t ndof) { return sdp_workspace_bytes(num_links, ndof); }

int sysid_sdp_solve(const sysid_sdp_desc* desc, const double* stats, int64_t stats_stride, int32_t batch,
                    double* x_out, sysid_sdp_info* info_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!desc || !stats || !x_out || !info_out || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    char msg[256] = "";
    int rc = sdp_solve_launch(*desc, stats, stats_stride, batch, x_out, info_out, workspace, workspace_bytes,
                              (cudaStream_t)stream, msg, sizeof(msg));
    if (rc != SYSID_OK) return fail(rc, "%s", msg);
    return SYSID_OK;
}

size_t sysid_sdp_plan_bytes(int32_t num_links) { return sizeof(double) * sdp_plan_total_doubles(num_links); }

int sysid_sdp_plan_create(const sysid_sdp_desc* desc, void* plan, size_t plan_bytes, void* stream) {
    if (!desc || !plan) return fail(SYSID_ERR_INVALID, "null argument");
    char msg[256] = "";
    int rc = sdp_plan_upload(*desc, (double*)plan, plan_bytes, (cudaStream_t)stream, msg, sizeof(msg));
    if (rc != SYSID_OK) return fail(rc, "%s", msg);
    return SYSID_OK;
}

size_t sysid_sdp_solve_workspace_bytes(int32_t num_links, int32_t ndof, int32_t batch) {
    return sizeof(double) * sdp_ws_doubles(num_links, ndof) * (size_t)(batch > 0 ? batch : 1);
}

size_t sysid_sdp_warm_len(int32_t num_links, int32_t ndof) { return sdp_warm_doubles(num_links, ndof); }

int sysid_sdp_solve_plan(const sysid_sdp_desc* desc, const void* plan, const double* stats, int64_t stats_stride, int32_t batch,
                         double* x_out, sysid_sdp_info* info_out, void* workspace, size_t workspace_bytes,
                         const double* warm_in, double* warm_out, void* stream) {
    if (!desc || !plan || !stats || !x_out || !info_out || !workspace) return fail(SYSID_ERR_INVALID, "null argument");
    char msg[256] = "";
    int rc = sdp_solve_planned(*desc, (const double*)plan, stats, stats_stride, batch, x_out, info_out, workspace, workspace_bytes,
                               warm_in, warm_out, (cudaStream_t)stream, msg, sizeof(msg));
    if (rc != SYSID_OK) return fail(rc, "%s", msg);
    return SYSID_OK;
}

}  // extern "C"
