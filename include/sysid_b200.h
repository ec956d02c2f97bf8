/* sysid_b200.h -- C ABI of the B200-native inertial-identification hot path.
 *
 * Drop-in boundary for xiaohu97/system_identification (reference, pure Python).  The reference has
 * no FFI layer of its own; the native code its hot path reaches lives in third-party libraries.
 * Each entry point below names the reference call site(s) it replaces:
 *
 *   sysid_model_create        pin.buildModelFromUrdf(path, JointModelFreeFlyer()) + gravity + foot frame ids
 *                             reference src/sys_identification.py:16,22,51-54  (host flattens the URDF; this uploads it)
 *   sysid_regressor_batch     pin.computeJointTorqueRegressor(model, data, q, dq, ddq)
 *                             reference src/sys_identification.py:395,406
 *   sysid_projected_batch     get_proj_regressor_torque + get_proj_friction_regressors
 *                             (_update_fk, _compute_J_c, _compute_null_space_proj = I - pinv(J_c) J_c, P@Y, P@S^T@tau,
 *                             P@S^T@diag(dq), P@S^T@diag(sign dq))   reference src/sys_identification.py:113-135,401-418
 *   sysid_gram_accumulate     the demo stacking loops + the normal equations MOSEK forms internally:
 *                             reference demo/solo_identification.py:36-55,79-84 and src/solver.py:186-190
 *   sysid_gram_accumulate_host  the same call for arrays still in host memory, exactly as read_data leaves them
 *                             (reference demo/solo_identification.py:9-33): upload and kernel overlap chunk by chunk
 *   sysid_gram_accumulate_host_ex  the same for the mixed float32 / float64 arrays read_data really returns (quirk Q8)
 *   sysid_gram_from_stack     same statistics from an already stacked (rows x c) matrix, for callers that built
 *                             Y_proj/B_v/B_c through the per-sample API and hand them to Solver(...)  src/solver.py:6-29
 *   sysid_sdp_solve           Solver.solve_fully_consistent: cvxpy problem + problem.solve(solver=cp.MOSEK)
 *                             reference src/solver.py:123-210
 *   sysid_sdp_plan_create / sysid_sdp_solve_plan   the same, split into the once-per-prior host part (Solver.__init__ and the
 *                             problem build, src/solver.py:6-29,55-121) and a launch without host work; optional warm start
 *   sysid_gram_accumulate_host_presolve   streaming statistics with the LMI fit of the first chunk solved behind the stream
 *   sysid_filtfilt / sysid_savgol  the scipy.signal.filtfilt / savgol_filter calls of read_data
 *                             reference demo/solo_identification.py:15-32 (the step immediately before the path)
 *   sysid_predict_rmse        SystemIdentification.print_tau_prediction_rmse   reference src/sys_identification.py:421-437
 *   sysid_tsqr                Solver.solve_llsq_svd: np.linalg.svd of the stacked regressor   reference src/solver.py:32-39
 *                             (the stack is first reduced to its (c+1) x (c+1) triangular factor on the device)
 *   sysid_physical_consistency  SystemIdentification.get_physical_consistency   reference src/sys_identification.py:324-389
 *   sysid_dat_scan / sysid_dat_parse / sysid_dat_parse_ex   np.loadtxt(path + name + "_robot_q.dat", delimiter='\t', dtype=np.float32): the five
 *                             loads of read_data   reference spot_identification.py:9-14, demo/solo_identification.py:9-14
 *   sysid_fd_rate / sysid_contact_from_tau   the row loops of calculate_low_motor_ddq (joint and body angular
 *                             accelerations by finite differences of the tick, contact labels from the ankle torques)
 *                             reference g1-data/low_ddq_contact_tick.py:46-81, low_ddq_tick.py:19-33, low_ddq.py:19-33
 *   sysid_round_dat           np.savetxt(fmt='%.6f') then np.loadtxt(dtype=np.float32): what csv2dat + read_data do to
 *                             a CSV column   reference g1-data/csv2dat.py:50-55
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / CUDA types in signatures (stream is a void* cudaStream_t).
 *   - Every data pointer is a DEVICE pointer unless its name ends in _host.  The caller owns all buffers; the
 *     library owns only the immutable model handle.  No hidden allocation on the hot path: scratch is passed in
 *     (sysid_gram_workspace_bytes / sysid_sdp_workspace_bytes).
 *   - Sample arrays are CHANNEL-MAJOR exactly as the reference's read_data returns them
 *     (demo/solo_identification.py:9-33): q (nq x N), dq (nv x N), ddq (nv x N), tau (d x N), contact (n_ee x N),
 *     fp64, element (ch, i) at base[ch * ld + i].
 *   - All calls are asynchronous on the given stream.  Return value: 0 = ok, negative = sysid_status.
 *     Never throws across the ABI; sysid_last_error() returns a thread-local message.
 *   - A model handle is immutable after creation and may be shared by threads/streams.
 */
#ifndef SYSID_B200_H
#define SYSID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYSID_ABI_VERSION 1

typedef enum sysid_status {
    SYSID_OK = 0,
    SYSID_ERR_INVALID = -1,      /* null pointer, bad size, malformed tree */
    SYSID_ERR_UNSUPPORTED = -2,  /* tree outside the compiled kernel envelope (see sysid_limits) */
    SYSID_ERR_CUDA = -3,         /* a CUDA runtime call failed; message has the cudaError string */
    SYSID_ERR_NOT_OPTIMAL = -4,  /* SDP did not reach tolerance ("The problem did not solve to optimality.") */
    SYSID_ERR_WORKSPACE = -5     /* workspace too small */
} sysid_status;

enum { SYSID_JT_FREEFLYER = 0, SYSID_JT_RX = 1, SYSID_JT_RY = 2, SYSID_JT_RZ = 3, SYSID_JT_RU = 4 };

/* Flat kinematic tree in pinocchio's joint numbering: index 0 = universe, 1 = free-flyer root_joint,
 * 2.. = revolute joints (depth-first, children by joint name).  All pointers are HOST pointers. */
typedef struct sysid_tree_desc {
    int32_t njoints;            /* including the universe */
    int32_t n_ee;               /* number of end-effector (foot) frames, in contact-channel order */
    const int32_t* parent;      /* [njoints] */
    const int32_t* jtype;       /* [njoints] SYSID_JT_*; entry 0 ignored */
    const double* axis;         /* [njoints*3] unit axis in the joint frame (used for SYSID_JT_RU) */
    const double* place_R;      /* [njoints*9] row-major rotation of the joint placement in the parent joint frame */
    const double* place_p;      /* [njoints*3] translation of the joint placement */
    const int32_t* ee_joint;    /* [n_ee] joint each foot frame is rigidly attached to */
    const double* ee_offset;    /* [n_ee*3] foot point in that joint's frame */
    double gravity[3];          /* linear gravity, reference uses (0, 0, -9.81) */
} sysid_tree_desc;

typedef struct sysid_model sysid_model;   /* opaque */

typedef struct sysid_dims {
    int32_t nq, nv, nbodies, ndof;        /* ndof = actuated joints d = nv - 6 */
    int32_t nparams;                      /* 10 * nbodies */
    int32_t ncols;                        /* nparams + 2 * ndof (with friction columns) */
    int32_t n_ee;
} sysid_dims;

/* Compile-time envelope of the kernels in this build. */
typedef struct sysid_limits {
    int32_t max_bodies, max_nv, max_ee, max_depth, max_cols_padded;
} sysid_limits;

int sysid_abi_version(void);
const char* sysid_last_error(void);
void sysid_get_limits(sysid_limits* out);

int sysid_model_create(const sysid_tree_desc* desc, sysid_model** out);
void sysid_model_destroy(sysid_model* model);
int sysid_model_dims(const sysid_model* model, sysid_dims* out);

/* Y_out: N x nv x nparams row-major, pinocchio column order per body [m, mcx,mcy,mcz, Ixx,Ixy,Iyy,Ixz,Iyz,Izz]. */
int sysid_regressor_batch(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                          int64_t N, int64_t ld, double* Y_out, void* stream);

/* A_out: N x nv x ncols row-major = P [Y | S^T diag(dq_j) | S^T diag(sign dq_j)]  (ncols = nparams + 2 ndof when
 * friction != 0, else nparams);  b_out: N x nv = P S^T tau;  P_out (nullable): N x nv x nv.
 * Contact rule: foot k is in stance iff contact[k] != 0 (reference quirk: state 2 counts). */
int sysid_projected_batch(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                          const double* tau, const double* contact, int64_t N, int64_t ld, int32_t friction,
                          double* A_out, double* b_out, double* P_out, void* stream);

/* Sufficient statistics of the stacked least-squares system, never materialising the stack.
 * stats layout (fp64, c = ncols as above):  G (c x c, row-major, full symmetric) | r (c) | s (1) | n (1)
 *   G = sum_i A_i^T A_i,  r = sum_i A_i^T b_i,  s = sum_i b_i^T b_i,  n = nv * N (rows of the stack, reference quirk Q4).
 * The call ADDS into stats (zero it first for a fresh accumulation), so shards/chunks can be streamed.
 * weights (nullable): per-sample non-negative weights w_i (bootstrap multiplicities); then n += nv * sum_i w_i.
 * info (nullable, device, 2 x int64): [0] += samples whose contact Jacobian lost rank (a dependent row was dropped),
 *                                     [1] += samples with a non-finite input (skipped). */
size_t sysid_stats_len(const sysid_model* model, int32_t friction);           /* doubles in stats */
size_t sysid_gram_workspace_bytes(const sysid_model* model);
int sysid_gram_accumulate(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                          const double* tau, const double* contact, int64_t N, int64_t ld, const double* weights,
                          int32_t friction, double* stats, int64_t* info, void* workspace, size_t workspace_bytes,
                          void* stream);

/* Block bootstrap (BASELINE.json configs[4]: many resamples of one log).  sysid_gram_blocks: ONE launch of the fused kernel that
 * returns one statistics vector per block of `block` consecutive samples (stats_blocks + k * stats_stride, k < ceil(N / block));
 * the statistics are additive, so the statistics of a resample that draws block k w_k times are sum_k w_k stats_k:
 * sysid_combine_stats forms out (B x slen) = weights (B x K, row-major) x stats_blocks (K x slen) for all B resamples in one launch
 * (fp64 tensor pipe).  The reference has no bootstrap; each resample is the reference's identification (demo/solo_identification.py:
 * 67-88) of the log with its blocks repeated.  slen = c*c + c + 2 must be a multiple of 8 (it is for c = 154 and c = 130). */
size_t sysid_gram_blocks_workspace_bytes(const sysid_model* model, int64_t N, int64_t block);
int sysid_gram_blocks(const sysid_model* model, const double* q, const double* dq, const double* ddq, const double* tau,
                      const double* contact, int64_t N, int64_t ld, int64_t block, int32_t friction, double* stats_blocks,
                      int64_t stats_stride, int64_t* info, void* workspace, size_t workspace_bytes, void* stream);
int sysid_combine_stats(const double* weights, int64_t B, int64_t K, const double* stats_blocks, int64_t slen, double* out, void* stream);

/* Same statistics from HOST arrays (the layout read_data returns; pinned memory for full PCIe speed): the log is
 * streamed through two device staging buffers in chunks of `chunk` samples; the copy of chunk k+1 (internal copy
 * stream, cudaMemcpy2DAsync) overlaps the kernels of chunk k on `stream`.  stats / info / workspace are DEVICE pointers;
 * weights_host is nullable.  The host arrays must stay valid until `stream` has drained.  ADDS into stats. */
size_t sysid_gram_host_workspace_bytes(const sysid_model* model, int64_t chunk);
int sysid_gram_accumulate_host(const sysid_model* model, const double* q_host, const double* dq_host, const double* ddq_host,
                               const double* tau_host, const double* contact_host, int64_t N, int64_t ld_host,
                               const double* weights_host, int32_t friction, double* stats, int64_t* info,
                               void* workspace, size_t workspace_bytes, int64_t chunk, void* stream);

/* The same call for host arrays of MIXED precision, which is what the reference's read_data actually returns
 * (demo/solo_identification.py:10-14: np.loadtxt(dtype=float32) for all five; dq / ddq / tau become float64 through
 * scipy's filtfilt, q and contact stay float32 -- SURVEY quirk Q8).  arrays_host[5] = {q, dq, ddq, tau, contact},
 * dtypes[a] in {SYSID_DTYPE_F64, SYSID_DTYPE_F32}, lds_host[a] = leading dimension of array a in ELEMENTS.  float32
 * arrays cross PCIe as float32 and are widened exactly on the device. */
#define SYSID_DTYPE_F64 0
#define SYSID_DTYPE_F32 1
int sysid_gram_accumulate_host_ex(const sysid_model* model, const void* const* arrays_host, const int32_t* dtypes,
                                  const int64_t* lds_host, int64_t N, const double* weights_host, int32_t friction,
                                  double* stats, int64_t* info, void* workspace, size_t workspace_bytes, int64_t chunk,
                                  void* stream);

/* Same statistics from a stacked matrix A (rows x c, row-major, device) and vector b (rows). */
int sysid_gram_from_stack(const double* A, const double* b, int64_t rows, int32_t c, double* stats,
                          void* workspace, size_t workspace_bytes, void* stream);
size_t sysid_gram_from_stack_workspace_bytes(int32_t c);

/* Pre-processing of read_data (reference demo/solo_identification.py:15-32), channel-major device arrays (channels x N, ld).
 * sysid_filtfilt = scipy.signal.filtfilt(b, a, x, axis=1) with its defaults (padtype 'odd', padlen 3 max(nb, na),
 * method 'pad'); b_host / a_host are HOST coefficient arrays (at most 9 each).  y may alias x.  pad_float32 != 0: the
 * values came from a float32 log (np.loadtxt(dtype=np.float32), demo/solo_identification.py:10-14): scipy then forms
 * the odd extension in float32, and so does this call (the pad samples are rounded to float32); bit-identical pads.
 * sysid_savgol = scipy.signal.savgol_filter(x, window_length, polyorder) with its defaults (deriv 0, mode 'interp');
 * y must not alias x. */
size_t sysid_filtfilt_workspace_bytes(int32_t channels, int64_t N, int32_t ncoef /* max(nb, na) */);
int sysid_filtfilt(const double* b_host, int32_t nb, const double* a_host, int32_t na, const double* x, double* y,
                   int32_t channels, int64_t N, int64_t ld, int32_t pad_float32, void* workspace, size_t workspace_bytes,
                   void* stream);
size_t sysid_savgol_workspace_bytes(int32_t window_length);
int sysid_savgol(int32_t window_length, int32_t polyorder, const double* x, double* y, int32_t channels, int64_t N,
                 int64_t ld, void* workspace, size_t workspace_bytes, void* stream);

/* Log ingest.
 * sysid_dat_scan + sysid_dat_parse = np.loadtxt(file, delimiter=<delimiter>, dtype=np.float32 | np.float64) for the text
 * np.savetxt(fmt='%.6f', delimiter='\t') writes (reference g1-data/csv2dat.py:50-55): `text` is the file's bytes in DEVICE
 * memory (4-byte aligned, trailing blank lines trimmed by the caller).  scan counts rows and columns (dims_host[0..1], HOST;
 * it synchronises the stream because the caller sizes `out` from it) and leaves the field offsets in the workspace; parse
 * converts every field into out (rows x cols, leading dimension ld, fp64; round_float32 != 0: rounded through float32
 * and widened, which is what dtype=np.float32 yields).  Conversion is exact (equal to strtod's) on the fields it accepts:
 * up to 2^53 as a digit string with a decimal exponent within +-22, which covers every "%.6f" field below 9.007e9, plus
 * nan / inf.  Anything else is counted, never guessed: info_host = {bad fields, index of the first, misplaced row ends,
 * fields found}; with info_host non-null the call synchronises and returns SYSID_ERR_INVALID when any count is non-zero
 * (np.loadtxt raises ValueError in those cases). */
size_t sysid_dat_workspace_bytes(int64_t nbytes);
int sysid_dat_scan(const void* text, int64_t nbytes, int32_t delimiter, void* workspace, size_t workspace_bytes,
                   int64_t* dims_host, void* stream);
int sysid_dat_parse(const void* text, int64_t nbytes, int32_t delimiter, void* workspace, size_t workspace_bytes,
                    int64_t rows, int64_t cols, double* out, int64_t ld, int32_t round_float32, int64_t* info_host, void* stream);
/* The same with options, for the logger CSV the G1 scripts start from (pd.read_csv, reference g1-data/csv2dat.py:15,
 * low_ddq_contact_tick.py:21; pass the text AFTER its header line, delimiter ','): SYSID_DAT_TRANSPOSE writes element
 * (row, col) to out[col * ld + row], so a row-per-sample CSV lands channel-major; SYSID_DAT_EMPTY_IS_NAN reads an empty
 * field as NaN (what pandas does) instead of counting it as an error (what np.loadtxt does). */
enum { SYSID_DAT_ROUND_FLOAT32 = 1, SYSID_DAT_TRANSPOSE = 2, SYSID_DAT_EMPTY_IS_NAN = 4 };
int sysid_dat_parse_ex(const void* text, int64_t nbytes, int32_t delimiter, void* workspace, size_t workspace_bytes,
                       int64_t rows, int64_t cols, double* out, int64_t ld, int32_t flags, int64_t* info_host, void* stream);

/* The row loop of calculate_low_motor_ddq (reference g1-data/low_ddq_contact_tick.py:46-70), all channels at once:
 * y[ch][0] = NaN; for i >= 1 with dt = tick[i] - tick[i-1], dx = x[ch][i] - x[ch][i-1]:  dt > 0 -> (dx * scale) / dt;
 * else dx == 0 -> 0; else NaN.  scale = 1000 for millisecond ticks (low_ddq_tick.py:28), 1 for low_ddq.py:27. */
int sysid_fd_rate(const double* tick, const double* x, double* y, int32_t channels, int64_t N, int64_t ld_x, int64_t ld_y,
                  double scale, void* stream);
/* out[i] = tau[i] >= hi ? 1 : (tau[i] > lo ? 2 : 0): the contact labels of low_ddq_contact_tick.py:72-81 (hi 10, lo -5). */
int sysid_contact_from_tau(const double* tau, double* out, int64_t N, double hi, double lo, void* stream);
/* y = the value x has after np.savetxt(fmt='%.6f') and np.loadtxt (to_float32 != 0: with dtype=np.float32), computed
 * without the text: exact decimal rounding (ties to even on the exact binary value, as printf does) and the correctly
 * rounded read-back.  y may alias x. */
int sysid_round_dat(const double* x, double* y, int32_t channels, int64_t N, int64_t ld_x, int64_t ld_y, int32_t to_float32,
                    void* stream);

/* Communication-avoiding QR of the stacked system [A | b] (A: rows x c row-major, b: rows or NULL; device): R_out receives
 * the (c+1) x (c+1) upper-triangular factor [R z; 0 rho] (row-major, zeros below the diagonal), R^T R = A^T A, z = Q^T b,
 * rho = the least-squares residual norm (up to sign).  R has the singular values and right singular vectors of A, so
 * Solver.solve_llsq_svd (reference src/solver.py:32-39) is V diag(1/sigma_i, sigma_i > 1e-15 sigma_max) U_R^T z with
 * U_R S V^T the SVD of the c x c R -- without squaring the condition number as the Gram route would.  Deterministic. */
size_t sysid_tsqr_workspace_bytes(int32_t c);
int sysid_tsqr(const double* A, const double* b, int64_t rows, int32_t c, double* R_out, void* workspace, size_t workspace_bytes,
               void* stream);

/* get_physical_consistency (reference src/sys_identification.py:324-389) for `batch` parameter vectors (phi + i phi_stride,
 * 10 num_links each, reference order m, h, Ixx Ixy Ixz Iyy Iyz Izz): out[i][0..4][link] = min eig I_bar, min eig of the 6x6
 * spatial inertia, min eig J, min eig C, tr(J Q).  Matrices the reference builds as np.float32 are rounded entry by entry
 * to float32; eigenvalues by fp64 Jacobi (the reference's float32 LAPACK answer agrees to float32 precision). */
int sysid_physical_consistency(const double* phi, int64_t phi_stride, int32_t batch, int32_t num_links, const double* semi_axes,
                               const double* centers, double* out, void* stream);

/* LMI-constrained fit (reference src/solver.py:123-210).  All pointers in the desc are HOST pointers. */
enum { SYSID_REG_CONSTANT_PULLBACK = 0, SYSID_REG_EUCLIDEAN = 1 };
typedef struct sysid_sdp_desc {
    int32_t num_links;            /* L */
    int32_t ndof;                 /* friction coefficients per kind (0 = no friction identification) */
    const double* phi_prior;      /* [10 L] reference order [m, hx,hy,hz, Ixx,Ixy,Ixz,Iyy,Iyz,Izz]; float32 values widened */
    const double* semi_axes;      /* [3 L] bounding-ellipsoid semi axes */
    const double* centers;        /* [3 L] bounding-ellipsoid centres */
    double total_mass;
    double lambda_reg;            /* reference default 1e-1 */
    int32_t reg_type;             /* SYSID_REG_* */
    double epsilon;               /* LMI margin, reference 1e-6 */
    double tol;                   /* reference default 1e-10 (MOSEK rel-gap); here KKT residuals are driven below 10 * tol (relative) */
    int32_t max_iters;            /* cap on Newton steps (reference: MOSEK iterations, default 1000; 0 = library default) */
} sysid_sdp_desc;

typedef struct sysid_sdp_info {
    int32_t status;               /* 0 optimal, 1 optimal-inaccurate (residuals within 1e3 x tolerance at the iteration cap;
                                     accepted like cvxpy's OPTIMAL_INACCURATE), SYSID_ERR_NOT_OPTIMAL otherwise */
    int32_t iterations;           /* Newton steps */
    int32_t refactorizations;     /* augmented-Lagrangian (multiplier) updates */
    int32_t reserved;
    double primal_residual, dual_residual, rho, objective;   /* rho = final penalty sigma */
    double min_eig_J, min_eig_C;  /* smallest eigenvalue over links of J+eps I and C+eps I at the solution */
    double mass_residual;
} sysid_sdp_info;

size_t sysid_sdp_workspace_bytes(int32_t num_links, int32_t ndof);
/* stats: device, layout above with c = 10 L + 2 ndof.  x_out: device, c doubles [phi | b_v | b_c].
 * info_out: device, one sysid_sdp_info per problem.  batch >= 1 solves `batch` independent problems whose stats are
 * stats + k * stats_stride (same priors/ellipsoids), one thread block each. */
int sysid_sdp_solve(const sysid_sdp_desc* desc, const double* stats, int64_t stats_stride, int32_t batch,
                    double* x_out, sysid_sdp_info* info_out, void* workspace, size_t workspace_bytes, void* stream);

/* The same solve split in two, for callers that solve more than once with the same prior / ellipsoids / lambda (a streaming
 * identify(), bootstrap batches): sysid_sdp_plan_create does the host work of Solver.__init__ + the problem build (reference
 * src/solver.py:6-29,55-121: pull-back metrics, svec maps, float32 Q) once and uploads it (synchronises `stream`);
 * sysid_sdp_solve_plan then launches with NO host work and NO synchronisation.  `desc` supplies only the scalars there
 * (num_links, ndof, total_mass, epsilon, tol, max_iters).  workspace: sysid_sdp_solve_workspace_bytes(L, nd, batch).
 * Warm start: warm_out (nullable, device, batch * sysid_sdp_warm_len doubles) receives [x | multipliers | sigma | valid] in
 * unscaled units; warm_in (nullable) starts the iteration from such a record of a NEARBY problem -- the optimum is unique, so
 * the start changes the number of Newton steps, not the answer.  A record that is not finite or not valid is ignored. */
size_t sysid_sdp_plan_bytes(int32_t num_links);
int sysid_sdp_plan_create(const sysid_sdp_desc* desc, void* plan, size_t plan_bytes, void* stream);
size_t sysid_sdp_solve_workspace_bytes(int32_t num_links, int32_t ndof, int32_t batch);
size_t sysid_sdp_warm_len(int32_t num_links, int32_t ndof);
int sysid_sdp_solve_plan(const sysid_sdp_desc* desc, const void* plan, const double* stats, int64_t stats_stride, int32_t batch,
                         double* x_out, sysid_sdp_info* info_out, void* workspace, size_t workspace_bytes,
                         const double* warm_in, double* warm_out, void* stream);

/* sysid_gram_accumulate_host_ex with an LMI PRE-SOLVE hidden behind the stream (pre nullable = plain streaming).  The statistics
 * are additive, so the fit of the first `samples` samples of the log is a point and a set of multipliers within the statistical
 * noise of the final fit: it is solved as ONE thread block on an internal stream while the remaining chunks are uploaded and
 * contracted (they leave it one SM), and its record (warm_out) warm-starts the final sysid_sdp_solve_plan, which then needs a
 * handful of Newton steps instead of ~55 (fewer still with the second stage, refine_at).  On return, work submitted to `stream` is ordered after the pre-solve.  Skipped (record
 * left invalid) when the log is shorter than 2 * samples.  All pointers in the struct are DEVICE pointers except desc. */
typedef struct sysid_presolve {
    const sysid_sdp_desc* desc;       /* host; scalars only */
    const void* plan;                 /* sysid_sdp_plan_create */
    void* sdp_workspace; size_t sdp_workspace_bytes;   /* sysid_sdp_solve_workspace_bytes(L, nd, 1) */
    double* stats_snapshot;           /* stats_len doubles */
    double* x_scratch;                /* c doubles */
    sysid_sdp_info* info_scratch;     /* one record */
    double* warm_out;                 /* sysid_sdp_warm_len doubles */
    int64_t samples;                  /* length of the first chunk (0 = `chunk`) */
    int64_t refine_at;                /* > 0: once that many samples are in, a SECOND pre-solve of the statistics so far is queued behind
                                         the first one, warm-started from it; its record replaces the first one's in warm_out.  Pays on
                                         logs long enough to hide both (~1 M samples); 0 = one stage */
    double* stats_snapshot2;          /* stats_len doubles (needed when refine_at > 0) */
    double first_tol;                 /* > 0: tolerance of the FIRST pre-solve (the second stage and the final solve keep desc->tol): with a
                                         second stage behind it, the first only has to get the multipliers roughly right */
} sysid_presolve;
int sysid_gram_accumulate_host_presolve(const sysid_model* model, const void* const* arrays_host, const int32_t* dtypes,
                                        const int64_t* lds_host, int64_t N, const double* weights_host, int32_t friction,
                                        double* stats, int64_t* info, void* workspace, size_t workspace_bytes, int64_t chunk,
                                        const sysid_presolve* pre, void* stream);

/* tau-prediction error of phi (nparams, multiplies the pinocchio-ordered regressor as is -- reference quirk Q1):
 * out (device): [0] = mean_i ||e_i||^2 (the reference's "total", no root), [1..ndof] = per-joint RMSE,
 * with e_i = (P Y phi)[6:] - (P S^T tau)[6:]. */
int sysid_predict_rmse(const sysid_model* model, const double* q, const double* dq, const double* ddq,
                       const double* tau, const double* contact, int64_t N, int64_t ld, const double* phi,
                       double* out, void* workspace, size_t workspace_bytes, void* stream);
size_t sysid_predict_rmse_workspace_bytes(const sysid_model* model);

#ifdef __cplusplus
}
#endif
#endif /* SYSID_B200_H */
