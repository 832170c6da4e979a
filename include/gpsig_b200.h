/*
 * gpsig_b200.h -- C ABI of the B200-native signature-kernel covariance path.
 *
 * This is the drop-in boundary: every entry point replaces one TensorFlow-graph function of tgcsaba/GPSig
 * (file:line relative to the reference checkout) and is what an FFI binding on the reference side would call
 * (see INTEGRATION.md for the ctypes stub).  Conventions, all entry points:
 *   - plain C types only; device pointers are fp32, row-major, owned by the caller; nothing is allocated for the
 *     caller and no global mutable state is kept (re-entrant per stream);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous;
 *   - return value: 0 = ok, negative = GPSIG_E_* argument error, positive = cudaError_t of a failed launch;
 *   - level stacks are (num_levels+1, ...) leading-axis, level 0 == 1 (signature_algs.py:20-23,35).
 * Arithmetic is fp32 (the reference is fp64 on TF; tolerance stated in tests/ and DESIGN.md).
 */
#ifndef GPSIG_B200_H
#define GPSIG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSIG_B200_VERSION 200 /* 0.2.0 */

enum {
    GPSIG_OK = 0,
    GPSIG_E_BADARG = -1,      /* null pointer / non-positive size / inconsistent shapes */
    GPSIG_E_UNSUPPORTED = -2, /* valid in the reference but outside what this build implements (see message) */
    GPSIG_E_WORKSPACE = -3,   /* workspace too small: call the matching *_workspace_bytes() */
    GPSIG_E_ALIGN = -4,       /* pointer / stride alignment required by the kernel not met */
    GPSIG_E_DRIVER = -5       /* cuTensorMapEncodeTiled unavailable or failed */
};

/* static (state-space) kernels, kernels.py:786-993 */
enum {
    GPSIG_KERN_LINEAR = 0,   /* SignatureLinear   kernels.py:799-806 */
    GPSIG_KERN_RBF = 1,      /* SignatureRBF      kernels.py:862-864 (+ _square_dist :765-776) */
    GPSIG_KERN_COSINE = 2,   /* SignatureCosine   kernels.py:820-828 */
    GPSIG_KERN_POLY = 3,     /* SignaturePoly     kernels.py:844-848   params = {gamma, degree} */
    GPSIG_KERN_MIX = 4,      /* SignatureMix      kernels.py:881-892   params = {mixing} */
    GPSIG_KERN_MATERN12 = 5, /* kernels.py:955-958 */
    GPSIG_KERN_MATERN32 = 6, /* kernels.py:974-977 */
    GPSIG_KERN_MATERN52 = 7, /* kernels.py:991-993 */
    GPSIG_KERN_SPECTRAL = 8  /* SignatureSpectral kernels.py:894-942, families 'gauss' / 'exp' ('mixed' is broken upstream);
                                params = {family (0 gauss, 1 exp), Q, d, alpha[Q], omega[Q*d], gamma[Q*d]}, Q <= 8, d <= 16 */
};

int gpsig_version(void);
/* static string for a return code of any function below (never NULL) */
const char* gpsig_error_string(int code);
/* free-form detail of the last GPSIG_E_* raised on the calling thread (thread-local, "" if none) */
const char* gpsig_last_error_detail(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Measurement hooks (no reference counterpart: the reference has no profiling subsystem).
 *   gpsig_launch_count(): number of kernels this library has launched since it was loaded (monotone, all threads).
 *   gpsig_profile_enable(1): from now on every launch of the classes below is bracketed by two CUDA events on the
 *     stream it is launched on; gpsig_profile_read() synchronises those events and returns, for one class, the summed
 *     device time, the number of launches and the work units they processed (sequence pairs / (z, n) pairs);
 *     gpsig_profile_reset() drops the records.  Disabled by default (zero overhead beyond one relaxed load).
 * ------------------------------------------------------------------------------------------------------------- */
enum {
    GPSIG_PROF_PREP = 0,       /* scaling / time increments / norms of the points */
    GPSIG_PROF_PRODUCER = 1,   /* increment-Gram chunk producer (a3 + signature_algs.py:26) */
    GPSIG_PROF_RECURSION = 2,  /* TMA-staged first-order recursion (a4) -- the dominant kernel */
    GPSIG_PROF_RECURSION_OTHER = 3, /* generic first-order and higher-order recursions (a4 fallback, a5) */
    GPSIG_PROF_EPILOGUE = 4,   /* normalise / weight / sum, mirror (a7) */
    GPSIG_PROF_TENS = 5,       /* inducing-tensor kernels (a9-a11) and low-rank kernels (a15) */
    GPSIG_PROF_FUSED = 6,      /* fused increment-Gram + recursion kernel (a3 + a4 in one launch, no HBM intermediate) */
    GPSIG_PROF_VJP = 7,        /* reverse-mode kernels (gpsig_*_vjp) */
    GPSIG_PROF_NUM_CLASSES = 8
};
long long gpsig_launch_count(void);
/* Tuning / experiment knobs.  Their defaults come from the environment ONCE, at first use (GPSIG_WARPFUSED,
 * GPSIG_WARPFUSED_WARPS, GPSIG_STREAM_NCW / _R / _S, GPSIG_TENS_TC); afterwards only this call changes them
 * (names: "warpfused", "warpfused_warps", "stream_ncw", "stream_r", "stream_s", "tens_tc").  Not thread safe against
 * concurrent launches. */
int gpsig_set_knob(const char* name, int value);
int gpsig_profile_enable(int on);
int gpsig_profile_reset(void);
int gpsig_profile_read(int cls, double* total_ms, long long* launches, double* units);

/* ---------------------------------------------------------------------------------------------------------------
 * a2  kernels.py:342-364 (_apply_scaling_and_lags_to_sequences, no-lags branch) and :366-398 (tensors):
 *     out[r, c] = X[r, c] * inv_lengthscales[c % num_features]        (inv_lengthscales may be NULL = copy)
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_scale_features(const float* X, long rows, int d, const float* inv_lengthscales, int num_features,
                         float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a2  gpsig/lags.py:7-63 (add_lags_to_sequences; called at kernels.py:352-353): out[n, l, 0, :] = X[n, l, :] and
 *     out[n, l, 1 + p, :] = X[n] linearly interpolated at time max(l / (L-1) - lags[p], 0).
 *     X (n, L, d), lags (num_lags) device pointers; out (n, L, (num_lags + 1) * d).
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_add_lags(const float* X, long n, int L, int d, const float* lags, int num_lags, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a3  static-kernel Gram, kernels.py:225-230 (`self._base_kern(X, X2)`) on already-scaled points.
 *     X (rows1, d), X2 (rows2, d) row-major; out (rows1, rows2) with leading dimension ld (elements).
 *     X2 == NULL means X2 = X.  `params` are the kernel's extra scalars (see enum), may be NULL.
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_gram(int kind, const float* X, long rows1, const float* X2, long rows2, int d, const float* params,
               float* out, long ld, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a4/a5  signature_algs.py:8-35 (signature_kern_first_order) and :37-74 (signature_kern_higher_order).
 *     M is the Gram tensor [n1, L1, n2, L2] addressed as M[i*stride_i + s*stride_s + j*stride_j + t] (elements,
 *     unit stride along t).  A 3-D (n, L, L) tensor (signature_algs.py:21-23) is the case n1 = 1, n2 = n,
 *     stride_s = L2, stride_j = L1*L2.  out_levels is (num_levels+1, n1, n2) dense.
 *     order == 1 -> first order; 1 < order <= num_levels -> higher order.  difference as signature_algs.py:25-26.
 *     upper_only != 0 (requires n1 == n2): only entries j >= i are computed and written (caller mirrors).
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_sigkern_levels(const float* M, int n1, int L1, int n2, int L2, long stride_i, long stride_s, long stride_j,
                         int num_levels, int order, int difference, int upper_only, float* out_levels, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a6  kernels.py:208-237 (_K_seq) and :188-205 (_K_seq_diag): Gram + recursion without ever holding the full
 *     [n1,L1,n2,L2] tensor: row blocks of the (increment) Gram are produced into `workspace` and consumed by the
 *     recursion kernel.  X (n1, L1, d), X2 (n2, L2, d) are RAW sequences; inv_lengthscales (d) may be NULL.
 *     X2 == NULL -> symmetric K(X, X): only tiles j >= i are computed; mirror != 0 then fills j < i from the
 *     transpose (needs the full problem).  Rows [row_begin, row_end) of X are computed (row sharding over GPUs:
 *     each rank owns row blocks) and written to out_levels (num_levels+1, out_rows_total, n2) at row
 *     out_row0 + (i - row_begin).  The whole problem is row_begin = 0, row_end = n1, out_row0 = 0,
 *     out_rows_total = n1.   diag variant: out_levels (num_levels+1, n).
 * ------------------------------------------------------------------------------------------------------------- */
size_t gpsig_seq_kern_workspace_bytes(int n1, int L1, int n2, int L2, int d, size_t budget_bytes);
int gpsig_seq_kern_levels(int kind, const float* params, const float* X, int n1, int L1, const float* X2, int n2,
                          int L2, int d, const float* inv_lengthscales, int num_levels, int order, int difference,
                          int row_begin, int row_end, float* out_levels, long out_row0, long out_rows_total, int mirror,
                          void* workspace, size_t workspace_bytes, void* stream);
/* The same for a LIST of row blocks (a GPU's shard of the symmetric problem; no reference counterpart):
 *     row_blocks = host array {begin_0, end_0, begin_1, end_1, ...} of global row ranges; out_levels is the compact stack
 *     (num_levels+1, out_rows_total = sum of block sizes, n2) holding block after block.  One launch covers all blocks
 *     where the fused kernel applies.  Symmetric (X2 == NULL): only entries j >= i of every row are written. */
int gpsig_seq_kern_levels_blocks(int kind, const float* params, const float* X, int n1, int L1, const float* X2, int n2,
                                 int L2, int d, const float* inv_lengthscales, int num_levels, int order, int difference,
                                 const int* row_blocks, int num_blocks, float* out_levels, long out_rows_total,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* workspace of gpsig_seq_kern_diag_levels for n sequences (the prepared points of ALL n sequences live in it) */
size_t gpsig_seq_kern_diag_workspace_bytes(int n, int L, int d, size_t budget_bytes);
/* levels[m][i][j] = levels[m][j][i] for i > j, levels (nl, n, n) */
int gpsig_mirror_upper(float* levels, int nl, int n, void* stream);
/* Multi-GPU assembly of a symmetric matrix from gathered row shards (no reference counterpart):
 *     K[i][j] = rows[row_src[i]][j] for j >= i, rows[row_src[j]][i] for j < i;  rows (*, n) holds every global row i at
 *     position row_src[i] (n ints, device) with only its entries j >= i valid; K (n, n). */
int gpsig_assemble_symmetric(const float* rows, const int* row_src, int n, float* K, void* stream);
int gpsig_seq_kern_diag_levels(int kind, const float* params, const float* X, int n, int L, int d,
                               const float* inv_lengthscales, int num_levels, int order, int difference,
                               float* out_levels, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a7  kernels.py:430-433 / :455-469 (normalisation), :471 (x sigma*variances), :473-476 (sum over levels).
 *     levels (nl, n1, n2) in place.  diag1 (nl, n1), diag2 (nl, n2): level diagonals K_m(x_i,x_i), may be NULL
 *     (= no normalisation on that side).  symmetric != 0 reproduces kernels.py:431-433: jitter is added to the
 *     diagonal ENTRIES of `levels` first and diag1 == diag2 == that diagonal (both pointers ignored).
 *     diag_cols (n1 ints, may be NULL): row i of `levels` is row diag_cols[i] of a symmetric problem whose other
 *     rows live elsewhere (GPU row shard); jitter is added at column diag_cols[i] and both diagonals are used, which
 *     makes the shard bit-identical to the same rows of the symmetric single-GPU result.
 *     weights (nl) = sigma * variances.  out_sum (n1, n2) may be NULL (levels only); if levels_out is NULL the
 *     weighted levels are not written back.
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_normalize_weight_sum(const float* levels, int nl, long n1, long n2, const float* diag1, const float* diag2,
                               const int* diag_cols, float jitter, int symmetric, const float* weights,
                               float* levels_out, float* out_sum, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a9  signature_algs.py:76-99 (tensor_kern): M (T, nz, nz2), T = num_levels(num_levels+1)/2 -> (num_levels+1,nz,nz2)
 *     increments != 0: M is the raw Gram (T, nz, 2, nz2, 2) and the 2x2 increment of kernels.py:275-277 is fused.
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_tensor_kern_levels(const float* M, int num_levels, long nz, long nz2, int increments, float* out_levels,
                             void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a10/a11  signature_algs.py:101-127 / :129-160 (tensor vs sequence).  M (T, nz, n, L) dense, or with
 *     increments != 0 the raw Gram (T, nz, 2, n, L) of kernels.py:329 whose z-increment (:330) is fused.
 *     out_levels (num_levels+1, nz, n).
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_tens_vs_seq_levels(const float* M, int num_levels, long nz, long n, int L, int order, int difference,
                             int increments, float* out_levels, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a10+a3 fused (kernels.py:313-340, _K_tens_vs_seq): Z (T, nz, d) or (T, nz, 2, d), X (n, L, d) raw, never
 *     materialising the (T, nz, n, L) Gram.  out_levels (num_levels+1, nz, n).
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_tens_seq_kern_levels(int kind, const float* params, const float* Z, long nz, int increments, const float* X,
                               long n, int L, int d, const float* inv_lengthscales, int num_levels, int order,
                               int difference, float* out_levels, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * f1  Reverse mode of the two first-order recursions.  The reference has no counterpart in its own sources: it trains
 *     through signature_algs.py:28-33 and :116-125 by TensorFlow autodiff (gpsig/training.py:140-203 drives it); a binding
 *     registers these as the gradients of the ops that replace signature_kern_first_order /
 *     signature_kern_tens_vs_seq_first_order.  Both act on the INCREMENTS, i.e. on what
 *     signature_algs.py:26 / :114 (and kernels.py:330) produce -- the forward value is gpsig_sigkern_levels /
 *     gpsig_tens_vs_seq_levels with difference = 0, increments = 0 on the same tensor.
 *       gpsig_sigkern_levels_vjp:     Delta addressed like M of gpsig_sigkern_levels (n1, L1, n2, L2 = increment counts),
 *                                     G (num_levels+1, n1, n2) = dL/d levels, Delta_bar dense (n1, L1, n2, L2) = dL/dDelta.
 *       gpsig_tens_vs_seq_levels_vjp: H (T, nz, n, Lh) dense, G (num_levels+1, nz, n), H_bar (T, nz, n, Lh).
 *     First order only (order == 1).  Nothing per entry is stored between the sweeps: the forward state is run backwards.
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_sigkern_levels_vjp(const float* Delta, int n1, int L1, int n2, int L2, long stride_i, long stride_s,
                             long stride_j, int num_levels, const float* G, float* Delta_bar, void* stream);
int gpsig_tens_vs_seq_levels_vjp(const float* H, int num_levels, long nz, long n, int Lh, const float* G, float* H_bar,
                                 void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a15  low-rank mode.  The random projection of low_rank_calculations.py:152-193 (very sparse Gaussian JL, matrix R of
 *     shape (k1*k2, r); row q pairs A[..., q % k1] with B[..., q / k1]) is passed in compressed-sparse-column form:
 *     colptr (r+1), and per non-zero ia = row % k1, ib = row / k1, val = R[row, col]; scale = sqrt(s / r).
 *     The coordinate-subsampling variant (:104-127) is one non-zero per column (val = Rademacher sign), scale = 1.
 *     Randomness is the caller's (TensorFlow's streams are not reproducible): parity is defined on given draws.
 *       gpsig_lr_hadamard_csc: C[x, c] = scale * sum_e A[x, ia[e]] * B[x, ib[e]] * val[e];  A (rows,k1) B (rows,k2) C (rows,r)
 *       gpsig_lr_seq_level   : one iteration of signature_algs.py:182-188 --
 *                              P_out[n,t,:] = proj(U[n,t,:], sum_{t'<t} P_in[n,t',:]),  phi[n,:] = sum_t P_out[n,t,:]
 *                              U (n, Lr, k1), P_in (n, Lr, k2), P_out (n, Lr, r), phi (n, r)
 * ------------------------------------------------------------------------------------------------------------- */
int gpsig_lr_hadamard_csc(const float* A, long rows, int k1, const float* B, int k2, const int* colptr, const int* ia,
                          const int* ib, const float* val, int r, float scale, float* out, void* stream);
int gpsig_lr_seq_level(const float* U, const float* P_in, long n, int Lr, int k1, int k2, const int* colptr, const int* ia,
                       const int* ib, const float* val, int r, float scale, float* P_out, float* phi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPSIG_B200_H */
