/* gpinv_b200 — C ABI of the B200 (sm_100a) StVGP / GPMC hot path.
 *
 * The reference (fujii-team/GPinv) has no FFI: its hot path is a TensorFlow graph built in
 * Python.  Each entry point below replaces the group of TF ops emitted by the cited reference
 * lines, and is what a ctypes binding on the reference side would load.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp64 (or int32 for `info`) unless stated otherwise;
 *    the library never allocates or frees device memory — workspaces are passed in;
 *  - matrices are row-major with an explicit leading dimension (elements);
 *  - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no synchronisation);
 *  - return value: 0 ok, <0 argument / CUDA error (message from gpinv_last_error());
 *    numerical failure of the Cholesky is reported LAPACK-style in the device int `info`
 *    (first failing pivot, 1-based), like tf.cholesky's InvalidArgumentError;
 *  - sample matrices are stored sample-major [R][S][n] (the reference's eps layout
 *    [R,N,n,1], GPinv/stvgp.py:164).
 */
#ifndef GPINV_B200_H
#define GPINV_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPINV_MODE_RBF 0   /* GPinv/kernels.py:105-107 */
#define GPINV_MODE_CSYM 1  /* GPinv/kernels.py:115-120 */
#define GPINV_MODE_CASYM 2 /* GPinv/kernels.py:140-145 */

int gpinv_version(void);
const char* gpinv_last_error(void);

/* Optional scratch for the GEMM kernel's split-K wave-quantisation fix: device memory owned by
 * the caller, ZERO-INITIALISED, registered once per device (first 64 KB are tile counters).
 * Without it the GEMMs run data-parallel only.  Calls that use it must be stream-ordered. */
int gpinv_set_workspace(void* ptr, size_t bytes);

/* Strided 2-D copy on `stream` (cudaMemcpy2DAsync): `height` rows of `width_bytes`, row pitches in
 * bytes; kind 1 = host->device, 2 = device->host, 3 = device->device.  Used by the host API of
 * Model._objective (the boundary the reference's GPflow Model._objective closure forms,
 * pattern at GPinv/model.py:64-75) to move only the lower trapezoids of the q_sqrt gradient into
 * pinned host memory. */
int gpinv_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch,
                         size_t width_bytes, size_t height, int kind, void* stream);

/* Host half of the same boundary, upload direction: x_host is the pageable row-major [nq,nq]
 * q_sqrt block of the optimiser's state vector (GPinv/stvgp.py:71-73; only its lower triangle is
 * read, :157).  Row blocks of rb rows are gathered (columns [0, r1) of rows [r0, r1)) into the
 * pinned buffer at element offsets off[k] by nthreads host threads, and each block's contiguous
 * host->device copy into dstage + off[k] is issued on `stream` as soon as the block is staged.
 * take[k] == 0 skips block k.  Returns once every copy is issued (not completed).  dstage == NULL
 * gathers only. */
int gpinv_upload_tril_blocks_f64(const double* x_host, int nq, int rb, double* pinned, double* dstage,
                                 const long long* off, const int* take, int nblocks, int nthreads,
                                 void* stream);

/* --- T1-T3: kernel core matrix.  Replaces Kern._Kcore + jitter (GPinv/kernels.py:57-58,
 * 105-107,115-120,140-145; GPflow Stationary.square_dist).  X [n,D], X2 [n2,D] or NULL
 * (then K is n x n and `jitter` is added to its diagonal); ls device [D] (ard) or [1]. */
size_t gpinv_kbuild_ws_bytes(int n, int n2, int D);
int gpinv_kbuild_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
                     const double* ls, int ard, int mode, double jitter, double* K, int ldk,
                     void* ws, void* stream);
/* reverse: ls_bar[D or 1] = sum_ij Kbar_ij dcore_ij/dls (replaces tf.gradients through _Kcore) */
int gpinv_kbuild_bwd_f64(const double* X, const double* X2, int n, int n2, int D, int ldx,
                         int ldx2, const double* ls, int ard, int mode, const double* Kbar,
                         int ldkb, double* ls_bar, void* ws, void* stream);

/* --- T4: Cholesky, in place, lower, strict upper zeroed.  Replaces tf.cholesky at
 * GPinv/kernels.py:59, GPinv/kronecker.py:54,57.  `info` is a device int[2]: info[0] receives the
 * LAPACK-style status, info[1] is scratch (a ticket counter used by the panel kernel). */
int gpinv_potrf_f64(double* A, int n, int lda, int* info, void* stream);
/* reverse (replaces TF's CholeskyGrad): Abar holds Lbar (lower) on entry, the full symmetric
 * dObj/dA on exit. */
size_t gpinv_potrf_bwd_ws_bytes(int n);
int gpinv_potrf_bwd_f64(const double* L, int ldl, double* Abar, int ldab, int n, void* ws,
                        void* stream);

/* --- T1-T4 fused: L = chol(core(X,X) + jitter I) and its reverse.  Replaces the whole body of
 * Kern.Cholesky up to the variance scaling (GPinv/kernels.py:56-59: _Kcore + eye*jitter +
 * tf.cholesky) and, in reverse, TF's CholeskyGrad followed by the gradient of _Kcore w.r.t. the
 * lengthscales.  Only the lower triangle of the core is ever evaluated; the reverse keeps the
 * blocked sweep's lower-triangular result (strict part doubled) and reduces it straight into
 * ls_bar [D if ard else 1], so no symmetric n x n gradient is materialised.  L [n,n] receives a
 * clean lower-triangular factor; info as in gpinv_potrf_f64. */
size_t gpinv_chol_core_ws_bytes(int n, int D);
int gpinv_chol_core_f64(const double* X, int n, int D, int ldx, const double* ls, int ard, int mode,
                        double jitter, double* L, int ldl, int* info, void* ws, void* stream);
/* The same with W = L^-1 (Winv [n,n]; scratch [n,n]) built alongside: row block by row block inside the
 * factorisation on a low-priority stream of the library (n > 512), else right behind it on the side stream.  The
 * inverse is what the reverse pass (gpinv_chol_core_bwd2_f64) multiplies with; `stream` does NOT wait for it:
 * order a stream after it with gpinv_trtri_side_join. */
int gpinv_chol_core_inv_f64(const double* X, int n, int D, int ldx, const double* ls, int ard, int mode,
                            double jitter, double* L, int ldl, int* info, double* Winv, double* scratch, void* ws,
                            void* stream);
size_t gpinv_chol_core_bwd_ws_bytes(int n, int D);
int gpinv_chol_core_bwd_f64(const double* X, int n, int D, int ldx, const double* ls, int ard,
                            int mode, const double* L, int ldl, const double* Lbar, int ldlb,
                            double* ls_bar, void* ws, void* stream);
/* Same reverse with the explicit inverse W = L^-1 [n,n] supplied (built ahead of time by
 * gpinv_trtri_side_f64 on the library's side stream, overlapping the sampling / likelihood stages:
 * the inverse depends on L only).  winv_ready = 0: the call first makes `stream` wait for that side-stream
 * work; 1: the caller has already done so with gpinv_trtri_side_join (needed when the forward and the reverse
 * pass are captured into different CUDA graphs: an event of one capture cannot be waited on in another). */
int gpinv_chol_core_bwd2_f64(const double* X, int n, int D, int ldx, const double* ls, int ard,
                             int mode, const double* L, int ldl, const double* Lbar, int ldlb,
                             const double* Winv, int winv_ready, double* ls_bar, void* ws, void* stream);
/* Winv <- L^-1 on the side stream once the work already enqueued on `stream` is done; `scratch` [n,n].
 * gpinv_trtri_side_join makes `stream` wait for the most recent such inversion. */
int gpinv_trtri_side_f64(const double* L, int ldl, double* Winv, double* scratch, int n, void* stream);
int gpinv_trtri_side_join(void* stream);
/* Sample sharding over G GPUs, n = 64 * 2^k with (n/2) % 4G == 0: like gpinv_trtri_side_f64, but the last merge
 * level of the inversion, W21 = -W22 L21 W11 with blocks of n/2 (three quarters of its flops), is computed for
 * this rank's column slices of W21 only: 2G slices of width n / 4G, rank r owns slices r and 2G-1-r.  The other
 * slices of Winv[n/2:, :n/2] stay zero; the caller all-gathers them (gpinv_b200/dist.py). */
int gpinv_trtri_side_shard_f64(const double* L, int ldl, double* Winv, double* scratch, int n, int rank, int G,
                               void* stream);
/* host-side wait for the side stream (before an inverse buffer is recycled outside a reverse pass) */
int gpinv_trtri_side_sync(void);
/* --- T22: X <- X L^{-T} (row-wise forward substitution; X is m x n).  Replaces
 * tf.batch_matrix_triangular_solve at GPinv/conditionals.py:44 (applied to Kmn^T). */
int gpinv_trsm_rlt_f64(const double* L, int ldl, double* X, int ldx, int m, int n, void* stream);

/* --- generic DMMA GEMM: C = alpha op(A) op(B) + beta C.  transa/transb as in row-major BLAS
 * ('N' = 0, 'T' = 1).  lower != 0: only the lower triangle of a square C is computed. */
int gpinv_gemm_f64(int transa, int transb, int m, int n, int k, double alpha, const double* A,
                   int lda, const double* B, int ldb, double beta, double* C, int ldc, int lower,
                   void* stream);

/* Bring-up / test entry: the same GEMM with the launcher's knobs exposed.  la / lb: 0 = element (r,k)
 * at p[r*ld + k], 1 = at p[k*ld + r]; kmode: triangular k-range skipping (0 none, 1: k < n0+BN,
 * 2: k >= n0, 3: k < m0+BM, 4: k >= m0, 5: both 2 and 3); batch problems `stride*` elements apart. */
int gpinv_gemm_ex_f64(int la, int lb, int m, int n, int k, double alpha, const double* A, int lda,
                      long long strideA, const double* B, int ldb, long long strideB, double beta, double* C,
                      int ldc, long long strideC, int batch, int lower, int kmode, int allow_split, void* stream);

/* --- T7: tril(q_sqrt) with the reference's [n,n,R] -> [R,n,n] transpose
 * (GPinv/stvgp.py:156-157) and its reverse. */
int gpinv_tril_unpack_f64(const double* q_sqrt_nnR, double* Q_Rnn, int n, int R, void* stream);
int gpinv_tril_pack_f64(const double* Qbar_Rnn, double* out_nnR, int n, int R, void* stream);

/* --- T8-T13: u = q_mu + tril(q_sqrt) eps, KL_MC, f = sqrt(v) Lc u + mean
 * (GPinv/stvgp.py:144-185 with Kern.Cholesky's sqrt(variance) scaling, GPinv/kernels.py:60-63,
 * folded in as a scalar).  Q: [R,n,n] lower (q_diag=0) or [R,n] (q_diag=1).
 * Outputs U [R,S,n]; F and/or expF [R,S,n] (either may be NULL, not both);
 * kl_out device [4] = {KL, sum eps^2, sum u^2, sum log Q_ii^2}. */
int gpinv_sample_fwd_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q,
                         int q_diag, const double* q_mu, const double* mean, const double* E,
                         int n, int S, int R, double* U, double* F, double* expF, double* kl_out,
                         void* stream);
/* The first half of that stage, u = q_mu + tril(q_sqrt) eps and the KL partial sums (GPinv/stvgp.py:162-172), does
 * not depend on the Cholesky factor of GPinv/kernels.py:59.  gpinv_side_mark_inputs marks the point of `stream`
 * where its inputs exist (call it BEFORE gpinv_chol_core_*); gpinv_sample_u_early_f64 (call it AFTER) runs that
 * half, CTA-capped, on a low-priority stream of the library, behind the marked point and the middle of the most
 * recent factorisation -- the product hides inside the factorisation's latency chain; gpinv_sample_fwd2_f64
 * with u_ready = 1 waits for it (gpinv_side_early_join) and continues with f = sqrt(v) Lc u + mean.  n > 512. */
int gpinv_side_mark_inputs(void* stream);
int gpinv_side_early_join(void* stream);
int gpinv_sample_u_early_f64(const double* Q, int q_diag, const double* q_mu, const double* E, int n, int S, int R,
                             double* U, double* kl_out, int max_ctas);
int gpinv_sample_fwd2_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q, int q_diag,
                          const double* q_mu, const double* mean, const double* E, int n, int S, int R, double* U,
                          double* F, double* expF, double* kl_out, int u_ready, void* stream);
/* reverse.  Fbar [R,S,n] = dObj/dF, klbar device scalar dObj/dKL (NULL = 0).
 * Outputs: Ubar [R,S,n]; qmu_bar [R,n]; Qbar [R,n,n] lower (or [R,n]); Lc_bar [n,n] lower
 * (summed over r); svdot [R] = sv_r <tril(Fbar_r^T U_r), Lc>; mean_bar [R,n]. */
int gpinv_sample_bwd_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q,
                         int q_diag, const double* E, const double* U, const double* Fbar,
                         const double* klbar, int n, int S, int R, double* Ubar, double* qmu_bar,
                         double* Qbar, double* Lc_bar, int ldlb, double* svdot, double* mean_bar,
                         void* stream);

/* --- T6 kept factored: Kronecker core chol(K1) (x) chol(K2) of kronecker.RBF2D
 * (GPinv/kronecker.py:10-22 kronecker_product, :49-66 RBF2D.Cholesky, which the reference multiplies
 * out into a dense [n1 n2, n1 n2] matrix).  Row order i*n2 + j (testing/test_kronecker.py:37-41).
 * gpinv_kron_apply_f64: F[b] = L1 U[b] L2^T (trans = 0, i.e. (L1 (x) L2) u) or L1^T U[b] L2 (trans = 1)
 * for `batch` vectors stored as rows of U / F [batch, n1*n2]; T_ws [batch, n1*n2] scratch. */
int gpinv_kron_apply_f64(const double* L1, int n1, const double* L2, int n2, const double* U, int batch,
                         double* T_ws, double* F, int trans, void* stream);
/* Sampling stage of GPinv/stvgp.py:144-185 with that factored core: same contract as
 * gpinv_sample_fwd_f64 / gpinv_sample_bwd_f64 with Lc replaced by (L1 [n1,n1], L2 [n2,n2]), n = n1*n2.
 * T [R,S,n] (= U L2^T per sample) is an extra output kept for the reverse, which returns L1bar
 * [n1,n1] and L2bar [n2,n2] (lower triangles) and svdot [R] = dObj/d sqrt(variance_r). */
int gpinv_kron_sample_fwd_f64(const double* L1, int n1, const double* L2, int n2, const double* sqrt_v,
                              const double* Q, int q_diag, const double* q_mu, const double* mean,
                              const double* E, int S, int R, double* U, double* T, double* F, double* expF,
                              double* kl_out, void* stream);
size_t gpinv_kron_bwd_ws_bytes(int n1, int n2, int S, int R);
int gpinv_kron_sample_bwd_f64(const double* L1, int n1, const double* L2, int n2, const double* sqrt_v,
                              const double* Q, int q_diag, const double* E, const double* U, const double* T,
                              const double* Fbar, const double* klbar, int S, int R, double* Ubar,
                              double* qmu_bar, double* Qbar, double* L1bar, double* L2bar, double* svdot,
                              double* mean_bar, void* ws, void* stream);

/* --- T14: Gaussian likelihood (GPinv/likelihoods.py:70-76): sum (F - Y)^2 over [R,S,n] with
 * Y [n,R]; reverse Fbar = coef (Y - F), coef a device scalar. */
int gpinv_gauss_sumsq_f64(const double* F, const double* Y, int n, int S, int R, double* out,
                          void* stream);
int gpinv_gauss_bwd_f64(const double* F, const double* Y, int n, int S, int R, const double* coef,
                        double* Fbar, void* stream);

/* --- T15: Abel transform + Gaussian residual (notebooks/Abel_inversion.ipynb:274-283).
 * Rs [S,M] = 1 y^T - expF [S,n] A[M,n]^T, *sumsq = sum Rs^2;
 * reverse Fbar [S,n] = coef expF o (Rs A), coef device scalar (= dObj/dlik / variance). */
int gpinv_abel_fwd_f64(const double* expF, const double* A, int lda, const double* Y, int n, int M,
                       int S, double* Rs, double* sumsq, void* stream);
int gpinv_abel_bwd_f64(const double* Rs, const double* A, int lda, const double* expF,
                       const double* coef, int n, int M, int S, double* Fbar, void* stream);

/* --- T16: spectroscopic Abel transform (notebooks/Spectroscopic_Abel_inversion.ipynb:316-356).
 * F [3,S,n] = (log a, log sigma, v); At, cosTt: the notebook's [m,n] geometry matrices TRANSPOSED to
 * [n,m] (chord index contiguous, so that every access of both kernels is coalesced); lam [k];
 * Y [k,m] (may be NULL with sumsq); n <= 9000.  dlam: spacing of the wavelength grid when it is uniform
 * (lam_k = lam_0 + k dlam, the notebook's np.linspace) -- the Gaussian line shape is then advanced along k by
 * a two-term recurrence restarted every 16 wavelengths instead of one exp per term -- or 0 for arbitrary lam.
 * fwd: P [S,k,m] (may be NULL), *sumsq = sum (P - Y)^2 (may be NULL);
 * bwd: Fbar [3,S,n] = coef * sum_{k,m} (P - Y) dP/dF, coef a device scalar. */
int gpinv_spec_fwd_f64(const double* F, const double* At, const double* cosTt, const double* lam,
                       const double* Y, int n, int m, int k, int S, double dlam, double* P, double* sumsq,
                       void* stream);
int gpinv_spec_bwd_f64(const double* F, const double* At, const double* cosTt, const double* lam,
                       const double* Y, const double* P, const double* coef, int n, int m, int k,
                       int S, double dlam, double* Fbar, void* stream);

/* --- T19 and its reverse: the optimiser's packed free state x (GPflow sorted_params order; the x of
 * GPinv/model.py:64-75) <-> constrained parameter blocks.  Block b covers x[off[b], off[b+1]);
 * transform[b] = 0 (identity) or 1 (GPflow transforms.positive: y = log(1 + exp(x)) + 1e-6, used at
 * GPinv/stvgp.py:69, GPinv/likelihoods.py:68 and by the Stationary kernels).  off / transform / dst / gval
 * are HOST arrays (copied into the launch).  gpinv_unpack_f64 writes the constrained values to dst[b];
 * gpinv_pack_f64 writes gx[off[b] + k] = scale * gval[b][k] * dy/dx (gval[b] == NULL: zeros), i.e. the
 * packed gradient that GPflow's _objective closure returns (scale = -1 for "-gradient"). */
int gpinv_unpack_f64(const double* x, int nblocks, const long long* off, const int* transform,
                     double* const* dst, void* stream);
int gpinv_pack_f64(const double* x, int nblocks, const long long* off, const int* transform,
                   const double* const* gval, double scale, double* gx, void* stream);

/* --- StVGP._analytical_KL (GPinv/stvgp.py:187-195 -> GPflow kullback_leiblers.gauss_kl_white[_diag]):
 * KL = 1/2 sum m^2 - 1/2 nR - 1/2 sum log Q_ii^2 + 1/2 sum tril(Q)^2 with q_mu [n,R] and q_sqrt in the
 * reference's layout [n,n,R] (q_diag = 0) or [n,R] (q_diag = 1).  kl_out device [4] = {KL, partial sums}.
 * Reverse: mu_bar = gbar m, q_bar = gbar (tril(Q) - diag(1/Q_ii)) in the same layouts. */
int gpinv_kl_white_f64(const double* q_mu, const double* q_sqrt, int n, int R, int q_diag, double* kl_out,
                       void* stream);
int gpinv_kl_white_bwd_f64(const double* q_mu, const double* q_sqrt, int n, int R, int q_diag,
                           const double* gbar, double* mu_bar, double* q_bar, void* stream);

/* --- Gaussian log-density from the residual sum of squares (GPflow densities.gaussian behind
 * GPinv/likelihoods.py:70-76 and the notebooks' Abel / spectroscopic likelihoods), all device scalars:
 * out = -count/2 (log 2 pi + log var) - ss / (2 var);  reverse: ss_bar = -g / (2 var),
 * var_bar = g (-count / (2 var) + ss / (2 var^2)). */
int gpinv_gauss_logp_f64(const double* ss, const double* var, double count, double* out, void* stream);
int gpinv_gauss_logp_bwd_f64(const double* ss, const double* var, double count, const double* g, double* ss_bar,
                             double* var_bar, void* stream);

/* --- optimiser step of the training loop (GPinv/model.py:206-240 drives a tf.train optimiser;
 * testing/test_stvgp.py:31-35: tf.train.AdamOptimizer).  TensorFlow's Adam update in place on the
 * device-resident state: t = step + 1; lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);
 * m = beta1 m + (1 - beta1) g; v = beta2 v + (1 - beta2) g^2; x -= lr_t m / (sqrt(v) + epsilon); step += 1.
 * `step` is a device int64 so that the launch can be replayed inside a CUDA graph. */
int gpinv_adam_step_f64(double* x, double* m, double* v, const double* g, long long count, double lr,
                        double beta1, double beta2, double epsilon, long long* step, void* stream);

/* --- sample sharding over G GPUs (the reference has no distributed path).  The reverse of
 * the Cholesky at GPinv/kernels.py:59 is linear in Lbar and only its contraction with dK/d(lengthscale) is
 * needed: with W = L^-1, G_K = W^T Phi(L^T Lbar) W = sum over column blocks J of (W^T Phi[:,J]) W[J,:], and
 * Phi[:,J] needs Lbar[:,J] only.  The ranks therefore reduce-scatter Lbar by column blocks of width
 * w = n / 2G (rank r owns blocks r and 2G-1-r), reverse their own blocks, and sum the scalar results.
 *   gpinv_lbar_pack_f64: out = G chunks of (n + w) w doubles, chunk r = [tril(Lbar)[r w:, block r] ;
 *     tril(Lbar)[(2G-1-r) w:, block 2G-1-r]] (rows x w, row-major) -- the reduce-scatter input.
 *   gpinv_chol_core_bwd_shard_f64: Dsh = this rank's reduced chunk, c0s (HOST) = first column of each of its
 *     nblk blocks in chunk order, Winv = L^-1 from gpinv_trtri_side_f64 (the caller has already ordered
 *     `stream` after it with gpinv_trtri_side_join); ls_bar <- this rank's share of dObj/d(lengthscales).
 *     ws: gpinv_chol_core_bwd_ws_bytes(n, D).
 *   gpinv_tri_compact_f64 / gpinv_tri_expand_f64: lower triangle of an n x n row-major matrix <-> packed
 *     rows [i (i + 1) / 2 + c]; the q_sqrt gradient crosses NVLink in this form (half the bytes). */
int gpinv_lbar_pack_f64(const double* Lbar, int ldlb, int n, int w, int G, double* out, void* stream);
int gpinv_chol_core_bwd_shard_f64(const double* X, int n, int D, int ldx, const double* ls, int ard, int mode,
                                  const double* L, int ldl, const double* Winv, const double* Dsh, int w,
                                  const int* c0s, int nblk, double* ls_bar, void* ws, void* stream);
int gpinv_tri_compact_f64(const double* src, int ld, int n, double* dst, void* stream);
int gpinv_tri_expand_f64(const double* src, double* dst, int ld, int n, void* stream);

/* --- reductions used by T8/T11/T17 */
int gpinv_sumsq_f64(const double* x, size_t count, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPINV_B200_H */
