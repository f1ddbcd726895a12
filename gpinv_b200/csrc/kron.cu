// Kronecker-factored sampling stage (T6 + T8-T13 for kronecker.RBF2D).
//
// The reference multiplies chol(K1) (x) chol(K2) out into a dense [n1 n2, n1 n2] matrix
// (GPinv/kronecker.py:10-22,49-66) and then treats it like any other Cholesky factor in
// GPinv/stvgp.py:174-183.  Here the product stays factored:
//
//     (L1 (x) L2) vec(U) = vec(L1 U L2^T),   U = mat(u) [n1, n2],  row index i*n2 + j
//                                            (the order pinned by testing/test_kronecker.py:37-41)
//
// and every sample matrix keeps its [R][S][n1][n2] layout, so no transposed copy is ever made:
//
//     T      = U_flat L2^T                    one GEMM  [(R S n1) x n2] x [n2 x n2]      (L2 lower)
//     F_r[b] = sv_r L1 T_r[b] + mean_r        batched GEMM over the S samples of output r, L1 shared
//
// 2 n (n1 + n2) flops per sample instead of 2 n^2.  Reverse (hand-written):
//
//     G_r[b]  = sv_r L1^T Fbar_r[b]                      batched, L1 shared
//     Ubar    = G_flat L2 (+ klbar U)                    one GEMM
//     L2bar   = tril(G_flat^T U_flat)                    one GEMM, K = R S n1
//     L1bar_r = sum_b Fbar_r[b] T_r[b]^T                 batched products + one column-sum pass
//     svbar_r = <L1bar_r, L1>,   L1bar = sum_r sv_r tril(L1bar_r)
#include "gemm.cuh"

namespace gpinv {

// from sample.cu
int sample_u_fwd_f64(const double* Q, int q_diag, const double* q_mu, const double* E, int n, int S, int R,
                     double* U, double* kl_out, cudaStream_t st);
int sample_q_bwd_f64(const double* Q, int q_diag, const double* E, const double* Ubar, const double* klbar,
                     int n, int S, int R, double* qmu_bar, int qmu_from_colsum, double* Qbar, cudaStream_t st);
int colsum_f64(const double* X, int n, int S, int R, double* out, cudaStream_t st);
int kl_finish_f64(double* kl_out, int S, cudaStream_t st);

// F[b] = alpha * op(L1) U[b] op(L2)^T  for b < batch, written as two DMMA GEMMs through T_ws [batch,n1,n2].
//   trans = 0: F[b] = L1 U[b] L2^T   ( (L1 (x) L2) u )        trans = 1: F[b] = L1^T U[b] L2   ( ... ^T u )
// alpha_dev (device scalar, may be NULL) multiplies alpha; Cin [n1,n2] (may be NULL) is added to every
// F[b] (the mean of one latent output); expF (may be NULL) receives exp(F).
int kron_apply_f64(const double* L1, int ldl1, int n1, const double* L2, int ldl2, int n2, const double* U,
                   int batch, double alpha, const double* alpha_dev, const double* Cin, double* T_ws, double* F,
                   double* expF, int trans, cudaStream_t st) {
  if (n1 <= 0 || n2 <= 0 || batch <= 0) return GPINV_ERR_ARG;
  const long long n = (long long)n1 * n2;
  int rc;
  {  // T = U_flat op(L2)^T : [(batch n1) x n2]
    GemmParams p = {};
    p.A = U; p.lda = n2;
    p.B = L2; p.ldb = ldl2;
    p.M = batch * n1; p.N = n2; p.K = n2;
    p.alpha = 1.0;
    p.out = T_ws; p.ldo = n2;
    p.splitk = 1; p.allow_split = 1;
    // trans = 0: (U L2^T)(i,c) = sum_k U(i,k) L2[c][k]   -> B is LK,  k <= column  (kmode 1)
    // trans = 1: (U L2)(i,c)   = sum_k U(i,k) L2[k][c]   -> B is LMN, k >= column  (kmode 2)
    p.kmode = trans ? 2 : 1;
    rc = gemm_launch(LK, trans ? LMN : LK, EF_BETA, p, st);
    if (rc) return rc;
  }
  {  // F[b] = alpha op(L1) T[b] (+ Cin) : batched, op(L1) shared
    GemmParams p = {};
    p.A = L1; p.lda = ldl1; p.strideA = 0;
    p.B = T_ws; p.ldb = n2; p.strideB = n;
    p.M = n1; p.N = n2; p.K = n1;
    p.alpha = alpha; p.alpha_dev = alpha_dev;
    p.beta = Cin ? 1.0 : 0.0; p.Cin = Cin; p.ldcin = n2; p.strideCin = 0;
    p.out = F; p.ldo = n2; p.strideO = n;
    p.out2 = expF; p.ldo2 = n2;
    p.batch = batch;
    p.splitk = 1;
    // trans = 0: A(i,k) = L1[i][k], zero for k > i (kmode 3);  trans = 1: A(i,k) = L1[k][i], zero for k < i (kmode 4)
    p.kmode = trans ? 4 : 3;
    rc = gemm_launch(trans ? LMN : LK, LMN, expF ? (EF_BETA | EF_EXP2) : EF_BETA, p, st);
    if (rc) return rc;
  }
  return GPINV_OK;
}

// out[r][i][k] lower: L1bar = sum_r sv_r tril(L1u_r); svdot_r = <tril(L1u_r), L1>
__global__ void __launch_bounds__(256) kron_l1_finish_kernel(const double* __restrict__ L1u, const double* __restrict__ L1,
                                                             int ldl1, const double* __restrict__ sqrt_v, int n1, int R,
                                                             double* __restrict__ L1bar, int ldb,
                                                             double* __restrict__ svdot) {
  __shared__ double red[32];
  const int total = n1 * n1;
  double dots[8];
  for (int r = 0; r < 8; ++r) dots[r] = 0.0;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx / n1, k = idx % n1;
    double acc = 0.0;
    if (k <= i) {
      const double l = L1[(size_t)i * ldl1 + k];
      for (int r = 0; r < R; ++r) {
        const double v = L1u[(size_t)r * total + idx];
        acc += sqrt_v[r] * v;
        if (r < 8) dots[r] += v * l;
      }
    }
    L1bar[(size_t)i * ldb + k] = acc;
  }
  for (int r = 0; r < R && r < 8; ++r) {
    const double s = block_sum(dots[r], red);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(svdot + r, s);
  }
}

// svdot for outputs r >= 8 (rare): one block per r
__global__ void __launch_bounds__(256) kron_svdot_kernel(const double* __restrict__ L1u, const double* __restrict__ L1,
                                                         int ldl1, int n1, int r0, double* __restrict__ svdot) {
  __shared__ double red[32];
  const int r = r0 + blockIdx.x;
  const int total = n1 * n1;
  double acc = 0.0;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int i = idx / n1, k = idx % n1;
    if (k <= i) acc += L1u[(size_t)r * total + idx] * L1[(size_t)i * ldl1 + k];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) svdot[r] = acc;
}

size_t kron_fwd_ws_bytes(int n1, int n2, int S, int R) { return (size_t)R * S * n1 * n2 * sizeof(double); }
size_t kron_bwd_ws_bytes(int n1, int n2, int S, int R) {
  // G [R,S,n] + per-sample products [S,n1,n1] + L1u [R,n1,n1]
  return ((size_t)R * S * n1 * n2 + (size_t)S * n1 * n1 + (size_t)R * n1 * n1) * sizeof(double);
}

// Forward of the sampling stage with a Kronecker core.  T [R,S,n] is kept for the reverse pass.
int kron_sample_fwd_f64(const double* L1, int n1, const double* L2, int n2, const double* sqrt_v, const double* Q,
                        int q_diag, const double* q_mu, const double* mean, const double* E, int S, int R, double* U,
                        double* T, double* F, double* expF, double* kl_out, cudaStream_t st) {
  const int n = n1 * n2;
  if (n1 <= 0 || n2 <= 0 || S <= 0 || R <= 0 || !F) return GPINV_ERR_ARG;
  int rc = sample_u_fwd_f64(Q, q_diag, q_mu, E, n, S, R, U, kl_out, st);
  if (rc) return rc;
  const size_t sn = (size_t)S * n;
  for (int r = 0; r < R; ++r) {
    rc = kron_apply_f64(L1, n1, n1, L2, n2, n2, U + r * sn, S, 1.0, sqrt_v + r, mean + (size_t)r * n, T + r * sn,
                        F + r * sn, expF ? expF + r * sn : nullptr, 0, st);
    if (rc) return rc;
  }
  return kl_finish_f64(kl_out, S, st);
}

// Reverse.  Outputs: Ubar [R,S,n] (workspace, kept for inspection), qmu_bar [R,n], Qbar ([R,n,n] lower or [R,n]),
// L1bar [n1,n1] lower, L2bar [n2,n2] lower, svdot [R] = dObj/d sqrt_v, mean_bar [R,n].
int kron_sample_bwd_f64(const double* L1, int n1, const double* L2, int n2, const double* sqrt_v, const double* Q,
                        int q_diag, const double* E, const double* U, const double* T, const double* Fbar,
                        const double* klbar, int S, int R, double* Ubar, double* qmu_bar, double* Qbar, double* L1bar,
                        double* L2bar, double* svdot, double* mean_bar, double* ws, cudaStream_t st) {
  const int n = n1 * n2;
  if (n1 <= 0 || n2 <= 0 || S <= 0 || R <= 0) return GPINV_ERR_ARG;
  const size_t sn = (size_t)S * n;
  double* G = ws;                               // [R,S,n]
  double* PP = G + (size_t)R * sn;              // [S,n1,n1]
  double* L1u = PP + (size_t)S * n1 * n1;       // [R,n1,n1]
  GPINV_TRY(cudaMemsetAsync(svdot, 0, (size_t)R * sizeof(double), st));
  int rc = colsum_f64(Fbar, n, S, R, mean_bar, st);
  if (rc) return rc;
  for (int r = 0; r < R; ++r) {
    // G_r[b] = sv_r L1^T Fbar_r[b]
    {
      GemmParams p = {};
      p.A = L1; p.lda = n1; p.strideA = 0;
      p.B = Fbar + r * sn; p.ldb = n2; p.strideB = n;
      p.M = n1; p.N = n2; p.K = n1;
      p.alpha = 1.0; p.alpha_dev = sqrt_v + r;
      p.out = G + r * sn; p.ldo = n2; p.strideO = n;
      p.batch = S; p.splitk = 1; p.kmode = 4;
      rc = gemm_launch(LMN, LMN, EF_BETA, p, st);
      if (rc) return rc;
    }
    // per-sample products Fbar_r[b] T_r[b]^T -> PP [S,n1,n1], then summed over the samples
    {
      GemmParams p = {};
      p.A = Fbar + r * sn; p.lda = n2; p.strideA = n;
      p.B = T + r * sn; p.ldb = n2; p.strideB = n;
      p.M = n1; p.N = n1; p.K = n2;
      p.alpha = 1.0;
      p.out = PP; p.ldo = n1; p.strideO = (long long)n1 * n1;
      p.batch = S; p.splitk = 1; p.kmode = 0;
      rc = gemm_launch(LK, LK, EF_BETA, p, st);
      if (rc) return rc;
      rc = colsum_f64(PP, n1 * n1, S, 1, L1u + (size_t)r * n1 * n1, st);
      if (rc) return rc;
    }
  }
  {  // Ubar = G_flat L2 + klbar U
    GemmParams p = {};
    p.A = G; p.lda = n2;
    p.B = L2; p.ldb = n2;
    p.M = R * S * n1; p.N = n2; p.K = n2;
    p.alpha = 1.0;
    p.beta = klbar ? 1.0 : 0.0; p.beta_dev = klbar;
    p.Cin = U; p.ldcin = n2;
    p.out = Ubar; p.ldo = n2;
    p.kmode = 2; p.splitk = 1; p.allow_split = 1;
    rc = gemm_launch(LK, LMN, EF_BETA, p, st);
    if (rc) return rc;
  }
  {  // L2bar = tril(G_flat^T U_flat),  K = R S n1
    GemmParams p = {};
    p.A = G; p.lda = n2;
    p.B = U; p.ldb = n2;
    p.M = n2; p.N = n2; p.K = R * S * n1;
    p.alpha = 1.0;
    p.out = L2bar; p.ldo = n2;
    p.kmode = 0; p.splitk = 1; p.allow_split = 1;
    rc = gemm_launch(LMN, LMN, EF_BETA | EF_LOWER, p, st);
    if (rc) return rc;
  }
  kron_l1_finish_kernel<<<min(cdiv(n1 * n1, 256), 64), 256, 0, st>>>(L1u, L1, n1, sqrt_v, n1, R, L1bar, n1, svdot);
  GPINV_CHECK_LAUNCH();
  if (R > 8) {
    kron_svdot_kernel<<<R - 8, 256, 0, st>>>(L1u, L1, n1, n1, 8, svdot);
    GPINV_CHECK_LAUNCH();
  }
  return sample_q_bwd_f64(Q, q_diag, E, Ubar, klbar, n, S, R, qmu_bar, 1, Qbar, st);
}

}  // namespace gpinv
