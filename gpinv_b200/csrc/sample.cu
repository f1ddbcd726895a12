// StVGP Monte-Carlo sampling stage (T7-T13), likelihood stages (T14, T15) and their
// hand-written reverse passes.  All sample matrices are stored sample-major: [R][S][n]
// (row = one MC sample, n contiguous) which is the reference's eps layout [R,S,n,1]
// (GPinv/stvgp.py:164) and makes every product a K-contiguous DMMA GEMM:
//
//   U_r = E_r Q_r^T + 1 m_r^T           (GPinv/stvgp.py:162-168, S.R mat-vecs in the reference)
//   F_r = sv_r U_r Lc^T + 1 mu_r^T      (GPinv/stvgp.py:174-183)
//   KL  = -S/2 sum log Q_ii^2 - 1/2 sum E^2 + 1/2 sum U^2     (GPinv/stvgp.py:160-161,170-172)
//   Abel:  R = 1 y^T - exp(F) A^T ,  lik = -SM/2 log(2 pi s2) - sum R^2 / (2 s2)
//          (notebooks/Abel_inversion.ipynb:274-283 + GPflow densities.gaussian)
#include "gemm.cuh"

namespace gpinv {

// ------------------------------ small HBM-bound kernels -------------------------------

// dst[r][i][j] = (j <= i) ? src[(i*n + j)*R + r] : 0      (GPinv/stvgp.py:157 band_part)
// One CTA per (row i, chunk of 256 columns): the chunk's 256*R source doubles are contiguous and are
// read coalesced (through shared memory when R > 1, so that the R destination rows are written
// coalesced too); columns above the diagonal are not read at all.  R == 1: 16-byte accesses.
constexpr int TRIL_JC = 256;
__global__ void __launch_bounds__(256) tril_unpack_kernel(const double* __restrict__ src, double* __restrict__ dst, int n,
                                                          int R) {
  extern __shared__ __align__(16) double tsm[];
  const int i = blockIdx.y;
  const int j0 = blockIdx.x * TRIL_JC;
  const int jn = min(TRIL_JC, n - j0);          // columns of this chunk
  const int jv = max(0, min(jn, i + 1 - j0));   // ... of which at or below the diagonal
  if (R == 1) {
    const double* s = src + (size_t)i * n + j0;
    double* d = dst + (size_t)i * n + j0;
    if ((((size_t)i * n + j0) & 1) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
      const int c = 2 * threadIdx.x;
      if (c + 1 < jn) {
        double2 v = make_double2(0.0, 0.0);
        if (c + 1 < jv) v = *reinterpret_cast<const double2*>(s + c);
        else if (c < jv) v.x = s[c];
        *reinterpret_cast<double2*>(d + c) = v;
      } else if (c < jn) {
        d[c] = (c < jv) ? s[c] : 0.0;
      }
    } else {
      for (int c = threadIdx.x; c < jn; c += blockDim.x) d[c] = (c < jv) ? s[c] : 0.0;
    }
    return;
  }
  const double* s = src + ((size_t)i * n + j0) * R;
  for (int k = threadIdx.x; k < jv * R; k += blockDim.x) tsm[k] = s[k];
  __syncthreads();
  for (int r = 0; r < R; ++r) {
    double* d = dst + ((size_t)r * n + i) * n + j0;
    for (int c = threadIdx.x; c < jn; c += blockDim.x) d[c] = (c < jv) ? tsm[c * R + r] : 0.0;
  }
}

// dst[(i*n + j)*R + r] = (j <= i) ? src[r][i][j] : 0     (reverse of tril_unpack)
__global__ void __launch_bounds__(256) tril_pack_kernel(const double* __restrict__ src, double* __restrict__ dst, int n,
                                                        int R) {
  extern __shared__ __align__(16) double tsm[];
  const int i = blockIdx.y;
  const int j0 = blockIdx.x * TRIL_JC;
  const int jn = min(TRIL_JC, n - j0);
  const int jv = max(0, min(jn, i + 1 - j0));
  if (R == 1) {
    const double* s = src + (size_t)i * n + j0;
    double* d = dst + (size_t)i * n + j0;
    if ((((size_t)i * n + j0) & 1) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
      const int c = 2 * threadIdx.x;
      if (c + 1 < jn) {
        double2 v = make_double2(0.0, 0.0);
        if (c + 1 < jv) v = *reinterpret_cast<const double2*>(s + c);
        else if (c < jv) v.x = s[c];
        *reinterpret_cast<double2*>(d + c) = v;
      } else if (c < jn) {
        d[c] = (c < jv) ? s[c] : 0.0;
      }
    } else {
      for (int c = threadIdx.x; c < jn; c += blockDim.x) d[c] = (c < jv) ? s[c] : 0.0;
    }
    return;
  }
  for (int r = 0; r < R; ++r) {
    const double* s = src + ((size_t)r * n + i) * n + j0;
    for (int c = threadIdx.x; c < jv; c += blockDim.x) tsm[c * R + r] = s[c];
  }
  __syncthreads();
  double* d = dst + ((size_t)i * n + j0) * R;
  for (int k = threadIdx.x; k < jn * R; k += blockDim.x) d[k] = (k < jv * R) ? tsm[k] : 0.0;
}

// *out += sum x^2 ; 128-bit loads when aligned
__global__ void __launch_bounds__(256) sumsq_kernel(const double* __restrict__ x, size_t count,
                                                    double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    const size_t n2 = count / 2;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    for (size_t k = i; k < n2; k += stride) {
      const double2 v = x2[k];
      acc += v.x * v.x + v.y * v.y;
    }
    if (i == 0 && (count & 1)) acc += x[count - 1] * x[count - 1];
  } else {
    for (size_t k = i; k < count; k += stride) acc += x[k] * x[k];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// out[r][i] += sum_s X[r][s][i]
__global__ void __launch_bounds__(256) colsum_kernel(const double* __restrict__ X, int n, int S,
                                                     int schunk, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.z;
  if (i >= n) return;
  const int s0 = blockIdx.y * schunk, s1 = min(S, s0 + schunk);
  const double* p = X + ((size_t)r * S) * n + i;
  double acc = 0.0;
  for (int s = s0; s < s1; ++s) acc += p[(size_t)s * n];
  atomicAdd(out + (size_t)r * n + i, acc);
}

// logdet[0] += sum_i log(q_i^2) over the diagonals (full: Q[r][i][i], diag: q[r][i])
__global__ void __launch_bounds__(256) diag_logsq_kernel(const double* __restrict__ Q, int n, int R,
                                                         int q_diag, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  const int total = n * R;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const int r = k / n, i = k % n;
    const double q = q_diag ? Q[(size_t)r * n + i] : Q[((size_t)r * n + i) * n + i];
    acc += log(q * q);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// parts = {sumE2, sumU2, logdet};  kl = -S/2 logdet - sumE2/2 + sumU2/2
__global__ void kl_finish_kernel(const double* __restrict__ parts, double S, double* __restrict__ kl) {
  kl[0] = -0.5 * S * parts[2] - 0.5 * parts[0] + 0.5 * parts[1];
}

// Qbar[r][i][i] += klbar * (-S) / Q[r][i][i]   (full)   /   qbar[r][i] likewise (diag)
__global__ void diag_term_kernel(double* __restrict__ Qbar, const double* __restrict__ Q, int n,
                                 int R, int q_diag, const double* __restrict__ klbar, double S) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n * R) return;
  const int r = k / n, i = k % n;
  const size_t off = q_diag ? ((size_t)r * n + i) : (((size_t)r * n + i) * n + i);
  Qbar[off] += klbar[0] * (-S) / Q[off];
}

// q_diag forward: U = m + q o E, *sumU2 += sum U^2
__global__ void __launch_bounds__(256) udiag_fwd_kernel(const double* __restrict__ q,
                                                        const double* __restrict__ qmu,
                                                        const double* __restrict__ E, int n, int S,
                                                        double* __restrict__ U,
                                                        double* __restrict__ sumU2) {
  __shared__ double red[32];
  const int r = blockIdx.z;
  const size_t total = (size_t)S * n;
  const size_t base = (size_t)r * total;
  double acc = 0.0;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
       k += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(k % n);
    const double u = qmu[(size_t)r * n + i] + q[(size_t)r * n + i] * E[base + k];
    U[base + k] = u;
    acc += u * u;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(sumU2, acc);
}

// q_diag backward: qbar[r][i] += sum_s Ubar E ; mbar[r][i] += sum_s Ubar
__global__ void __launch_bounds__(256) udiag_bwd_kernel(const double* __restrict__ Ubar,
                                                        const double* __restrict__ E, int n, int S,
                                                        int schunk, double* __restrict__ qbar,
                                                        double* __restrict__ mbar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.z;
  if (i >= n) return;
  const int s0 = blockIdx.y * schunk, s1 = min(S, s0 + schunk);
  const size_t base = ((size_t)r * S) * n + i;
  double aq = 0.0, am = 0.0;
  for (int s = s0; s < s1; ++s) {
    const double ub = Ubar[base + (size_t)s * n];
    aq += ub * E[base + (size_t)s * n];
    am += ub;
  }
  atomicAdd(qbar + (size_t)r * n + i, aq);
  if (mbar) atomicAdd(mbar + (size_t)r * n + i, am);
}

// Gaussian likelihood on latent samples: *out += sum_{r,s,i} (F[r][s][i] - Y[i*R + r])^2
__global__ void __launch_bounds__(256) gauss_sumsq_kernel(const double* __restrict__ F,
                                                          const double* __restrict__ Y, int n,
                                                          int S, int R, double* __restrict__ out) {
  __shared__ double red[32];
  const size_t total = (size_t)R * S * n;
  double acc = 0.0;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
       k += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(k % n);
    const int r = (int)(k / ((size_t)S * n));
    const double d = F[k] - Y[(size_t)i * R + r];
    acc += d * d;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// Fbar[r][s][i] = coef * (Y[i*R + r] - F[r][s][i])
__global__ void __launch_bounds__(256) gauss_bwd_kernel(const double* __restrict__ F,
                                                        const double* __restrict__ Y, int n, int S,
                                                        int R, const double* __restrict__ coef,
                                                        double* __restrict__ Fbar) {
  const size_t total = (size_t)R * S * n;
  const double c = coef[0];
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
       k += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(k % n);
    const int r = (int)(k / ((size_t)S * n));
    Fbar[k] = c * (Y[(size_t)i * R + r] - F[k]);
  }
}

// out = exp(x), 128-bit accesses (the sampling stage's second output for likelihoods that work on
// exp(f), notebooks/Abel_inversion.ipynb:279-283).  68 MB of traffic at the headline size: ~15 us,
// against +250 us when the same exp sat in the GEMM epilogue (all DMMA warps leave the tensor pipe
// idle while they run the ~60-instruction exp sequence 128 times each).
__global__ void __launch_bounds__(256) exp_kernel(const double* __restrict__ x, double* __restrict__ out, size_t count) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    const size_t n2 = count / 2;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    double2* o2 = reinterpret_cast<double2*>(out);
    for (size_t k = i; k < n2; k += stride) {
      const double2 v = x2[k];
      o2[k] = make_double2(exp(v.x), exp(v.y));
    }
    if (i == 0 && (count & 1)) out[count - 1] = exp(x[count - 1]);
  } else {
    for (size_t k = i; k < count; k += stride) out[k] = exp(x[k]);
  }
}

static inline int ew_grid(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = 148 * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ------------------------------------ host drivers ------------------------------------

int tril_unpack_f64(const double* src, double* dst, int n, int R, cudaStream_t st) {
  if (R > 16) return GPINV_ERR_ARG;
  dim3 grid(cdiv(n, TRIL_JC), n);
  tril_unpack_kernel<<<grid, R == 1 ? 128 : 256, R == 1 ? 0 : TRIL_JC * R * sizeof(double), st>>>(src, dst, n, R);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int tril_pack_f64(const double* src, double* dst, int n, int R, cudaStream_t st) {
  if (R > 16) return GPINV_ERR_ARG;
  dim3 grid(cdiv(n, TRIL_JC), n);
  tril_pack_kernel<<<grid, R == 1 ? 128 : 256, R == 1 ? 0 : TRIL_JC * R * sizeof(double), st>>>(src, dst, n, R);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int sumsq_f64(const double* x, size_t count, double* out, cudaStream_t st) {
  GPINV_TRY(cudaMemsetAsync(out, 0, sizeof(double), st));
  sumsq_kernel<<<ew_grid(count / 2 + 1), 256, 0, st>>>(x, count, out);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// U = q_mu + tril(Q) eps (GPinv/stvgp.py:162-168) and the partial sums of the stochastic KL
// (GPinv/stvgp.py:160-161,170-172).  kl_out: device [4] = {KL, sumE2, sumU2, logdet}; KL itself is
// written by kl_finish_f64 once sumU2 is complete.
int sample_u_fwd_ex_f64(const double* Q, int q_diag, const double* q_mu, const double* E, int n, int S, int R,
                        double* U, double* kl_out, int max_ctas, cudaStream_t st);

int sample_u_fwd_f64(const double* Q, int q_diag, const double* q_mu, const double* E, int n, int S, int R,
                     double* U, double* kl_out, cudaStream_t st) {
  return sample_u_fwd_ex_f64(Q, q_diag, q_mu, E, n, S, R, U, kl_out, 0, st);
}

// max_ctas > 0: the GEMM leaves SMs free (it runs beside the Cholesky factorisation, see sample_u_early_f64)
int sample_u_fwd_ex_f64(const double* Q, int q_diag, const double* q_mu, const double* E, int n, int S, int R,
                        double* U, double* kl_out, int max_ctas, cudaStream_t st) {
  GPINV_TRY(cudaMemsetAsync(kl_out, 0, 4 * sizeof(double), st));
  const size_t sn = (size_t)S * n;
  sumsq_kernel<<<ew_grid((size_t)R * sn / 2 + 1), 256, 0, st>>>(E, (size_t)R * sn, kl_out + 1);
  GPINV_CHECK_LAUNCH();
  diag_logsq_kernel<<<min(cdiv(n * R, 256), 1024), 256, 0, st>>>(Q, n, R, q_diag, kl_out + 3);
  GPINV_CHECK_LAUNCH();
  if (q_diag) {
    dim3 grid(ew_grid(sn), 1, R);
    udiag_fwd_kernel<<<grid, 256, 0, st>>>(Q, q_mu, E, n, S, U, kl_out + 2);
    GPINV_CHECK_LAUNCH();
    return GPINV_OK;
  }
  for (int r = 0; r < R; ++r) {
    GemmParams p = {};
    p.A = E + r * sn; p.lda = n;
    p.B = Q + (size_t)r * n * n; p.ldb = n;
    p.M = S; p.N = n; p.K = n;
    p.alpha = 1.0;
    p.colvec = q_mu + (size_t)r * n;
    p.out = U + r * sn; p.ldo = n;
    p.sumsq = kl_out + 2;
    p.kmode = 1; p.splitk = 1; p.allow_split = 1;
    p.max_ctas = max_ctas;
    const int rc = gemm_launch(LK, LK, EF_COLVEC | EF_SUMSQ, p, st);
    if (rc) return rc;
  }
  return GPINV_OK;
}

// helpers of the library's stream set (potrf.cu)
int side_early_begin(cudaStream_t* out);
int side_early_end();
int side_early_join(cudaStream_t st);

// U = E tril(Q)^T + q_mu and the KL partial sums do not depend on the Cholesky factor: this entry runs them on
// the library's low-priority stream, behind (a) the point marked with side_mark_inputs (their inputs exist)
// and (b) the middle of the most recent factorisation (from there on its trailing updates no longer fill the
// machine), so that the 0.5 ms product hides inside the factorisation's latency chain.  sample_fwd2_f64 with
// u_ready = 1 joins it and continues with F = sqrt(v) U Lc^T + mean.
int sample_u_early_f64(const double* Q, int q_diag, const double* q_mu, const double* E, int n, int S, int R,
                       double* U, double* kl_out, int max_ctas) {
  cudaStream_t s2 = nullptr;
  int rc = side_early_begin(&s2);
  if (rc) return rc;
  rc = sample_u_fwd_ex_f64(Q, q_diag, q_mu, E, n, S, R, U, kl_out, max_ctas, s2);
  if (rc) return rc;
  return side_early_end();
}

int kl_finish_f64(double* kl_out, int S, cudaStream_t st) {
  // parts layout for kl_finish: {sumE2, sumU2, logdet} = kl_out+1
  kl_finish_kernel<<<1, 1, 0, st>>>(kl_out + 1, (double)S, kl_out);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// kl_out: device [4] = {KL, sumE2, sumU2, logdet}
int sample_fwd2_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q, int q_diag,
                    const double* q_mu, const double* mean, const double* E, int n, int S, int R,
                    double* U, double* F, double* expF, double* kl_out, int u_ready, cudaStream_t st);

int sample_fwd_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q, int q_diag,
                   const double* q_mu, const double* mean, const double* E, int n, int S, int R,
                   double* U, double* F, double* expF, double* kl_out, cudaStream_t st) {
  return sample_fwd2_f64(Lc, ldl, sqrt_v, Q, q_diag, q_mu, mean, E, n, S, R, U, F, expF, kl_out, 0, st);
}

// u_ready: U and the partial sums in kl_out were produced by sample_u_early_f64
int sample_fwd2_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q, int q_diag,
                    const double* q_mu, const double* mean, const double* E, int n, int S, int R,
                    double* U, double* F, double* expF, double* kl_out, int u_ready, cudaStream_t st) {
  if (n <= 0 || S <= 0 || R <= 0) return GPINV_ERR_ARG;
  if (!F && !expF) return GPINV_ERR_ARG;
  int rc = u_ready ? side_early_join(st) : sample_u_fwd_f64(Q, q_diag, q_mu, E, n, S, R, U, kl_out, st);
  if (rc) return rc;
  const size_t sn = (size_t)S * n;
  for (int r = 0; r < R; ++r) {
    GemmParams p = {};
    p.A = U + r * sn; p.lda = n;
    p.B = Lc; p.ldb = ldl;
    p.M = S; p.N = n; p.K = n;
    p.alpha = 1.0; p.alpha_dev = sqrt_v + r;
    p.colvec = mean + (size_t)r * n;
    // exp(F) is produced by a streaming pass, not in the epilogue (see exp_kernel); when the caller
    // wants only exp(F) the GEMM writes F into that buffer and the pass runs in place
    double* Fr = F ? F + r * sn : expF + r * sn;
    p.out = Fr; p.ldo = n;
    p.kmode = 1; p.splitk = 1; p.allow_split = 1;
    rc = gemm_launch(LK, LK, EF_COLVEC, p, st);
    if (rc) return rc;
    if (expF) {
      exp_kernel<<<ew_grid(sn / 2 + 1), 256, 0, st>>>(Fr, expF + r * sn, sn);
      GPINV_CHECK_LAUNCH();
    }
  }
  return kl_finish_f64(kl_out, S, st);
}

int colsum_f64(const double* X, int n, int S, int R, double* out, cudaStream_t st) {
  GPINV_TRY(cudaMemsetAsync(out, 0, (size_t)R * n * sizeof(double), st));
  const int schunk = 64;
  dim3 grid(cdiv(n, 256), cdiv(S, schunk), R);
  colsum_kernel<<<grid, 256, 0, st>>>(X, n, S, schunk, out);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// Tail of the reverse pass that does not depend on the form of the Cholesky factor: from Ubar to
// qmu_bar (column sums, unless the caller's Ubar GEMM already produced them), Qbar = tril(Ubar^T E)
// (or the diagonal version) and the d/dQ_ii term of the stochastic KL.
int sample_q_bwd_f64(const double* Q, int q_diag, const double* E, const double* Ubar, const double* klbar,
                     int n, int S, int R, double* qmu_bar, int qmu_from_colsum, double* Qbar, cudaStream_t st) {
  const size_t sn = (size_t)S * n;
  int rc;
  if (qmu_from_colsum) {
    rc = colsum_f64(Ubar, n, S, R, qmu_bar, st);
    if (rc) return rc;
  }
  if (q_diag) {
    GPINV_TRY(cudaMemsetAsync(Qbar, 0, (size_t)R * n * sizeof(double), st));
    const int schunk = 64;
    dim3 grid(cdiv(n, 256), cdiv(S, schunk), R);
    udiag_bwd_kernel<<<grid, 256, 0, st>>>(Ubar, E, n, S, schunk, Qbar, nullptr);
    GPINV_CHECK_LAUNCH();
  } else {
    for (int r = 0; r < R; ++r) {   // Qbar_r = tril(Ubar_r^T E_r)
      GemmParams p = {};
      p.A = Ubar + r * sn; p.lda = n;
      p.B = E + r * sn; p.ldb = n;
      p.M = n; p.N = n; p.K = S;
      p.alpha = 1.0;
      p.out = Qbar + (size_t)r * n * n; p.ldo = n;
      p.kmode = 0; p.splitk = 1; p.allow_split = 1;
      rc = gemm_launch(LMN, LMN, EF_LOWER, p, st);
      if (rc) return rc;
    }
  }
  if (klbar) {
    diag_term_kernel<<<cdiv(n * R, 256), 256, 0, st>>>(Qbar, Q, n, R, q_diag, klbar, (double)S);
    GPINV_CHECK_LAUNCH();
  }
  return GPINV_OK;
}

// Reverse of sample_fwd.  klbar: device scalar dObj/dKL (nullable = 0).  Outputs:
//   Ubar [R,S,n] (workspace, returned for inspection), qmu_bar [R,n], Qbar [R,n,n] lower
//   (or [R,n] when q_diag), Lc_bar [n,n] lower (summed over r), svdot [R] =
//   sv_r <tril(Fbar_r^T U_r), Lc>, mean_bar [R,n].
int sample_bwd_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q, int q_diag,
                   const double* E, const double* U, const double* Fbar, const double* klbar,
                   int n, int S, int R, double* Ubar, double* qmu_bar, double* Qbar,
                   double* Lc_bar, int ldlb, double* svdot, double* mean_bar, cudaStream_t st) {
  if (n <= 0 || S <= 0 || R <= 0) return GPINV_ERR_ARG;
  const size_t sn = (size_t)S * n;
  GPINV_TRY(cudaMemsetAsync(qmu_bar, 0, (size_t)R * n * sizeof(double), st));
  GPINV_TRY(cudaMemsetAsync(svdot, 0, (size_t)R * sizeof(double), st));
  int rc = colsum_f64(Fbar, n, S, R, mean_bar, st);
  if (rc) return rc;
  for (int r = 0; r < R; ++r) {
    {  // Ubar_r = sv_r Fbar_r Lc + klbar U_r ; qmu_bar = column sums (fused)
      GemmParams p = {};
      p.A = Fbar + r * sn; p.lda = n;
      p.B = Lc; p.ldb = ldl;
      p.M = S; p.N = n; p.K = n;
      p.alpha = 1.0; p.alpha_dev = sqrt_v + r;
      p.beta = klbar ? 1.0 : 0.0; p.beta_dev = klbar;
      p.Cin = U + r * sn; p.ldcin = n;
      p.out = Ubar + r * sn; p.ldo = n;
      p.colsum = qmu_bar + (size_t)r * n;
      p.kmode = 2; p.splitk = 1; p.allow_split = 1;
      rc = gemm_launch(LK, LMN, EF_BETA | EF_COLSUM, p, st);
      if (rc) return rc;
    }
    {  // Lc_bar (+)= sv_r tril(Fbar_r^T U_r) ; svdot_r = <that, Lc>
      GemmParams p = {};
      p.A = Fbar + r * sn; p.lda = n;
      p.B = U + r * sn; p.ldb = n;
      p.M = n; p.N = n; p.K = S;
      p.alpha = 1.0; p.alpha_dev = sqrt_v + r;
      p.beta = (r == 0) ? 0.0 : 1.0;
      p.Cin = Lc_bar; p.ldcin = ldlb;
      p.out = Lc_bar; p.ldo = ldlb;
      p.W = Lc; p.ldw = ldl;
      p.dot = svdot + r;
      p.kmode = 0; p.splitk = 1; p.allow_split = 1;
      rc = gemm_launch(LMN, LMN, EF_LOWER | EF_BETA | EF_DOT, p, st);
      if (rc) return rc;
    }
  }
  return sample_q_bwd_f64(Q, q_diag, E, Ubar, klbar, n, S, R, qmu_bar, 0, Qbar, st);
}

int gauss_sumsq_f64(const double* F, const double* Y, int n, int S, int R, double* out,
                    cudaStream_t st) {
  GPINV_TRY(cudaMemsetAsync(out, 0, sizeof(double), st));
  gauss_sumsq_kernel<<<ew_grid((size_t)R * S * n), 256, 0, st>>>(F, Y, n, S, R, out);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int gauss_bwd_f64(const double* F, const double* Y, int n, int S, int R, const double* coef,
                  double* Fbar, cudaStream_t st) {
  gauss_bwd_kernel<<<ew_grid((size_t)R * S * n), 256, 0, st>>>(F, Y, n, S, R, coef, Fbar);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// Rs[S,M] = 1 y^T - expF[S,n] A[M,n]^T ; *sumsq = sum Rs^2
int abel_fwd_f64(const double* expF, const double* A, int lda, const double* Y, int n, int M, int S,
                 double* Rs, double* sumsq, cudaStream_t st) {
  GPINV_TRY(cudaMemsetAsync(sumsq, 0, sizeof(double), st));
  GemmParams p = {};
  p.A = expF; p.lda = n;
  p.B = A; p.ldb = lda;
  p.M = S; p.N = M; p.K = n;
  p.alpha = -1.0;
  p.colvec = Y;
  p.out = Rs; p.ldo = M;
  p.sumsq = sumsq;
  p.kmode = 0; p.splitk = 1; p.allow_split = 1;
  return gemm_launch(LK, LK, EF_COLVEC | EF_SUMSQ, p, st);
}

// Fbar[S,n] = coef * expF o (Rs[S,M] A[M,n])
int abel_bwd_f64(const double* Rs, const double* A, int lda, const double* expF, const double* coef,
                 int n, int M, int S, double* Fbar, cudaStream_t st) {
  GemmParams p = {};
  p.A = Rs; p.lda = M;
  p.B = A; p.ldb = lda;
  p.M = S; p.N = n; p.K = M;
  p.alpha = 1.0; p.alpha_dev = coef;
  p.mul = expF; p.ldmul = n;
  p.out = Fbar; p.ldo = n;
  p.kmode = 0; p.splitk = 1; p.allow_split = 1;
  return gemm_launch(LK, LMN, EF_MULMAT, p, st);
}

}  // namespace gpinv
