// extern "C" surface of libgpinv_sm100.so (declared in include/gpinv_b200.h).
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpinv_b200.h"
#include "gemm.cuh"

namespace gpinv {
static thread_local std::string g_last_error;
void set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }

// implemented in the other translation units
size_t kbuild_ws_bytes(int n, int n2, int D);
int kbuild_f64(const double*, const double*, int, int, int, int, int, const double*, int, int,
               double, double*, int, double*, cudaStream_t);
int kbuild_bwd_f64(const double*, const double*, int, int, int, int, int, const double*, int, int,
                   const double*, int, double*, double*, cudaStream_t);
int potrf_f64(double*, int, int, int*, cudaStream_t);
size_t potrf_bwd_ws_bytes(int);
int potrf_bwd_f64(const double*, int, double*, int, int, double*, cudaStream_t);
int trsm_rlt_f64(const double*, int, double*, int, int, int, cudaStream_t);
size_t potrf_bwd_inv_ws_bytes(int);
int potrf_bwd_inv_f64(const double*, int, double*, int, int, double*, cudaStream_t);
int potrf_bwd_hybrid_f64(const double*, int, double*, int, int, double*, cudaStream_t);
int potrf_bwd_hybrid_block();
int potrf_bwd_hybrid_ex_f64(const double*, int, double*, int, int, double*, int, int, cudaStream_t,
                            const double* W_pre = nullptr);
int trtri_side_f64(const double*, int, double*, double*, int, cudaStream_t);
int trtri_side_shard_f64(const double*, int, double*, double*, int, int, int, cudaStream_t);
int trtri_side_join_f64(cudaStream_t);
int trtri_side_sync_f64();
int potrf_ex_f64(double*, int, int, int*, int, cudaStream_t, double* Winv = nullptr, double* Tinv = nullptr,
                 int* inv_done_out = nullptr);
int tril_copy_f64(const double*, int, double*, int, int, cudaStream_t);
int lbar_pack_f64(const double*, int, int, int, int, double*, cudaStream_t);
int tri_compact_f64(const double*, int, int, double*, cudaStream_t);
int tri_expand_f64(const double*, double*, int, int, cudaStream_t);
int potrf_bwd_shard_f64(const double*, int, const double*, int, int, int, const int*, int, const double*, double*,
                        double*, double*, cudaStream_t);
int kbuild_ex_f64(const double*, const double*, int, int, int, int, int, const double*, int, int, double,
                  double*, int, double*, int, cudaStream_t);
int kbuild_bwd_ex_f64(const double*, const double*, int, int, int, int, int, const double*, int, int,
                      const double*, int, double*, double*, int, cudaStream_t);
int tril_unpack_f64(const double*, double*, int, int, cudaStream_t);
int tril_pack_f64(const double*, double*, int, int, cudaStream_t);
int sumsq_f64(const double*, size_t, double*, cudaStream_t);
int side_mark_inputs(cudaStream_t);
int side_early_join(cudaStream_t);
int sample_u_early_f64(const double*, int, const double*, const double*, int, int, int, double*, double*, int);
int sample_fwd2_f64(const double*, int, const double*, const double*, int, const double*, const double*,
                    const double*, int, int, int, double*, double*, double*, double*, int, cudaStream_t);
int sample_fwd_f64(const double*, int, const double*, const double*, int, const double*,
                   const double*, const double*, int, int, int, double*, double*, double*, double*,
                   cudaStream_t);
int sample_bwd_f64(const double*, int, const double*, const double*, int, const double*,
                   const double*, const double*, const double*, int, int, int, double*, double*,
                   double*, double*, int, double*, double*, cudaStream_t);
int gauss_sumsq_f64(const double*, const double*, int, int, int, double*, cudaStream_t);
int gauss_bwd_f64(const double*, const double*, int, int, int, const double*, double*,
                  cudaStream_t);
int abel_fwd_f64(const double*, const double*, int, const double*, int, int, int, double*, double*,
                 cudaStream_t);
int abel_bwd_f64(const double*, const double*, int, const double*, const double*, int, int, int,
                 double*, cudaStream_t);
int kron_apply_f64(const double*, int, int, const double*, int, int, const double*, int, double, const double*,
                   const double*, double*, double*, double*, int, cudaStream_t);
size_t kron_bwd_ws_bytes(int, int, int, int);
int kron_sample_fwd_f64(const double*, int, const double*, int, const double*, const double*, int, const double*,
                        const double*, const double*, int, int, double*, double*, double*, double*, double*,
                        cudaStream_t);
int kron_sample_bwd_f64(const double*, int, const double*, int, const double*, const double*, int, const double*,
                        const double*, const double*, const double*, const double*, int, int, double*, double*,
                        double*, double*, double*, double*, double*, double*, cudaStream_t);
int unpack_f64(const double*, int, const long long*, const int*, double* const*, cudaStream_t);
int pack_f64(const double*, int, const long long*, const int*, const double* const*, double, double*, cudaStream_t);
int kl_white_f64(const double*, const double*, int, int, int, double*, cudaStream_t);
int kl_white_bwd_f64(const double*, const double*, int, int, int, const double*, double*, double*, cudaStream_t);
int adam_step_f64(double*, double*, double*, const double*, long long, double, double, double, double, long long*,
                  cudaStream_t);
int gauss_logp_f64(const double*, const double*, double, double*, cudaStream_t);
int gauss_logp_bwd_f64(const double*, const double*, double, const double*, double*, double*, cudaStream_t);
int spec_fwd_f64(const double*, const double*, const double*, const double*, const double*, int, int,
                 int, int, double, double*, double*, cudaStream_t);
int spec_bwd_f64(const double*, const double*, const double*, const double*, const double*,
                 const double*, const double*, int, int, int, int, double, double*, cudaStream_t);
}  // namespace gpinv

using namespace gpinv;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int gpinv_version(void) { return 101; }

int gpinv_set_workspace(void* ptr, size_t bytes) {
  gemm_set_workspace(ptr, bytes);
  return GPINV_OK;
}
const char* gpinv_last_error(void) { return g_last_error.c_str(); }

int gpinv_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch,
                         size_t width_bytes, size_t height, int kind, void* stream) {
  if (width_bytes == 0 || height == 0) return GPINV_OK;
  if (!dst || !src || dpitch < width_bytes || spitch < width_bytes || kind < 1 || kind > 3) return GPINV_ERR_ARG;
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice
                         : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  const cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, k, ST(stream));
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e));
    return GPINV_ERR_CUDA;
  }
  return GPINV_OK;
}

size_t gpinv_kbuild_ws_bytes(int n, int n2, int D) { return kbuild_ws_bytes(n, n2, D); }

int gpinv_kbuild_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
                     const double* ls, int ard, int mode, double jitter, double* K, int ldk,
                     void* ws, void* stream) {
  return kbuild_f64(X, X2, n, n2, D, ldx, ldx2, ls, ard, mode, jitter, K, ldk,
                    static_cast<double*>(ws), ST(stream));
}

int gpinv_kbuild_bwd_f64(const double* X, const double* X2, int n, int n2, int D, int ldx,
                         int ldx2, const double* ls, int ard, int mode, const double* Kbar,
                         int ldkb, double* ls_bar, void* ws, void* stream) {
  return kbuild_bwd_f64(X, X2, n, n2, D, ldx, ldx2, ls, ard, mode, Kbar, ldkb, ls_bar,
                        static_cast<double*>(ws), ST(stream));
}

int gpinv_potrf_f64(double* A, int n, int lda, int* info, void* stream) {
  if (n <= 0 || lda < n) return GPINV_ERR_ARG;
  return potrf_f64(A, n, lda, info, ST(stream));
}

// GPINV_POTRF_BWD = murray | inv | hybrid (default): which reverse sweep of the Cholesky runs
static int bwd_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GPINV_POTRF_BWD");
    v = (e && strcmp(e, "murray") == 0) ? 1 : (e && strcmp(e, "inv") == 0) ? 2 : 0;
  }
  return v;
}

size_t gpinv_potrf_bwd_ws_bytes(int n) {
  const size_t a = potrf_bwd_ws_bytes(n), b = potrf_bwd_inv_ws_bytes(n);
  return a > b ? a : b;
}

int gpinv_potrf_bwd_f64(const double* L, int ldl, double* Abar, int ldab, int n, void* ws,
                        void* stream) {
  if (n <= 0 || ldl < n || ldab < n) return GPINV_ERR_ARG;
  const int mode = bwd_mode();
  if (mode == 1) return potrf_bwd_f64(L, ldl, Abar, ldab, n, static_cast<double*>(ws), ST(stream));
  if (mode == 0)
    return potrf_bwd_hybrid_f64(L, ldl, Abar, ldab, n, static_cast<double*>(ws), ST(stream));
  return potrf_bwd_inv_f64(L, ldl, Abar, ldab, n, static_cast<double*>(ws), ST(stream));
}

// ---- host side of the lower-trapezoid upload (Model._objective host boundary) ----
// x_host: pageable row-major [nq,nq] matrix (q_sqrt inside the optimiser's state vector).  Row
// blocks of `rb` rows are gathered -- columns [0, r1) of rows [r0, r1) -- into the pinned buffer at
// element offset off[k] by `nthreads` worker threads (chunks of 8 rows, handed out in block
// order), and the calling thread issues the block's contiguous host->device copy on `stream` as
// soon as the block is complete.  Returns when every copy has been ISSUED.  take[k] == 0 skips a
// block (another rank stages it).
int gpinv_upload_tril_blocks_f64(const double* x_host, int nq, int rb, double* pinned, double* dstage,
                                 const long long* off, const int* take, int nblocks, int nthreads,
                                 void* stream) {
  // dstage == NULL: gather only (no copies are issued; used by the CPU-side test of the packing)
  if (!x_host || !pinned || !off || !take || nq <= 0 || rb <= 0 || nblocks <= 0) return GPINV_ERR_ARG;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  constexpr int CH = 8;   // rows per work item
  // work items in block order
  std::vector<int> item_block, item_row;
  std::vector<int> block_items(nblocks, 0);
  for (int k = 0; k < nblocks; ++k) {
    if (!take[k]) continue;
    const int r0 = k * rb, r1 = (r0 + rb < nq) ? r0 + rb : nq;
    for (int r = r0; r < r1; r += CH) {
      item_block.push_back(k);
      item_row.push_back(r);
      ++block_items[k];
    }
  }
  const int nitems = (int)item_block.size();
  std::vector<std::atomic<int>> done(nblocks);
  for (int k = 0; k < nblocks; ++k) done[k].store(0);
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (;;) {
      const int it = next.fetch_add(1);
      if (it >= nitems) break;
      const int k = item_block[it], ra = item_row[it];
      const int r0 = k * rb, r1 = (r0 + rb < nq) ? r0 + rb : nq;
      const int rz = (ra + CH < r1) ? ra + CH : r1;
      const size_t width = (size_t)r1;                 // columns [0, r1)
      for (int r = ra; r < rz; ++r)
        std::memcpy(pinned + off[k] + (size_t)(r - r0) * width, x_host + (size_t)r * nq, width * sizeof(double));
      done[k].fetch_add(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
  int rc = GPINV_OK;
  for (int k = 0; k < nblocks; ++k) {
    if (!take[k]) continue;
    while (done[k].load(std::memory_order_acquire) < block_items[k]) std::this_thread::sleep_for(std::chrono::microseconds(20));
    const int r0 = k * rb, r1 = (r0 + rb < nq) ? r0 + rb : nq;
    const size_t bytes = (size_t)(r1 - r0) * r1 * sizeof(double);
    if (rc == GPINV_OK && dstage &&
        cudaMemcpyAsync(dstage + off[k], pinned + off[k], bytes, cudaMemcpyHostToDevice, ST(stream)) != cudaSuccess) {
      set_last_error(cudaGetErrorString(cudaGetLastError()));
      rc = GPINV_ERR_CUDA;
    }
  }
  for (auto& th : pool) th.join();
  return rc;
}

// ---- fused kernel build + Cholesky (the Kern.Cholesky core of GPinv/kernels.py:56-59) ----
size_t gpinv_chol_core_ws_bytes(int n, int D) { return kbuild_ws_bytes(n, n, D); }

int gpinv_chol_core_f64(const double* X, int n, int D, int ldx, const double* ls, int ard, int mode,
                        double jitter, double* L, int ldl, int* info, void* ws, void* stream) {
  if (n <= 0 || D <= 0 || ldl < n) return GPINV_ERR_ARG;
  int rc = kbuild_ex_f64(X, nullptr, n, n, D, ldx, ldx, ls, ard, mode, jitter, L, ldl,
                         static_cast<double*>(ws), 1, ST(stream));
  if (rc) return rc;
  return potrf_ex_f64(L, n, ldl, info, 1, ST(stream));
}

int gpinv_chol_core_inv_f64(const double* X, int n, int D, int ldx, const double* ls, int ard, int mode,
                            double jitter, double* L, int ldl, int* info, double* Winv, double* scratch, void* ws,
                            void* stream) {
  if (n <= 0 || D <= 0 || ldl < n || !Winv || !scratch) return GPINV_ERR_ARG;
  int rc = kbuild_ex_f64(X, nullptr, n, n, D, ldx, ldx, ls, ard, mode, jitter, L, ldl,
                         static_cast<double*>(ws), 1, ST(stream));
  if (rc) return rc;
  int interleaved = 0;
  rc = potrf_ex_f64(L, n, ldl, info, 1, ST(stream), Winv, scratch, &interleaved);
  if (rc || interleaved) return rc;
  return trtri_side_f64(L, ldl, Winv, scratch, n, ST(stream));     // small n: behind the factorisation
}

int gpinv_trtri_side_f64(const double* L, int ldl, double* Winv, double* scratch, int n, void* stream) {
  if (n <= 0 || ldl < n || !Winv || !scratch) return GPINV_ERR_ARG;
  return trtri_side_f64(L, ldl, Winv, scratch, n, ST(stream));
}

int gpinv_trtri_side_shard_f64(const double* L, int ldl, double* Winv, double* scratch, int n, int rank, int G,
                               void* stream) {
  if (n <= 0 || ldl < n || !Winv || !scratch) return GPINV_ERR_ARG;
  return trtri_side_shard_f64(L, ldl, Winv, scratch, n, rank, G, ST(stream));
}

int gpinv_trtri_side_join(void* stream) { return trtri_side_join_f64(ST(stream)); }
int gpinv_trtri_side_sync(void) { return trtri_side_sync_f64(); }

size_t gpinv_chol_core_bwd_ws_bytes(int n, int D) {
  return (size_t)n * n * sizeof(double) + potrf_bwd_inv_ws_bytes(n) + kbuild_ws_bytes(n, n, D);
}

int gpinv_chol_core_bwd_f64(const double* X, int n, int D, int ldx, const double* ls, int ard,
                            int mode, const double* L, int ldl, const double* Lbar, int ldlb,
                            double* ls_bar, void* ws, void* stream) {
  return gpinv_chol_core_bwd2_f64(X, n, D, ldx, ls, ard, mode, L, ldl, Lbar, ldlb, nullptr, 0, ls_bar, ws, stream);
}

int gpinv_chol_core_bwd2_f64(const double* X, int n, int D, int ldx, const double* ls, int ard,
                             int mode, const double* L, int ldl, const double* Lbar, int ldlb,
                             const double* Winv, int winv_ready, double* ls_bar, void* ws, void* stream) {
  if (n <= 0 || D <= 0 || ldl < n || ldlb < n) return GPINV_ERR_ARG;
  double* Ab = static_cast<double*>(ws);              // n x n: tril(Lbar), then the Murray-lower Kbar
  double* sweep_ws = Ab + (size_t)n * n;              // potrf_bwd_inv_ws_bytes(n)
  double* kws = sweep_ws + potrf_bwd_inv_ws_bytes(n) / sizeof(double);
  int rc = tril_copy_f64(Lbar, ldlb, Ab, n, n, ST(stream));
  if (rc) return rc;
  if (Winv && !winv_ready) {
    rc = trtri_side_join_f64(ST(stream));
    if (rc) return rc;
  }
  rc = potrf_bwd_hybrid_ex_f64(L, ldl, Ab, n, n, sweep_ws, 1, 0, ST(stream), Winv);
  if (rc) return rc;
  return kbuild_bwd_ex_f64(X, nullptr, n, n, D, ldx, ldx, ls, ard, mode, Ab, n, ls_bar, kws, 1, ST(stream));
}

int gpinv_trsm_rlt_f64(const double* L, int ldl, double* X, int ldx, int m, int n, void* stream) {
  if (n <= 0 || ldl < n || ldx < n) return GPINV_ERR_ARG;
  return trsm_rlt_f64(L, ldl, X, ldx, m, n, ST(stream));
}

int gpinv_gemm_f64(int transa, int transb, int m, int n, int k, double alpha, const double* A,
                   int lda, const double* B, int ldb, double beta, double* C, int ldc, int lower,
                   void* stream) {
  // row-major: op(A) is m x k.  'N': A stored m x k (k contiguous -> LK);  'T': stored k x m (LMN).
  // op(B) is k x n.  'N': B stored k x n (n contiguous -> LMN);  'T': stored n x k (LK).
  const int la = transa ? LMN : LK;
  const int lb = transb ? LK : LMN;
  GemmParams p = {};
  p.A = A; p.B = B; p.M = m; p.N = n; p.K = k; p.lda = lda; p.ldb = ldb;
  p.alpha = alpha; p.beta = beta; p.Cin = C; p.ldcin = ldc; p.out = C; p.ldo = ldc;
  p.splitk = 1; p.allow_split = 1;     // the launcher may k-split through the registered workspace
  return gemm_launch(la, lb, EF_BETA | (lower ? EF_LOWER : 0u), p, ST(stream));
}

// Test / bring-up entry: the DMMA GEMM with every knob of the library-internal launcher exposed
// (layouts, triangular k-range mode, lower output, batching, stream-K permission).
int gpinv_gemm_ex_f64(int la, int lb, int m, int n, int k, double alpha, const double* A, int lda,
                      long long strideA, const double* B, int ldb, long long strideB, double beta, double* C,
                      int ldc, long long strideC, int batch, int lower, int kmode, int allow_split, void* stream) {
  GemmParams p = {};
  p.A = A; p.B = B; p.M = m; p.N = n; p.K = k; p.lda = lda; p.ldb = ldb;
  p.alpha = alpha; p.beta = beta; p.Cin = C; p.ldcin = ldc; p.out = C; p.ldo = ldc;
  p.kmode = kmode; p.splitk = 1; p.allow_split = allow_split;
  p.batch = batch; p.strideA = strideA; p.strideB = strideB; p.strideO = strideC; p.strideCin = strideC;
  return gemm_launch(la, lb, EF_BETA | (lower ? EF_LOWER : 0u), p, ST(stream));
}

int gpinv_tril_unpack_f64(const double* src, double* dst, int n, int R, void* stream) {
  return tril_unpack_f64(src, dst, n, R, ST(stream));
}
int gpinv_tril_pack_f64(const double* src, double* dst, int n, int R, void* stream) {
  return tril_pack_f64(src, dst, n, R, ST(stream));
}

int gpinv_sample_fwd_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q,
                         int q_diag, const double* q_mu, const double* mean, const double* E,
                         int n, int S, int R, double* U, double* F, double* expF, double* kl_out,
                         void* stream) {
  return sample_fwd_f64(Lc, ldl, sqrt_v, Q, q_diag, q_mu, mean, E, n, S, R, U, F, expF, kl_out,
                        ST(stream));
}

int gpinv_side_mark_inputs(void* stream) { return side_mark_inputs(ST(stream)); }
int gpinv_side_early_join(void* stream) { return side_early_join(ST(stream)); }
int gpinv_sample_u_early_f64(const double* Q, int q_diag, const double* q_mu, const double* E, int n, int S, int R,
                             double* U, double* kl_out, int max_ctas) {
  if (n <= 0 || S <= 0 || R <= 0 || !Q || !q_mu || !E || !U || !kl_out) return GPINV_ERR_ARG;
  return sample_u_early_f64(Q, q_diag, q_mu, E, n, S, R, U, kl_out, max_ctas);
}
int gpinv_sample_fwd2_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q, int q_diag,
                          const double* q_mu, const double* mean, const double* E, int n, int S, int R, double* U,
                          double* F, double* expF, double* kl_out, int u_ready, void* stream) {
  return sample_fwd2_f64(Lc, ldl, sqrt_v, Q, q_diag, q_mu, mean, E, n, S, R, U, F, expF, kl_out, u_ready, ST(stream));
}

int gpinv_sample_bwd_f64(const double* Lc, int ldl, const double* sqrt_v, const double* Q,
                         int q_diag, const double* E, const double* U, const double* Fbar,
                         const double* klbar, int n, int S, int R, double* Ubar, double* qmu_bar,
                         double* Qbar, double* Lc_bar, int ldlb, double* svdot, double* mean_bar,
                         void* stream) {
  return sample_bwd_f64(Lc, ldl, sqrt_v, Q, q_diag, E, U, Fbar, klbar, n, S, R, Ubar, qmu_bar,
                        Qbar, Lc_bar, ldlb, svdot, mean_bar, ST(stream));
}

int gpinv_gauss_sumsq_f64(const double* F, const double* Y, int n, int S, int R, double* out,
                          void* stream) {
  return gauss_sumsq_f64(F, Y, n, S, R, out, ST(stream));
}
int gpinv_gauss_bwd_f64(const double* F, const double* Y, int n, int S, int R, const double* coef,
                        double* Fbar, void* stream) {
  return gauss_bwd_f64(F, Y, n, S, R, coef, Fbar, ST(stream));
}

int gpinv_abel_fwd_f64(const double* expF, const double* A, int lda, const double* Y, int n, int M,
                       int S, double* Rs, double* sumsq, void* stream) {
  return abel_fwd_f64(expF, A, lda, Y, n, M, S, Rs, sumsq, ST(stream));
}
int gpinv_abel_bwd_f64(const double* Rs, const double* A, int lda, const double* expF,
                       const double* coef, int n, int M, int S, double* Fbar, void* stream) {
  return abel_bwd_f64(Rs, A, lda, expF, coef, n, M, S, Fbar, ST(stream));
}

int gpinv_spec_fwd_f64(const double* F, const double* A, const double* cosT, const double* lam,
                       const double* Y, int n, int m, int k, int S, double dlam, double* P, double* sumsq,
                       void* stream) {
  return spec_fwd_f64(F, A, cosT, lam, Y, n, m, k, S, dlam, P, sumsq, ST(stream));
}
int gpinv_spec_bwd_f64(const double* F, const double* A, const double* cosT, const double* lam,
                       const double* Y, const double* P, const double* coef, int n, int m, int k,
                       int S, double dlam, double* Fbar, void* stream) {
  return spec_bwd_f64(F, A, cosT, lam, Y, P, coef, n, m, k, S, dlam, Fbar, ST(stream));
}

int gpinv_kron_apply_f64(const double* L1, int n1, const double* L2, int n2, const double* U, int batch,
                         double* T_ws, double* F, int trans, void* stream) {
  if (!L1 || !L2 || !U || !T_ws || !F) return GPINV_ERR_ARG;
  return kron_apply_f64(L1, n1, n1, L2, n2, n2, U, batch, 1.0, nullptr, nullptr, T_ws, F, nullptr, trans,
                        ST(stream));
}

size_t gpinv_kron_bwd_ws_bytes(int n1, int n2, int S, int R) { return kron_bwd_ws_bytes(n1, n2, S, R); }

int gpinv_kron_sample_fwd_f64(const double* L1, int n1, const double* L2, int n2, const double* sqrt_v,
                              const double* Q, int q_diag, const double* q_mu, const double* mean,
                              const double* E, int S, int R, double* U, double* T, double* F, double* expF,
                              double* kl_out, void* stream) {
  return kron_sample_fwd_f64(L1, n1, L2, n2, sqrt_v, Q, q_diag, q_mu, mean, E, S, R, U, T, F, expF, kl_out,
                             ST(stream));
}

int gpinv_kron_sample_bwd_f64(const double* L1, int n1, const double* L2, int n2, const double* sqrt_v,
                              const double* Q, int q_diag, const double* E, const double* U, const double* T,
                              const double* Fbar, const double* klbar, int S, int R, double* Ubar,
                              double* qmu_bar, double* Qbar, double* L1bar, double* L2bar, double* svdot,
                              double* mean_bar, void* ws, void* stream) {
  return kron_sample_bwd_f64(L1, n1, L2, n2, sqrt_v, Q, q_diag, E, U, T, Fbar, klbar, S, R, Ubar, qmu_bar, Qbar,
                             L1bar, L2bar, svdot, mean_bar, static_cast<double*>(ws), ST(stream));
}

int gpinv_unpack_f64(const double* x, int nblocks, const long long* off, const int* transform,
                     double* const* dst, void* stream) {
  if (!x || !off || !transform || !dst) return GPINV_ERR_ARG;
  return unpack_f64(x, nblocks, off, transform, dst, ST(stream));
}

int gpinv_pack_f64(const double* x, int nblocks, const long long* off, const int* transform,
                   const double* const* gval, double scale, double* gx, void* stream) {
  if (!x || !off || !transform || !gval || !gx) return GPINV_ERR_ARG;
  return pack_f64(x, nblocks, off, transform, gval, scale, gx, ST(stream));
}

int gpinv_kl_white_f64(const double* q_mu, const double* q_sqrt, int n, int R, int q_diag, double* kl_out,
                       void* stream) {
  return kl_white_f64(q_mu, q_sqrt, n, R, q_diag, kl_out, ST(stream));
}

int gpinv_kl_white_bwd_f64(const double* q_mu, const double* q_sqrt, int n, int R, int q_diag,
                           const double* gbar, double* mu_bar, double* q_bar, void* stream) {
  return kl_white_bwd_f64(q_mu, q_sqrt, n, R, q_diag, gbar, mu_bar, q_bar, ST(stream));
}

int gpinv_adam_step_f64(double* x, double* m, double* v, const double* g, long long count, double lr,
                        double beta1, double beta2, double epsilon, long long* step, void* stream) {
  if (!x || !m || !v || !g || !step) return GPINV_ERR_ARG;
  return adam_step_f64(x, m, v, g, count, lr, beta1, beta2, epsilon, step, ST(stream));
}

int gpinv_gauss_logp_f64(const double* ss, const double* var, double count, double* out, void* stream) {
  if (!ss || !var || !out) return GPINV_ERR_ARG;
  return gauss_logp_f64(ss, var, count, out, ST(stream));
}
int gpinv_gauss_logp_bwd_f64(const double* ss, const double* var, double count, const double* g, double* ss_bar,
                             double* var_bar, void* stream) {
  if (!ss || !var || !g || !ss_bar || !var_bar) return GPINV_ERR_ARG;
  return gauss_logp_bwd_f64(ss, var, count, g, ss_bar, var_bar, ST(stream));
}

int gpinv_lbar_pack_f64(const double* Lbar, int ldlb, int n, int w, int G, double* out, void* stream) {
  if (!Lbar || !out) return GPINV_ERR_ARG;
  return lbar_pack_f64(Lbar, ldlb, n, w, G, out, ST(stream));
}

int gpinv_tri_compact_f64(const double* src, int ld, int n, double* dst, void* stream) {
  if (!src || !dst) return GPINV_ERR_ARG;
  return tri_compact_f64(src, ld, n, dst, ST(stream));
}

int gpinv_tri_expand_f64(const double* src, double* dst, int ld, int n, void* stream) {
  if (!src || !dst) return GPINV_ERR_ARG;
  return tri_expand_f64(src, dst, ld, n, ST(stream));
}

int gpinv_chol_core_bwd_shard_f64(const double* X, int n, int D, int ldx, const double* ls, int ard, int mode,
                                  const double* L, int ldl, const double* Winv, const double* Dsh, int w,
                                  const int* c0s, int nblk, double* ls_bar, void* ws, void* stream) {
  if (n <= 0 || D <= 0 || ldl < n || !Winv || !Dsh || !c0s || nblk <= 0 || w <= 0 || 2 * w > n) return GPINV_ERR_ARG;
  double* Gacc = static_cast<double*>(ws);            // n x n partial of W^T Phi W
  double* P = Gacc + (size_t)n * n;                   // n x w
  double* A = P + (size_t)n * w;                      // n x w
  double* kws = A + (size_t)n * w;                    // (fits: 2 n w <= n^2 < the sweep's share of the workspace)
  // (the caller has ordered `stream` after the side-stream inversion: gpinv_trtri_side_join)
  int rc = potrf_bwd_shard_f64(L, ldl, Winv, n, n, w, c0s, nblk, Dsh, Gacc, P, A, ST(stream));
  if (rc) return rc;
  return kbuild_bwd_ex_f64(X, nullptr, n, n, D, ldx, ldx, ls, ard, mode, Gacc, n, ls_bar, kws, 0, ST(stream));
}

int gpinv_sumsq_f64(const double* x, size_t count, double* out, void* stream) {
  return sumsq_f64(x, count, out, ST(stream));
}

}  // extern "C"
