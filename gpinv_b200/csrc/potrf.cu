// Blocked fp64 Cholesky (T4) and its reverse-mode sweep for sm_100a.
//
// Forward: two-level right-looking.  Outer panels of NBO columns get a DMMA SYRK trailing
// update (K = NBO); inside a panel, 64-column sub-panels are factored by `panel_kernel`
// (diagonal block potrf fused with the substitution-form TRSM of all rows below; every CTA
// re-factors the 64x64 diagonal block in shared memory so there is no inter-CTA dependency)
// followed by a K=64 DMMA update of the remaining panel columns.
//
// Reverse: Murray's level-3 blocked reverse sweep (arXiv:1602.07527, Alg. "blocked
// reverse-mode"), two-level: NBO-wide outer steps whose GEMMs run on the DMMA kernel,
// 64-wide inner steps with a substitution TRSM leaf and a level-2 leaf reverse in smem.
// Replaces tf.cholesky + its CholeskyGrad op at GPinv/kernels.py:59 and
// GPinv/kronecker.py:54,57 (reference /root/reference).
#include <climits>
#include <cstdlib>

#include "gemm.cuh"

namespace gpinv {

namespace pcfg {
constexpr int NB = 64;      // leaf width
constexpr int NBO = 256;    // outer panel width
constexpr int RB = 64;      // TRSM rows per CTA in panel_kernel
constexpr int LDR = 65;     // row buffer stride (odd -> conflict-free thread-per-row access)
constexpr int LDT = 66;     // transposed / natural L block stride (even -> 16B-aligned rows)
constexpr int PANEL_SMEM = ((NB + RB) * LDR + NB * LDT + 8 * 9) * 8;
constexpr int TRSM_ROWS = 128;
constexpr int TRSM_SMEM = (NB * LDT + NB + TRSM_ROWS * LDR) * 8;
constexpr int LEAF_SMEM = (2 * NB * LDR) * 8;
}  // namespace pcfg

// ---------------------------------------------------------------------------------------
// panel_kernel: factor the b x b diagonal block at (i0,i0) and solve X L11^T = A21 for the
// rows below.  Thread-per-row, left-looking in 8-column register blocks.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(2 * (pcfg::NB + pcfg::RB)) panel_kernel(double* __restrict__ A, int lda,
                                                                          int n, int i0, int b,
                                                                          int* __restrict__ info) {
  using namespace pcfg;
  extern __shared__ __align__(16) double sm[];
  double* Rs = sm;                      // [NB+RB][LDR]
  double* Lt = Rs + (NB + RB) * LDR;    // [NB][LDT]   Lt[k][c] = L11[c][k]
  double* Ds = Lt + NB * LDT;           // [8][9]
  const int tid = threadIdx.x;
  constexpr int ROWS = NB + RB;         // rows held by the CTA: 64 of the diagonal block + 64 below
  constexpr int NTHR = 2 * ROWS;        // every row has an owner thread and a helper thread
  const int row_base = i0 + b + blockIdx.x * RB;
  const int nrows = max(0, min(RB, n - row_base));

  // L2-only (ld.global.cg) loads: the look-ahead runs this kernel concurrently with the
  // side-stream trailing update, so L1 must not be trusted.  All 16 loads a thread owes to each
  // of the two blocks are issued back to back (one exposed L2 round trip per block).
  {
    constexpr int PER = NB * NB / NTHR;   // 16 elements per thread per block
    double v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = u * NTHR + tid;
      const int r = idx >> 6, c = idx & 63;
      v[u] = (r == c) ? 1.0 : 0.0;
      if (r < b && c < b) v[u] = __ldcg(A + (size_t)(i0 + r) * lda + i0 + c);
    }
    double w[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = u * NTHR + tid;
      const int r = idx >> 6, c = idx & 63;
      w[u] = 0.0;
      if (r < nrows && c < b) w[u] = __ldcg(A + (size_t)(row_base + r) * lda + i0 + c);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = u * NTHR + tid;
      Rs[(idx >> 6) * LDR + (idx & 63)] = v[u];
      Rs[(NB + (idx >> 6)) * LDR + (idx & 63)] = w[u];
    }
  }
  __syncthreads();
  // Ticket: the diagonal block is overwritten in place, so only the CTA that is LAST to have
  // loaded it may write the factor back (CTAs can start arbitrarily late when another stream
  // shares the GPU).  info[1] is the ticket counter; the last CTA resets it.
  __shared__ int s_last;
  if (tid == 0) {
    const int ticket = atomicAdd(info + 1, 1);
    s_last = (ticket == (int)gridDim.x - 1);
    if (s_last) info[1] = 0;
  }

  // Left-looking sweep over 8-column blocks with two threads per row.  The OWNER of a row
  // (tid < ROWS) brings the current block up to date with the previous block's 8 columns,
  // factors the 8x8 pivot block (redundantly, in registers) and solves its row's 8 entries.
  // Meanwhile the HELPER of the row (tid >= ROWS) applies all columns finished BEFORE the current
  // block to the NEXT block's 8 entries: that is the long part of the left-looking update and it
  // does not depend on the pivot chain the owners are waiting on, so the two overlap.
  const bool owner = tid < ROWS;
  const int row = owner ? tid : tid - ROWS;
  int fail = INT_MAX;
  double* myrow = Rs + row * LDR;
#pragma unroll 1
  for (int cb = 0; cb < NB / 8; ++cb) {
    const int c0 = cb * 8;
    double s[8];
    bool active = false;
    if (owner) {
      active = (row >= c0);
      if (active) {
#pragma unroll
        for (int q = 0; q < 8; ++q) s[q] = myrow[c0 + q];
        // columns [c0-8, c0): finished during the previous block (everything older was applied
        // by the helper while that block was being factored)
        if (c0 > 0) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {      // exactly 8 columns: unrolled, loads ahead of the FMAs
            const int k = c0 - 8 + kk;
            const double li = myrow[k];
            const double2* lt = reinterpret_cast<const double2*>(Lt + k * LDT + c0);
            const double2 l01 = lt[0], l23 = lt[1], l45 = lt[2], l67 = lt[3];
            s[0] -= li * l01.x; s[1] -= li * l01.y; s[2] -= li * l23.x; s[3] -= li * l23.y;
            s[4] -= li * l45.x; s[5] -= li * l45.y; s[6] -= li * l67.x; s[7] -= li * l67.y;
          }
        }
      }
      if (row >= c0 && row < c0 + 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) Ds[(row - c0) * 9 + q] = s[q];
      }
    }
    __syncthreads();
    if (owner) {
      // every owner factors the 8x8 diagonal block redundantly in registers, right-looking so
      // the dependent chain per pivot is rsqrt -> scale -> one FMA
      double l[8][8], inv[8];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int bb = 0; bb <= a; ++bb) l[a][bb] = Ds[a * 9 + bb];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double v = l[j][j];
        if (!(v > 0.0)) fail = min(fail, c0 + j);
        const double y = fast_rsqrt(v);
        inv[j] = y;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) l[i][j] *= y;
#pragma unroll
        for (int i = j + 1; i < 8; ++i)
#pragma unroll
          for (int k = j + 1; k <= i; ++k) l[i][k] -= l[i][j] * l[k][j];
        l[j][j] = sqrt_from_rsqrt(v, y);
      }
      if (active) {
        double x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = s[q];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          x[q] *= inv[q];
#pragma unroll
          for (int m = q + 1; m < 8; ++m) x[m] -= x[q] * l[m][q];
        }
        if (row < c0 + 8) {  // rows of the diagonal 8x8 block: exact diagonal, zero upper part
          const int a = row - c0;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q == a) x[q] = l[q][q];
            if (q > a) x[q] = 0.0;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) myrow[c0 + q] = x[q];
        if (row < NB) {
#pragma unroll
          for (int q = 0; q < 8; ++q) Lt[(c0 + q) * LDT + row] = x[q];
        }
      }
    } else if (cb + 1 < NB / 8 && c0 > 0 && row >= c0 + 8) {
      // helper: next block's entries minus the contribution of columns [0, c0)
      const int n0 = c0 + 8;
      double t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = myrow[n0 + q];
#pragma unroll 4
      for (int k = 0; k < c0; ++k) {
        const double li = myrow[k];
        const double2* lt = reinterpret_cast<const double2*>(Lt + k * LDT + n0);
        const double2 l01 = lt[0], l23 = lt[1], l45 = lt[2], l67 = lt[3];
        t[0] -= li * l01.x; t[1] -= li * l01.y; t[2] -= li * l23.x; t[3] -= li * l23.y;
        t[4] -= li * l45.x; t[5] -= li * l45.y; t[6] -= li * l67.x; t[7] -= li * l67.y;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) myrow[n0 + q] = t[q];
    }
    __syncthreads();
  }

  if (s_last) {   // (s_last was written before the first barrier of the main loop)
    for (int idx = tid; idx < NB * NB; idx += NTHR) {
      const int r = idx >> 6, c = idx & 63;
      if (r < b && c < b) A[(size_t)(i0 + r) * lda + i0 + c] = (c <= r) ? Rs[r * LDR + c] : 0.0;
    }
    if (owner) {
      // `fail` is identical in every owner thread of the CTA
      if (tid == 0 && fail != INT_MAX && fail < b) atomicCAS(info, 0, i0 + fail + 1);
    }
  }
  for (int idx = tid; idx < RB * NB; idx += NTHR) {
    const int r = idx >> 6, c = idx & 63;
    if (r < nrows && c < b) A[(size_t)(row_base + r) * lda + i0 + c] = Rs[(NB + r) * LDR + c];
  }
}

// ---------------------------------------------------------------------------------------
// trsm_rln_kernel: X <- X D^{-1} for rows [r0,r1), columns [j, j+b) of X, with
// D = L[j:j+b, j:j+b] lower triangular (solve X D = C by back substitution over columns).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(pcfg::TRSM_ROWS) trsm_rln_kernel(const double* __restrict__ L,
                                                                   int ldl, double* __restrict__ X,
                                                                   int ldx, int j, int b, int r0,
                                                                   int r1) {
  using namespace pcfg;
  extern __shared__ __align__(16) double sm[];
  double* Dm = sm;                 // [NB][LDT]  natural layout D[k][c]
  double* invd = Dm + NB * LDT;    // [NB]
  double* Xs = invd + NB;          // [TRSM_ROWS][LDR]
  const int tid = threadIdx.x;
  const int row_base = r0 + blockIdx.x * TRSM_ROWS;
  const int nrows = max(0, min(TRSM_ROWS, r1 - row_base));

  for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
    const int k = idx >> 6, c = idx & 63;
    double v = (k == c) ? 1.0 : 0.0;
    if (k < b && c < b) v = (c <= k) ? L[(size_t)(j + k) * ldl + j + c] : 0.0;
    Dm[k * LDT + c] = v;
  }
  for (int idx = tid; idx < TRSM_ROWS * NB; idx += TRSM_ROWS) {
    const int r = idx >> 6, c = idx & 63;
    double v = 0.0;
    if (r < nrows && c < b) v = X[(size_t)(row_base + r) * ldx + j + c];
    Xs[r * LDR + c] = v;
  }
  __syncthreads();
  if (tid < NB) invd[tid] = 1.0 / Dm[tid * LDT + tid];
  __syncthreads();

  double* myrow = Xs + tid * LDR;
#pragma unroll 1
  for (int cb = NB / 8 - 1; cb >= 0; --cb) {
    const int c0 = cb * 8;
    double s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = myrow[c0 + q];
    for (int k = c0 + 8; k < NB; ++k) {
      const double xk = myrow[k];
      const double2* dk = reinterpret_cast<const double2*>(Dm + k * LDT + c0);
      const double2 d01 = dk[0], d23 = dk[1], d45 = dk[2], d67 = dk[3];
      s[0] -= xk * d01.x; s[1] -= xk * d01.y; s[2] -= xk * d23.x; s[3] -= xk * d23.y;
      s[4] -= xk * d45.x; s[5] -= xk * d45.y; s[6] -= xk * d67.x; s[7] -= xk * d67.y;
    }
    double x[8];
#pragma unroll
    for (int q = 7; q >= 0; --q) {
      double v = s[q];
#pragma unroll
      for (int m = q + 1; m < 8; ++m) v -= x[m] * Dm[(c0 + m) * LDT + c0 + q];
      x[q] = v * invd[c0 + q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) myrow[c0 + q] = x[q];
  }
  __syncthreads();
  for (int idx = tid; idx < TRSM_ROWS * NB; idx += TRSM_ROWS) {
    const int r = idx >> 6, c = idx & 63;
    if (r < nrows && c < b) X[(size_t)(row_base + r) * ldx + j + c] = Xs[r * LDR + c];
  }
}

// ---------------------------------------------------------------------------------------
// leaf_rev_kernel: level-2 reverse-mode of the b x b (b <= 64) diagonal block at (j,j):
// Abar[j:j+b, j:j+b] (lower, holding Lbar) <- Abar in Murray's convention.  One CTA.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) leaf_rev_kernel(const double* __restrict__ L, int ldl,
                                                       double* __restrict__ Ab, int ldab, int j0,
                                                       int b) {
  using namespace pcfg;
  extern __shared__ __align__(16) double sm[];
  double* D = sm;               // [NB][LDR]
  double* X = D + NB * LDR;     // [NB][LDR]
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    double d = (r == c) ? 1.0 : 0.0, x = 0.0;
    if (r < b && c < b && c <= r) {
      d = L[(size_t)(j0 + r) * ldl + j0 + c];
      x = Ab[(size_t)(j0 + r) * ldab + j0 + c];
    }
    D[r * LDR + c] = d;
    X[r * LDR + c] = x;
  }
  __syncthreads();

#pragma unroll 1
  for (int j = b - 1; j >= 0; --j) {
    // Phase A (each warp redundantly): dbar = (X_jj - <c, cbar>/d)/d
    double part = 0.0;
    for (int i = j + 1 + lane; i < b; i += 32) part += D[i * LDR + j] * X[i * LDR + j];
    const double tsum = warp_sum(part);
    const double rd = 1.0 / D[j * LDR + j];
    const double dbar = (X[j * LDR + j] - tsum * rd) * rd;
    // Phase B: rbar_k -= dbar r_k + sum_i cbar_i B_ik ; Bbar_ik -= cbar_i r_k  (cbar = X[:,j]/d)
    {
      const int k = tid >> 2, sub = tid & 3;
      double acc = 0.0;
      if (k < j)
        for (int i = j + 1 + sub; i < b; i += 4) acc += X[i * LDR + j] * D[i * LDR + k];
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (k < j && sub == 0) X[j * LDR + k] -= dbar * D[j * LDR + k] + acc * rd;
    }
    if (j > 0) {
      const int cnt = (b - 1 - j) * j;
      for (int idx = tid; idx < cnt; idx += 256) {
        const int i = j + 1 + idx / j, k = idx % j;
        X[i * LDR + k] -= X[i * LDR + j] * rd * D[j * LDR + k];
      }
    }
    __syncthreads();
    // Phase C: finalise column j (not read again)
    if (tid > j && tid < b) X[tid * LDR + j] *= rd;
    if (tid == j) X[j * LDR + j] = 0.5 * dbar;
  }
  __syncthreads();
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    if (r < b && c < b && c <= r) Ab[(size_t)(j0 + r) * ldab + j0 + c] = X[r * LDR + c];
  }
}

// ws[b x b] (ld = b) = Xl + Xl^T with Xl = tril(Ab[j:j+b, j:j+b])
__global__ void symmetrize_block_kernel(const double* __restrict__ Ab, int ldab, int j0, int b,
                                        double* __restrict__ ws) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= b || c >= b) return;
  double v;
  if (r == c) v = 2.0 * Ab[(size_t)(j0 + r) * ldab + j0 + c];
  else if (c < r) v = Ab[(size_t)(j0 + r) * ldab + j0 + c];
  else v = Ab[(size_t)(j0 + c) * ldab + j0 + r];
  ws[(size_t)r * b + c] = v;
}

__global__ void zero_upper_kernel(double* __restrict__ A, int n, int lda) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r < n && c < n && c > r) A[(size_t)r * lda + c] = 0.0;
}

// In place: lower triangle in Murray's convention -> full symmetric gradient G.
__global__ void finalize_sym_kernel(double* __restrict__ Ab, int n, int ld) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= n || c >= n || c >= r) return;
  const double v = 0.5 * Ab[(size_t)r * ld + c];
  Ab[(size_t)r * ld + c] = v;
  Ab[(size_t)c * ld + r] = v;
}

// ---------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------
// The opt-in to > 48 KB of dynamic shared memory is a per-DEVICE attribute of a kernel: one flag per device.
static bool* device_flag(bool (&flags)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return &flags[dev];
}

static int set_smem_attrs() {
  static bool done_dev[64] = {};
  bool& done = *device_flag(done_dev);
  if (done) return GPINV_OK;
  cudaError_t e;
  e = cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pcfg::PANEL_SMEM);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(trsm_rln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pcfg::TRSM_SMEM);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(leaf_rev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pcfg::LEAF_SMEM);
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e));
    return GPINV_ERR_CUDA;
  }
  done = true;
  return GPINV_OK;
}

static int pick_splitk(int M, int N, int K, bool lower) {
  const int tiles = gemm_tile_count(M, N, K, lower, 1);
  const int ktiles = cdiv(K, 16);
  int s = 148 / (tiles > 0 ? tiles : 1);
  s = min(s, ktiles / 8);   // keep >= 8 k-tiles (128 columns of K) per split
  return max(s, 1);
}

static int trtri_f64(const double* L, int ldl, double* W, double* T, int n, cudaStream_t st, int stop = 0x7fffffff,
                     int ws_slot = 0, int ldw = 0, int top_rank = -1, int top_G = 0);

// Side stream + events for the look-ahead (one set per device, created on first use).
struct LookAhead {
  cudaStream_t side = nullptr;
  cudaStream_t inv = nullptr;                  // lowest priority: the inversion interleaved with the factorisation
  cudaEvent_t evP = nullptr, evB = nullptr;
  cudaEvent_t evF = nullptr, evW = nullptr;    // forward factor ready / explicit inverse ready (see trtri_side_f64)
  cudaEvent_t evI = nullptr;                   // outer panel factored (interleaved inversion)
  cudaEvent_t evIn = nullptr, evMid = nullptr, evU = nullptr;   // early sampling product (side_early_*)
  bool ok = false;
};
static LookAhead* lookahead_for_current_device() {
  static LookAhead table[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  LookAhead& la = table[dev];
  if (!la.ok) {
    if (cudaStreamCreateWithFlags(&la.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    int prio_lo = 0, prio_hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&la.inv, cudaStreamNonBlocking, prio_lo) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evI, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evIn, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evMid, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evU, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evP, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evB, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evF, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&la.evW, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    la.ok = true;
  }
  return &la;
}

// Two-level right-looking Cholesky with look-ahead: after outer panel o is factored, the
// trailing update is split into (a) the columns of the NEXT outer panel, done on the caller's
// stream so panel o+1 can start at once, and (b) the rest of the trailing matrix, done on a side
// stream concurrently with (a) and with the factorisation of panel o+1.
// strict upper part of every NBO x NBO diagonal block <- 0
__global__ void zero_upper_diag_blocks_kernel(double* __restrict__ A, int n, int lda, int nbo) {
  const int base = blockIdx.z * nbo;
  const int c = base + blockIdx.x * blockDim.x + threadIdx.x;
  const int r = base + blockIdx.y * blockDim.y + threadIdx.y;
  if (r < n && c < n && c < base + nbo && r < base + nbo && c > r) A[(size_t)r * lda + c] = 0.0;
}

int potrf_ex_f64(double* A, int n, int lda, int* info, int upper_is_zero, cudaStream_t st, double* Winv = nullptr,
                 double* Tinv = nullptr, int* inv_done_out = nullptr);

int potrf_f64(double* A, int n, int lda, int* info, cudaStream_t st) {
  return potrf_ex_f64(A, n, lda, info, 0, st);
}

// GEMM with a CTA cap (runs beside the factorisation's kernel chain), split-K workspace half `ws_slot`
static int gemm_cap(int la_, int lb_, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
                    int ldb, double* out, int ldo, int kmode, int cap, int ws_slot, cudaStream_t st) {
  GemmParams p = {};
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.M = M; p.N = N; p.K = K;
  p.alpha = alpha; p.beta = 0.0; p.Cin = nullptr; p.ldcin = 0; p.out = out; p.ldo = ldo;
  p.kmode = kmode; p.splitk = 1; p.allow_split = 1; p.max_ctas = cap; p.ws_slot = ws_slot;
  return gemm_launch(la_, lb_, EF_BETA, p, st);
}

// Row block [r0, r1) of W = L^-1 (W, T: n x n, ld n; rows < r0 of W done; strict upper part of W zero):
//   W_II = L_II^-1;   W[I, :r0] = -W_II (L[I, :r0] W[:r0, :r0]).
// Needs only L[:r1, :r1], i.e. it can run as soon as the outer panel that ends at column r1 is factored.
static int trtri_row_block(const double* L, int ldl, double* W, double* T, int n, int r0, int r1, int cap,
                           cudaStream_t st) {
  const int bw = r1 - r0;
  int rc = trtri_f64(L + (size_t)r0 * ldl + r0, ldl, W + (size_t)r0 * n + r0, T + (size_t)r0 * n + r0, bw, st,
                     0x7fffffff, 1, n);
  if (rc || r0 == 0) return rc;
  rc = gemm_cap(LK, LMN, bw, r0, r0, 1.0, L + (size_t)r0 * ldl, ldl, W, n, T + (size_t)r0 * n, n, 2, cap, 1, st);
  if (rc) return rc;                                        // T = L[I, :r0] W[:r0, :r0]   (W lower: k >= column)
  return gemm_cap(LK, LMN, bw, r0, bw, -1.0, W + (size_t)r0 * n + r0, n, T + (size_t)r0 * n, n, W + (size_t)r0 * n, n,
                  3, cap, 1, st);                           // W[I, :r0] = -W_II T         (W_II lower: k <= row)
}

// upper_is_zero: the caller built only the lower triangle (strict upper part = 0).  The
// factorisation never reads above the diagonal and only dirties the upper parts of the NBO-wide
// diagonal blocks, so the final clean-up touches those instead of the whole matrix.
// Winv / Tinv (optional, n x n each; EXPERIMENT, see below): W = L^-1 built INSIDE the factorisation, row block
// by row block, on a third, lowest-priority stream.  Row block I needs the leading (I+1) outer panels of L only,
// its cost grows like I^2 while the trailing update shrinks like (N-I)^2, so the idea was that the inversion
// fills the SMs the late panels leave idle instead of competing with the sampling stage afterwards.
// *inv_done_out = 1 when the inversion was enqueued here (the caller joins evW through trtri_side_join_f64).
int potrf_ex_f64(double* A, int n, int lda, int* info, int upper_is_zero, cudaStream_t st, double* Winv,
                 double* Tinv, int* inv_done_out) {
  using namespace pcfg;
  int rc = set_smem_attrs();
  if (rc) return rc;
  if (inv_done_out) *inv_done_out = 0;
  GPINV_TRY(cudaMemsetAsync(info, 0, 2 * sizeof(int), st));
  static const bool no_la = (getenv("GPINV_NO_LOOKAHEAD") != nullptr);
  LookAhead* la = (n > 2 * NBO && !no_la) ? lookahead_for_current_device() : nullptr;
  bool side_pending = false;
  // Measured at n = 4096 (tests/gpu_cholinv_time.py, graph replay): factorisation 2.13 ms; factorisation + inversion
  // behind it (recursive doubling, whole machine) 3.10 ms; interleaved 3.25 ms at best (cap 148, from panel 0; 4.50 ms
  // with cap 64 from panel 6): the row-block form has 256-row GEMMs with few, long tiles, its heavy row blocks only
  // become available at the very end, and the diagonal-block inversions are latency chains.  So it is OFF by default
  // (GPINV_INTERLEAVED_INVERSE=1 to enable).
  static const bool inter = (getenv("GPINV_INTERLEAVED_INVERSE") != nullptr);
  static const int inv_cap = getenv("GPINV_INV_CAP") ? atoi(getenv("GPINV_INV_CAP")) : 148;
  static const int inv_from = getenv("GPINV_INV_FROM") ? atoi(getenv("GPINV_INV_FROM")) : 0;
  const bool inv_on = la && Winv && Tinv && inter && lda == n;
  int inv_rows = 0;                    // row blocks [0, inv_rows) of W are enqueued
  if (inv_on) {
    GPINV_TRY(cudaEventRecord(la->evF, st));
    GPINV_TRY(cudaStreamWaitEvent(la->inv, la->evF, 0));
    GPINV_TRY(cudaMemsetAsync(Winv, 0, (size_t)n * n * sizeof(double), la->inv));
  }
  // "middle" of the factorisation for work that wants to hide inside it (sample_u_early_f64): outer panel
  // 5/8 of the way (panel 10 of 16 at n = 4096; measured: start at panel 6 -> 79.2, 8 -> 80.5, 10 -> 81.3,
  // 12 -> 80.1 evaluations/s, against 78.7 without)
  static const int mid_env = getenv("GPINV_POTRF_MID") ? atoi(getenv("GPINV_POTRF_MID")) : -1;
  const int last_panel = (n - 1) / NBO;
  const int mid_panel = min(mid_env >= 0 ? mid_env : ((last_panel + 1) * 5) / 8, last_panel);
  for (int o = 0; o < n; o += NBO) {
    const int w = min(NBO, n - o);
    if (la && o / NBO == mid_panel)
      GPINV_TRY(cudaEventRecord(la->evMid, st));     // from here on the trailing updates no longer fill the machine
    // Inside an outer panel the 64-column sub-panels are brought up to date either right-looking (after
    // sub-panel j: update ALL remaining columns of the outer panel, K = 64) or left-looking (before
    // sub-panel j: update ITS 64 columns with everything finished in this outer panel, K = 64 j).  Both are
    // latency-bound launches on the critical chain; the left-looking ones are narrower (N = 64).
    static const bool inner_left = (getenv("GPINV_POTRF_INNER_LEFT") != nullptr);
    for (int i = o; i < o + w; i += NB) {
      const int b = min(NB, o + w - i);
      const int below = n - i - b;
      if (inner_left && i > o) {
        double* P = A + (size_t)i * lda + o;
        rc = gemm_simple(LK, LK, n - i, b, i - o, -1.0, P, lda, P, lda, 1.0, A + (size_t)i * lda + i, lda, false,
                         0, 1, st);
        if (rc) return rc;
      }
      panel_kernel<<<max(1, cdiv(below, RB)), 2 * (NB + RB), PANEL_SMEM, st>>>(A, lda, n, i, b, info);
      GPINV_CHECK_LAUNCH();
      if (!inner_left && i + b < o + w) {
        double* P = A + (size_t)(i + b) * lda + i;
        rc = gemm_simple(LK, LK, n - (i + b), o + w - (i + b), b, -1.0, P, lda, P, lda, 1.0,
                         A + (size_t)(i + b) * lda + (i + b), lda, false, 0, 1, st);
        if (rc) return rc;
      }
    }
    const int t0 = o + w;   // first trailing row / column
    if (inv_on) {
      const int pi = o / NBO;
      if (pi >= inv_from || t0 >= n) {
        GPINV_TRY(cudaEventRecord(la->evI, st));            // outer panels 0 .. pi are final
        GPINV_TRY(cudaStreamWaitEvent(la->inv, la->evI, 0));
        for (; inv_rows <= pi; ++inv_rows) {
          const int r0 = inv_rows * NBO, r1 = min(n, r0 + NBO);
          rc = trtri_row_block(A, lda, Winv, Tinv, n, r0, r1, inv_cap, la->inv);
          if (rc) return rc;
        }
      }
    }
    if (t0 >= n) break;
    double* P = A + (size_t)t0 * lda + o;
    if (!la) {
      rc = gemm_simple(LK, LK, n - t0, n - t0, w, -1.0, P, lda, P, lda, 1.0,
                       A + (size_t)t0 * lda + t0, lda, true, 0, 1, st);
      if (rc) return rc;
      continue;
    }
    const int nw = min(NBO, n - t0);   // width of the next outer panel
    const int t1 = t0 + nw;
    GPINV_TRY(cudaEventRecord(la->evP, st));
    GPINV_TRY(cudaStreamWaitEvent(la->side, la->evP, 0));
    if (side_pending) GPINV_TRY(cudaStreamWaitEvent(st, la->evB, 0));
    // (a) next panel's columns, all trailing rows (the upper part of its diagonal block is junk)
    rc = gemm_simple(LK, LK, n - t0, nw, w, -1.0, P, lda, P, lda, 1.0, A + (size_t)t0 * lda + t0,
                     lda, false, 0, 1, st);
    if (rc) return rc;
    side_pending = false;
    if (t1 < n) {  // (b) the rest, lower triangle, on the side stream
      double* P1 = A + (size_t)t1 * lda + o;
      // (b) runs beside the next panel's kernel chain.  A full grid would own every SM for one
      // CTA lifetime (~40 us at K = 256; the 128x128 tile's 65 K registers exclude co-residency)
      // and the chain's next launch would sit waiting for that long, so the grid is capped (the
      // persistent GEMM variant walks the tile list) and ~40 SMs stay free for the chain.
      static const int cap_env = getenv("GPINV_POTRF_B_CAP") ? atoi(getenv("GPINV_POTRF_B_CAP")) : -1;
      const int T = cdiv(n - t1, 128);
      const int tiles = T * (T + 1) / 2;
      (void)tiles;
      int cap = 108;   // measured optimum at n = 4096 (2.68 ms vs 3.01 ms uncapped; 80: 2.85, 132: 2.81)
      if (cap_env >= 0) cap = cap_env;
      GemmParams gp = {};
      gp.A = P1; gp.lda = lda; gp.B = P1; gp.ldb = lda; gp.M = n - t1; gp.N = n - t1; gp.K = w;
      gp.alpha = -1.0; gp.beta = 1.0; gp.Cin = A + (size_t)t1 * lda + t1; gp.ldcin = lda;
      gp.out = A + (size_t)t1 * lda + t1; gp.ldo = lda; gp.splitk = 1; gp.max_ctas = cap;
      gp.force_cfg = cap > 0 ? 1 : 0;
      rc = gemm_launch(LK, LK, EF_BETA | EF_LOWER, gp, la->side);
      if (rc) return rc;
      GPINV_TRY(cudaEventRecord(la->evB, la->side));
      side_pending = true;
    }
  }
  if (side_pending) GPINV_TRY(cudaStreamWaitEvent(st, la->evB, 0));
  if (inv_on) {
    GPINV_TRY(cudaEventRecord(la->evW, la->inv));
    if (inv_done_out) *inv_done_out = 1;
  }
  dim3 blk(32, 8);
  if (upper_is_zero) {
    dim3 grd(cdiv(NBO, 32), cdiv(NBO, 8), cdiv(n, NBO));
    zero_upper_diag_blocks_kernel<<<grd, blk, 0, st>>>(A, n, lda, NBO);
  } else {
    dim3 grd(cdiv(n, 32), cdiv(n, 8));
    zero_upper_kernel<<<grd, blk, 0, st>>>(A, n, lda);
  }
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// X[r0:r1, j:k] <- X[r0:r1, j:k] D^{-1}, D = L[j:k, j:k]; blocked over 64-wide leaves.
static int trsm_rln_blocked(const double* L, int ldl, double* X, int ldx, int j, int k, int r0,
                            int r1, cudaStream_t st) {
  using namespace pcfg;
  if (r1 <= r0) return GPINV_OK;
  int jb = k;
  while (jb > j) {
    const int ja = max(j, jb - NB);
    trsm_rln_kernel<<<cdiv(r1 - r0, TRSM_ROWS), TRSM_ROWS, TRSM_SMEM, st>>>(L, ldl, X, ldx, ja,
                                                                            jb - ja, r0, r1);
    GPINV_CHECK_LAUNCH();
    if (ja > j) {
      int rc = gemm_simple(LK, LMN, r1 - r0, ja - j, jb - ja, -1.0, X + (size_t)r0 * ldx + ja, ldx,
                           L + (size_t)ja * ldl + j, ldl, 1.0, X + (size_t)r0 * ldx + j, ldx, false,
                           0, 1, st);
      if (rc) return rc;
    }
    jb = ja;
  }
  return GPINV_OK;
}

// Reverse sweep restricted to the diagonal block [s,e) of (L, Ab).
static int potrf_bwd_range(const double* L, int ldl, double* Ab, int ldab, int s, int e, int nb,
                           bool two_level, double* ws, cudaStream_t st) {
  using namespace pcfg;
  int rc;
  int k = e;
  while (k > s) {
    const int j = max(s, k - nb);
    const int bw = k - j;
    if (k < e) {
      rc = trsm_rln_blocked(L, ldl, Ab, ldab, j, k, k, e, st);   // Cbar <- Cbar D^{-1}
      if (rc) return rc;
      const double* Cb = Ab + (size_t)k * ldab + j;
      if (j > s) {  // Bbar -= Cbar R
        rc = gemm_simple(LK, LMN, e - k, j - s, bw, -1.0, Cb, ldab, L + (size_t)j * ldl + s, ldl,
                         1.0, Ab + (size_t)k * ldab + s, ldab, false, 0, 1, st);
        if (rc) return rc;
      }
      {  // Dbar -= tril(Cbar^T C)
        const int sk = pick_splitk(bw, bw, e - k, true);
        rc = gemm_simple(LMN, LMN, bw, bw, e - k, -1.0, Cb, ldab, L + (size_t)k * ldl + j, ldl, 1.0,
                         Ab + (size_t)j * ldab + j, ldab, true, 0, sk, st);
        if (rc) return rc;
      }
    }
    if (two_level) {
      rc = potrf_bwd_range(L, ldl, Ab, ldab, j, k, NB, false, ws, st);
      if (rc) return rc;
    } else {
      leaf_rev_kernel<<<1, 256, LEAF_SMEM, st>>>(L, ldl, Ab, ldab, j, bw);
      GPINV_CHECK_LAUNCH();
    }
    if (j > s) {
      if (k < e) {  // Rbar -= Cbar^T B
        const int sk = pick_splitk(bw, j - s, e - k, false);
        rc = gemm_simple(LMN, LMN, bw, j - s, e - k, -1.0, Ab + (size_t)k * ldab + j, ldab,
                         L + (size_t)k * ldl + s, ldl, 1.0, Ab + (size_t)j * ldab + s, ldab, false,
                         0, sk, st);
        if (rc) return rc;
      }
      // Rbar -= (Dbar + Dbar^T) R
      dim3 blk(32, 8), grd(cdiv(bw, 32), cdiv(bw, 8));
      symmetrize_block_kernel<<<grd, blk, 0, st>>>(Ab, ldab, j, bw, ws);
      GPINV_CHECK_LAUNCH();
      rc = gemm_simple(LK, LMN, bw, j - s, bw, -1.0, ws, bw, L + (size_t)j * ldl + s, ldl, 1.0,
                       Ab + (size_t)j * ldab + s, ldab, false, 0, 1, st);
      if (rc) return rc;
    }
    k = j;
  }
  return GPINV_OK;
}

size_t potrf_bwd_ws_bytes(int n) {
  (void)n;
  // one outer-level and one inner-level symmetrised block
  return (size_t)(pcfg::NBO * pcfg::NBO + pcfg::NB * pcfg::NB) * sizeof(double);
}

// Ab holds Lbar (lower) on entry; on exit the full symmetric gradient dF/dA.
int potrf_bwd_f64(const double* L, int ldl, double* Ab, int ldab, int n, double* ws,
                  cudaStream_t st) {
  using namespace pcfg;
  int rc = set_smem_attrs();
  if (rc) return rc;
  // The inner level needs its own symmetrise buffer: it runs between the outer level's
  // symmetrise and its consumer only across iterations, never inside one, but keep them apart.
  double* ws_outer = ws;
  double* ws_inner = ws + (size_t)NBO * NBO;
  // outer level uses ws_outer; potrf_bwd_range passes `ws` down, so wrap manually:
  int k = n;
  while (k > 0) {
    const int j = max(0, k - NBO);
    const int bw = k - j;
    if (k < n) {
      rc = trsm_rln_blocked(L, ldl, Ab, ldab, j, k, k, n, st);
      if (rc) return rc;
      const double* Cb = Ab + (size_t)k * ldab + j;
      if (j > 0) {
        rc = gemm_simple(LK, LMN, n - k, j, bw, -1.0, Cb, ldab, L + (size_t)j * ldl, ldl, 1.0,
                         Ab + (size_t)k * ldab, ldab, false, 0, 1, st);
        if (rc) return rc;
      }
      const int sk = pick_splitk(bw, bw, n - k, true);
      rc = gemm_simple(LMN, LMN, bw, bw, n - k, -1.0, Cb, ldab, L + (size_t)k * ldl + j, ldl, 1.0,
                       Ab + (size_t)j * ldab + j, ldab, true, 0, sk, st);
      if (rc) return rc;
    }
    rc = potrf_bwd_range(L, ldl, Ab, ldab, j, k, NB, false, ws_inner, st);
    if (rc) return rc;
    if (j > 0) {
      if (k < n) {
        const int sk = pick_splitk(bw, j, n - k, false);
        rc = gemm_simple(LMN, LMN, bw, j, n - k, -1.0, Ab + (size_t)k * ldab + j, ldab,
                         L + (size_t)k * ldl, ldl, 1.0, Ab + (size_t)j * ldab, ldab, false, 0, sk,
                         st);
        if (rc) return rc;
      }
      dim3 blk(32, 8), grd(cdiv(bw, 32), cdiv(bw, 8));
      symmetrize_block_kernel<<<grd, blk, 0, st>>>(Ab, ldab, j, bw, ws_outer);
      GPINV_CHECK_LAUNCH();
      rc = gemm_simple(LK, LMN, bw, j, bw, -1.0, ws_outer, bw, L + (size_t)j * ldl, ldl, 1.0,
                       Ab + (size_t)j * ldab, ldab, false, 0, 1, st);
      if (rc) return rc;
    }
    k = j;
  }
  dim3 blk(32, 8), grd(cdiv(n, 32), cdiv(n, 8));
  finalize_sym_kernel<<<grd, blk, 0, st>>>(Ab, n, ldab);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// ---------------------------------------------------------------------------------------
// Inverse-based reverse sweep (default):   G = sym( L^{-T} Phi(L^T Lbar) L^{-1} ).
// W = L^{-1} is built level by level: all 64x64 diagonal blocks are inverted by one launch
// of the substitution kernel, then at level s every off-diagonal block
// W21 = -W22 (L21 W11) is two *batched* DMMA GEMMs.  No 64-step chain of small kernels:
// ~20 launches instead of ~640, 5n^3/3 flops instead of 2n^3/3, all at GEMM speed.
// ---------------------------------------------------------------------------------------
__global__ void set_identity_blocks_kernel(double* __restrict__ W, int n, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) W[(size_t)i * ld + i] = 1.0;
}

// one CTA per 64-wide diagonal block: X_bb <- X_bb L_bb^{-1} (X_bb = I on entry)
__global__ void __launch_bounds__(pcfg::NB) trtri_leaf_kernel(const double* __restrict__ L, int ldl,
                                                              double* __restrict__ W, int ldw,
                                                              int n) {
  using namespace pcfg;
  extern __shared__ __align__(16) double sm[];
  double* Dm = sm;               // [NB][LDT]
  double* invd = Dm + NB * LDT;  // [NB]
  double* Xs = invd + NB;        // [NB][LDR]
  const int tid = threadIdx.x;
  const int j = blockIdx.x * NB;
  const int b = min(NB, n - j);
  for (int idx = tid; idx < NB * NB; idx += NB) {
    const int k = idx >> 6, c = idx & 63;
    double v = (k == c) ? 1.0 : 0.0;
    if (k < b && c < b) v = (c <= k) ? L[(size_t)(j + k) * ldl + j + c] : 0.0;
    Dm[k * LDT + c] = v;
    Xs[k * LDR + c] = (k == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  invd[tid] = 1.0 / Dm[tid * LDT + tid];
  __syncthreads();
  double* myrow = Xs + tid * LDR;
#pragma unroll 1
  for (int cb = NB / 8 - 1; cb >= 0; --cb) {
    const int c0 = cb * 8;
    if (c0 > tid) continue;  // row tid of a lower-triangular inverse is zero right of column tid
    double s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = myrow[c0 + q];
    for (int k = c0 + 8; k <= tid; ++k) {
      const double xk = myrow[k];
      const double2* dk = reinterpret_cast<const double2*>(Dm + k * LDT + c0);
      const double2 d01 = dk[0], d23 = dk[1], d45 = dk[2], d67 = dk[3];
      s[0] -= xk * d01.x; s[1] -= xk * d01.y; s[2] -= xk * d23.x; s[3] -= xk * d23.y;
      s[4] -= xk * d45.x; s[5] -= xk * d45.y; s[6] -= xk * d67.x; s[7] -= xk * d67.y;
    }
    double x[8];
#pragma unroll
    for (int q = 7; q >= 0; --q) {
      double v = s[q];
#pragma unroll
      for (int m = q + 1; m < 8; ++m) v -= x[m] * Dm[(c0 + m) * LDT + c0 + q];
      x[q] = v * invd[c0 + q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) myrow[c0 + q] = x[q];
  }
  __syncthreads();
  for (int idx = tid; idx < NB * NB; idx += NB) {
    const int r = idx >> 6, c = idx & 63;
    if (r < b && c < b) W[(size_t)(j + r) * ldw + j + c] = (c <= r) ? Xs[r * LDR + c] : 0.0;
  }
}

__global__ void scale_diag_kernel(double* __restrict__ P, int n, int ld, double f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) P[(size_t)i * ld + i] *= f;
}

__global__ void mirror_lower_kernel(double* __restrict__ A, int n, int ld) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r < n && c < r) A[(size_t)c * ld + r] = A[(size_t)r * ld + c];
}

// W (n x n, ld = n, zero on entry) <- L^{-1};  T: n x n scratch.
// With `stop` < n only the diagonal blocks of size `stop` (a power-of-two multiple of NB, blocks
// aligned at multiples of `stop`) are inverted; the rest of W stays zero.
// top_G > 1 (sample sharding over top_G GPUs, n = 2 s_top): the LAST merge level -- W21 = -W22 L21 W11 with
// s_top = n / 2, three quarters of the inversion's flops -- is computed for this rank's column slices of W21
// only (2 top_G slices of width s_top / 2 top_G, rank r owns slices r and 2 top_G - 1 - r: the triangular k
// range of the first product makes early slices dearer, pairing evens it out); the caller all-gathers the slices.
static int trtri_f64(const double* L, int ldl, double* W, double* T, int n, cudaStream_t st, int stop, int ws_slot,
                     int ldw, int top_rank, int top_G) {
  if (ldw <= 0) ldw = n;        // W and T are n x n with leading dimension ldw
  using namespace pcfg;
  static bool attr_dev[64] = {};
  bool& attr = *device_flag(attr_dev);
  const int smem = (NB * LDT + NB + NB * LDR) * 8;
  if (!attr) {
    GPINV_TRY(cudaFuncSetAttribute(trtri_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem));
    attr = true;
  }
  trtri_leaf_kernel<<<cdiv(n, NB), NB, smem, st>>>(L, ldl, W, ldw, n);
  GPINV_CHECK_LAUNCH();
  int rc;
  for (int s = NB; s < n && s < stop; s *= 2) {
    // pairs (i, m = i+s, e = min(i+2s, n)); full pairs are batched, a ragged last one is separate
    int nfull = 0;
    while ((long long)(nfull + 1) * 2 * s <= n) ++nfull;
    const long long strL = (long long)2 * s * ldl + 2 * s;
    const long long strW = (long long)2 * s * ldw + 2 * s;
    if (top_G > 1 && 2 * s == n && s % (4 * top_G) == 0) {
      const int wsl = s / (2 * top_G);
      for (int t = 0; t < 2; ++t) {
        const int a = (t == 0 ? top_rank : 2 * top_G - 1 - top_rank) * wsl;      // first column of the slice
        // T21[:, a:a+wsl] = L21[:, a:] W11[a:, a:a+wsl]      (rows < a of that column slice of W11 are zero)
        rc = gemm_batched(LK, LMN, s, wsl, s - a, 1.0, L + (size_t)s * ldl + a, ldl, 0, W + (size_t)a * ldw + a, ldw, 0,
                          0.0, T + (size_t)s * ldw + a, ldw, 0, 1, false, 2, st, 1, ws_slot);
        if (rc) return rc;
        // W21[:, a:a+wsl] = -W22 T21[:, a:a+wsl]
        rc = gemm_batched(LK, LMN, s, wsl, s, -1.0, W + (size_t)s * ldw + s, ldw, 0, T + (size_t)s * ldw + a, ldw, 0,
                          0.0, W + (size_t)s * ldw + a, ldw, 0, 1, false, 3, st, 1, ws_slot);
        if (rc) return rc;
      }
      continue;
    }
    for (int pass = 0; pass < 2; ++pass) {
      const int i0 = pass == 0 ? 0 : nfull * 2 * s;
      const int batch = pass == 0 ? nfull : 1;
      const int m0 = i0 + s;
      const int rows = pass == 0 ? s : min(i0 + 2 * s, n) - m0;
      if (batch <= 0 || rows <= 0 || m0 >= n) continue;
      // T21 = L21 W11   (W11 lower: B(c,k) = W11[k][c] is zero for k < c)
      rc = gemm_batched(LK, LMN, rows, s, s, 1.0, L + (size_t)m0 * ldl + i0, ldl, strL,
                        W + (size_t)i0 * ldw + i0, ldw, strW, 0.0, T + (size_t)m0 * ldw + i0, ldw, strW,
                        batch, false, 2, st, 1, ws_slot);
      if (rc) return rc;
      // W21 = -W22 T21  (W22 lower: A(r,k) zero for k > r)
      rc = gemm_batched(LK, LMN, rows, s, rows, -1.0, W + (size_t)m0 * ldw + m0, ldw, strW,
                        T + (size_t)m0 * ldw + i0, ldw, strW, 0.0, W + (size_t)m0 * ldw + i0, ldw, strW,
                        batch, false, 3, st, 1, ws_slot);
      if (rc) return rc;
    }
  }
  return GPINV_OK;
}

size_t potrf_bwd_inv_ws_bytes(int n) {
  // W, S1, S2 (n x n each) + S3, S4 (block x block each, block <= 1024 by default)
  return ((size_t)3 * n * n + (size_t)2 * 2048 * 2048) * sizeof(double);
}

// Ab: Lbar (any upper part is ignored) on entry, full symmetric dObj/dA on exit.
int potrf_bwd_inv_f64(const double* L, int ldl, double* Ab, int ldab, int n, double* ws,
                      cudaStream_t st) {
  double* W = ws;
  double* P = W + (size_t)n * n;
  double* T1 = P + (size_t)n * n;
  GPINV_TRY(cudaMemsetAsync(W, 0, (size_t)n * n * sizeof(double), st));
  int rc = trtri_f64(L, ldl, W, P, n, st);
  if (rc) return rc;
  dim3 blk(32, 8), grd(cdiv(n, 32), cdiv(n, 8));
  zero_upper_kernel<<<grd, blk, 0, st>>>(Ab, n, ldab);
  GPINV_CHECK_LAUNCH();
  // P = tril(L^T Lbar), diagonal halved (Phi)
  {
    GemmParams p = {};
    p.A = L; p.lda = ldl; p.B = Ab; p.ldb = ldab; p.M = n; p.N = n; p.K = n;
    p.alpha = 1.0; p.out = P; p.ldo = n; p.kmode = 4; p.splitk = 1; p.allow_split = 1;
    rc = gemm_launch(LMN, LMN, EF_BETA | EF_LOWER, p, st);
    if (rc) return rc;
    scale_diag_kernel<<<cdiv(n, 256), 256, 0, st>>>(P, n, n, 0.5);
    GPINV_CHECK_LAUNCH();
  }
  // T1 = tril(P W)
  {
    GemmParams p = {};
    p.A = P; p.lda = n; p.B = W; p.ldb = n; p.M = n; p.N = n; p.K = n;
    p.alpha = 1.0; p.out = T1; p.ldo = n; p.kmode = 5; p.splitk = 1; p.allow_split = 1;
    rc = gemm_launch(LK, LMN, EF_BETA | EF_LOWER, p, st);
    if (rc) return rc;
  }
  // tril(G) = 1/2 tril(W^T T1) + 1/2 tril(T1^T W)
  {
    GemmParams p = {};
    p.A = W; p.lda = n; p.B = T1; p.ldb = n; p.M = n; p.N = n; p.K = n;
    p.alpha = 0.5; p.out = Ab; p.ldo = ldab; p.kmode = 4; p.splitk = 1; p.allow_split = 1;
    rc = gemm_launch(LMN, LMN, EF_BETA | EF_LOWER, p, st);
    if (rc) return rc;
    p.A = T1; p.B = W; p.beta = 1.0; p.Cin = Ab; p.ldcin = ldab;
    rc = gemm_launch(LMN, LMN, EF_BETA | EF_LOWER, p, st);
    if (rc) return rc;
  }
  mirror_lower_kernel<<<grd, blk, 0, st>>>(Ab, n, ldab);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// ---------------------------------------------------------------------------------------
// Hybrid reverse sweep (default for n > 2*BWD_BLOCK): Murray's blocked sweep over diagonal blocks
// of m = BWD_BLOCK (1024) columns, whose big products run as DMMA GEMMs of dimension >= m, with
//   * the triangular solve  Cbar <- Cbar D^{-1}  replaced by one GEMM with W_bb = L_bb^{-1}
//     (all W_bb are built first, level by level, batched over the blocks), and
//   * the diagonal-block reverse done by the inverse formula above on the m x m block.
// ~0.73 n^3 flops at n = 4m (the sweep's own 2n^3/3 + the block inverses and the inverse-form
// leaves) instead of 5n^3/3, in ~60 launches, none of them smaller than a 1024-wide GEMM except
// the first trtri levels.  Block results are kept in Murray's convention (strict lower part
// doubled) until finalize_sym_kernel.
// ---------------------------------------------------------------------------------------
static int gemm_ws(int la, int lb, unsigned ef, int M, int N, int K, double alpha, const double* A,
                   int lda, const double* B, int ldb, double beta, const double* Cin, int ldcin,
                   double* out, int ldo, int kmode, cudaStream_t st, int ws_slot = 0) {
  GemmParams p = {};
  p.ws_slot = ws_slot;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.M = M; p.N = N; p.K = K;
  p.alpha = alpha; p.beta = beta; p.Cin = Cin; p.ldcin = ldcin; p.out = out; p.ldo = ldo;
  p.kmode = kmode; p.splitk = 1; p.allow_split = 1;
  return gemm_launch(la, lb, ef, p, st);
}

// Block width of the reverse sweep.  Default: ONE block, i.e. the inverse formula applied to the whole
// matrix.  Measured against the longdouble arbiter (oracle/arbiter.py, lengthscale gradient at cfg4,
// n = 2048): whole matrix 4.6e-10; blocks of 64 / 128 (substitution / inverse-form leaves) 1e-8; but
// blocks of 512 / 1024 with inverse-form leaves 1e-7 .. 4e-6 -- the leaf's error (~cond(L_bb)^2 eps of the
// block norm) is fed through Rbar -= sym(Dbar) R and the following block solves -- and refining the
// leaf's triangular solves does not help.  The blocked sweep (0.73 n^3 flops instead of 5 n^3 / 3)
// stays selectable with GPINV_BWD_BLOCK for experiments.
static int bwd_block() {
  static const int m = [] {
    const char* e = getenv("GPINV_BWD_BLOCK");
    if (!e) return 1 << 30;
    int v = atoi(e);
    int p2 = pcfg::NB;
    while (p2 * 2 <= v) p2 *= 2;
    return p2;
  }();
  return m;
}

int potrf_bwd_hybrid_ex_f64(const double* L, int ldl, double* Ab, int ldab, int n, double* ws,
                            int ab_is_tril, int finalize, cudaStream_t st, const double* W_pre = nullptr);

// Reverse of ONE diagonal block by the inverse formula, with both triangular solves refined:
//   G = sym(L^-T Phi(L^T Lbar) L^-1),   W ~ L^-1 explicit.
//   stage 1   T  = P L^-1  :  T0 = tril(P W),  R1 = P - tril(T0 L),  T = T0 + tril(R1 W)
//   stage 2   H  = L^-T T  :  H0 = W^T T,      R2 = T - L^T H0,      H = H0 + W^T R2
//   Dbar <- tril(H) + tril(H^T)  with the diagonal halved  (Murray's convention: strict lower = 2 G)
// Products with an explicit inverse are not backward stable; inside the blocked sweep their error
// (~cond(L_bb) eps relative to the block) feeds the later blocks and the lengthscale reduction, a sum
// that cancels to ~1e-6 of its terms.  Measured against the longdouble arbiter (oracle/arbiter.py)
// at cfg4 (n = 2048, two blocks): unrefined 3.8e-6, substitution-based leaf 1.1e-8, refined
// no better (2e-6 .. 4e-6): the refinement is kept as an experiment only.
// GPINV_LEAF_REFINE = bitmask (1: stage 1, 2: stage 2; default 0 = plain inverse formula).
static int leaf_reverse_refined(const double* Lbb, int ldl, const double* Wbb, int ldw, double* Dbar,
                                int ldab, int bw, double* P, double* T, double* H, double* R2,
                                bool scratch_ok, cudaStream_t st) {
  static const int refine_env = getenv("GPINV_LEAF_REFINE") ? atoi(getenv("GPINV_LEAF_REFINE")) : 0;
  const int refine = scratch_ok ? refine_env : (refine_env & 1);
  int rc;
  rc = gemm_ws(LMN, LMN, EF_BETA | EF_LOWER, bw, bw, bw, 1.0, Lbb, ldl, Dbar, ldab, 0.0, nullptr, 0, P, bw,
               4, st);                                          // P = tril(L^T Lbar)
  if (rc) return rc;
  scale_diag_kernel<<<cdiv(bw, 256), 256, 0, st>>>(P, bw, bw, 0.5);   // Phi
  GPINV_CHECK_LAUNCH();
  if (refine & 2) GPINV_TRY(cudaMemsetAsync(T, 0, (size_t)bw * bw * sizeof(double), st));   // H0 reads all of T
  rc = gemm_ws(LK, LMN, EF_BETA | EF_LOWER, bw, bw, bw, 1.0, P, bw, Wbb, ldw, 0.0, nullptr, 0, T, bw, 5, st);
  if (rc) return rc;                                            // T0 = tril(P W)
  if (refine & 1) {
    rc = gemm_ws(LK, LMN, EF_BETA | EF_LOWER, bw, bw, bw, -1.0, T, bw, Lbb, ldl, 1.0, P, bw, P, bw, 5, st);
    if (rc) return rc;                                          // R1 = P - tril(T0 L)     (in place of P)
    rc = gemm_ws(LK, LMN, EF_BETA | EF_LOWER, bw, bw, bw, 1.0, P, bw, Wbb, ldw, 1.0, T, bw, T, bw, 5, st);
    if (rc) return rc;                                          // T = T0 + tril(R1 W)
  }
  if (!(refine & 2)) {
    rc = gemm_ws(LMN, LMN, EF_BETA | EF_LOWER, bw, bw, bw, 1.0, Wbb, ldw, T, bw, 0.0, nullptr, 0, Dbar, ldab, 4, st);
    if (rc) return rc;                                          // tril(W^T T)
  } else {
    rc = gemm_ws(LMN, LMN, EF_BETA, bw, bw, bw, 1.0, Wbb, ldw, T, bw, 0.0, nullptr, 0, H, bw, 4, st);
    if (rc) return rc;                                          // H0 = W^T T  (full; k >= row)
    rc = gemm_ws(LMN, LMN, EF_BETA, bw, bw, bw, -1.0, Lbb, ldl, H, bw, 1.0, T, bw, R2, bw, 4, st);
    if (rc) return rc;                                          // R2 = T - L^T H0  (full)
    rc = gemm_ws(LMN, LMN, EF_BETA | EF_LOWER, bw, bw, bw, 1.0, Wbb, ldw, R2, bw, 1.0, H, bw, Dbar, ldab, 4, st);
    if (rc) return rc;                                          // tril(H0 + W^T R2) = tril(H)
    rc = gemm_ws(LMN, LMN, EF_BETA | EF_LOWER, bw, bw, bw, 1.0, R2, bw, Wbb, ldw, 1.0, Dbar, ldab, Dbar, ldab, 2, st);
    if (rc) return rc;                                          // + tril(R2^T W)      (k >= column)
  }
  rc = gemm_ws(LMN, LMN, EF_BETA | EF_LOWER, bw, bw, bw, 1.0, T, bw, Wbb, ldw, 1.0, Dbar, ldab, Dbar, ldab, 4, st);
  if (rc) return rc;                                            // + tril(T^T W) = tril(H0^T)
  scale_diag_kernel<<<cdiv(bw, 256), 256, 0, st>>>(Dbar, bw, ldab, 0.5);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int potrf_bwd_hybrid_f64(const double* L, int ldl, double* Ab, int ldab, int n, double* ws,
                         cudaStream_t st) {
  return potrf_bwd_hybrid_ex_f64(L, ldl, Ab, ldab, n, ws, 0, 1, st);
}

// ab_is_tril: Ab already has a zero strict upper part.  finalize = 0 leaves the result as a lower
// triangle in Murray's convention (strict lower part = 2 x the symmetric gradient entry).
int potrf_bwd_hybrid_ex_f64(const double* L, int ldl, double* Ab, int ldab, int n, double* ws,
                            int ab_is_tril, int finalize, cudaStream_t st, const double* W_pre) {
  int m = bwd_block();
  if (m > n) {                          // one block: round n up to the leaf granularity
    m = pcfg::NB;
    while (m < n) m *= 2;
  }
  double* Wown = ws;                    // n x n, only the m x m diagonal blocks are non-zero
  const double* W = W_pre ? W_pre : Wown;   // W_pre: whole inverse already built (trtri_side_f64)
  double* S1 = Wown + (size_t)n * n;    // scratch: Cbar W_bb  /  P
  double* S2 = S1 + (size_t)n * n;      // scratch: T1  /  symmetrised block
  double* S3 = S2 + (size_t)n * n;      // scratch of the refined diagonal-block reverse (m x m each)
  double* S4 = S3 + (size_t)m * m;
  int rc = GPINV_OK;
  if (n > pcfg::NB && !(W_pre && m >= n)) {     // (a single 64-wide block is reversed by substitution)
    GPINV_TRY(cudaMemsetAsync(Wown, 0, (size_t)n * n * sizeof(double), st));
    rc = trtri_f64(L, ldl, Wown, S1, n, st, m);
    if (rc) return rc;
    W = Wown;
  }
  dim3 blk(32, 8), grd(cdiv(n, 32), cdiv(n, 8));
  if (!ab_is_tril) {
    zero_upper_kernel<<<grd, blk, 0, st>>>(Ab, n, ldab);
    GPINV_CHECK_LAUNCH();
  }
  const int nblocks = cdiv(n, m);
  static const bool no_fork = (getenv("GPINV_BWD_NO_FORK") != nullptr);
  LookAhead* la2 = no_fork ? nullptr : lookahead_for_current_device();
  bool forked = false;
  for (int b = nblocks - 1; b >= 0; --b) {
    const int s = b * m, e = min(n, s + m), bw = e - s, r = n - e;
    const double* Wbb = W + (size_t)s * n + s;
    const double* Lbb = L + (size_t)s * ldl + s;
    double* Dbar = Ab + (size_t)s * ldab + s;
    double* Cb = Ab + (size_t)e * ldab + s;            // rows below the block
    if (r > 0) {
      // Cbar <- Cbar D^{-1} by the explicit block inverse PLUS one step of iterative refinement:
      //   X0 = Cbar W_bb,   R = Cbar - X0 D,   X1 = X0 + R W_bb      (W_bb, D lower: k >= column)
      // Multiplying by an explicit inverse is not backward stable (residual ~ cond(D) eps), and the
      // lengthscale gradient -- a sum over Kbar that cancels to ~1e-6 of its terms -- sees it: against
      // the longdouble arbiter (oracle/arbiter.py) the unrefined sweep was off by 7.6e-6 at n = 2048
      // where the substitution-based sweep and the refined one are at ~1e-8.  Costs two more
      // r x m x m triangular GEMMs per block column; the copy-back of X disappears into the last one.
      static const bool trsm_subst0 = (getenv("GPINV_BWD_TRSM_SUBST") != nullptr);
      if (!trsm_subst0) {
        rc = gemm_ws(LK, LMN, EF_BETA, r, bw, bw, 1.0, Cb, ldab, Wbb, n, 0.0, nullptr, 0, S1, bw, 2, st);
        if (rc) return rc;
      }
      static const bool no_refine = (getenv("GPINV_BWD_NO_REFINE") != nullptr);
      static const bool trsm_subst = (getenv("GPINV_BWD_TRSM_SUBST") != nullptr);   // experiment
      if (trsm_subst) {
        rc = trsm_rln_blocked(L, ldl, Ab, ldab, s, e, e, n, st);
        if (rc) return rc;
      } else if (no_refine) {
        GPINV_TRY(cudaMemcpy2DAsync(Cb, (size_t)ldab * sizeof(double), S1, (size_t)bw * sizeof(double),
                                    (size_t)bw * sizeof(double), r, cudaMemcpyDeviceToDevice, st));
      } else {
        rc = gemm_ws(LK, LMN, EF_BETA, r, bw, bw, -1.0, S1, bw, Lbb, ldl, 1.0, Cb, ldab, S2, bw, 2, st);
        if (rc) return rc;
        rc = gemm_ws(LK, LMN, EF_BETA, r, bw, bw, 1.0, S2, bw, Wbb, n, 1.0, S1, bw, Cb, ldab, 2, st);
        if (rc) return rc;
      }
      // Two independent chains follow from Cbar: {Bbar -= Cbar R, Rbar -= Cbar^T B} (big GEMMs) and
      // {Dbar -= tril(Cbar^T C), diagonal-block reverse} (small, latency-bound GEMMs).  The first
      // runs on the side stream (second half of the split-K workspace) beside the second.
      forked = false;
      if (s > 0 && la2) {
        GPINV_TRY(cudaEventRecord(la2->evP, st));
        GPINV_TRY(cudaStreamWaitEvent(la2->side, la2->evP, 0));
        forked = true;
      }
      if (s > 0) {
        cudaStream_t s2 = forked ? la2->side : st;
        const int slot = forked ? 1 : 0;
        // Bbar -= Cbar R
        rc = gemm_ws(LK, LMN, EF_BETA, r, s, bw, -1.0, Cb, ldab, L + (size_t)s * ldl, ldl, 1.0,
                     Ab + (size_t)e * ldab, ldab, Ab + (size_t)e * ldab, ldab, 0, s2, slot);
        if (rc) return rc;
        // Rbar -= Cbar^T B
        double* Rb2 = Ab + (size_t)s * ldab;
        rc = gemm_ws(LMN, LMN, EF_BETA, bw, s, r, -1.0, Cb, ldab, L + (size_t)e * ldl, ldl, 1.0, Rb2, ldab,
                     Rb2, ldab, 0, s2, slot);
        if (rc) return rc;
        if (forked) GPINV_TRY(cudaEventRecord(la2->evB, la2->side));
      }
      // Dbar -= tril(Cbar^T C)
      rc = gemm_ws(LMN, LMN, EF_BETA | EF_LOWER, bw, bw, r, -1.0, Cb, ldab, L + (size_t)e * ldl + s, ldl,
                   1.0, Dbar, ldab, Dbar, ldab, 0, st);
      if (rc) return rc;
    }
    // diagonal block: Dbar (holding Lbar_bb) -> Murray-convention tril of the block gradient
    static const bool leaf_murray = (getenv("GPINV_BWD_LEAF_MURRAY") != nullptr);   // experiment
    if (leaf_murray) {
      rc = set_smem_attrs();
      if (rc) return rc;
      rc = potrf_bwd_range(L, ldl, Ab, ldab, s, e, pcfg::NBO, true, S2, st);
      if (rc) return rc;
    } else if (bw <= pcfg::NB) {
      // a single 64-wide block: the level-2 reverse in shared memory (substitution, one launch)
      rc = set_smem_attrs();
      if (rc) return rc;
      leaf_rev_kernel<<<1, 256, pcfg::LEAF_SMEM, st>>>(L, ldl, Ab, ldab, s, bw);
      GPINV_CHECK_LAUNCH();
    } else {
      rc = leaf_reverse_refined(Lbb, ldl, Wbb, n, Dbar, ldab, bw, S1, S2, S3, S4, bw <= 2048, st);
      if (rc) return rc;
    }
    if (s > 0) {
      double* Rb = Ab + (size_t)s * ldab;               // Rbar: block rows, columns [0, s)
      if (forked) GPINV_TRY(cudaStreamWaitEvent(st, la2->evB, 0));
      forked = false;
      // Rbar -= (Dbar + Dbar^T) R
      dim3 g2(cdiv(bw, 32), cdiv(bw, 8));
      symmetrize_block_kernel<<<g2, blk, 0, st>>>(Ab, ldab, s, bw, S2);
      GPINV_CHECK_LAUNCH();
      rc = gemm_ws(LK, LMN, EF_BETA, bw, s, bw, -1.0, S2, bw, L + (size_t)s * ldl, ldl, 1.0, Rb, ldab, Rb,
                   ldab, 0, st);
      if (rc) return rc;
    }
  }
  if (finalize) {
    finalize_sym_kernel<<<grd, blk, 0, st>>>(Ab, n, ldab);
    GPINV_CHECK_LAUNCH();
  }
  return GPINV_OK;
}

// dst (ld = ldd) <- tril(src)
__global__ void tril_copy_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst,
                                 int ldd, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r < n && c < n) dst[(size_t)r * ldd + c] = (c <= r) ? src[(size_t)r * lds + c] : 0.0;
}

int tril_copy_f64(const double* src, int lds, double* dst, int ldd, int n, cudaStream_t st) {
  dim3 blk(32, 8), grd(cdiv(n, 32), cdiv(n, 8));
  tril_copy_kernel<<<grd, blk, 0, st>>>(src, lds, dst, ldd, n);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int potrf_bwd_hybrid_block() { return bwd_block(); }

// W <- L^-1 (n x n, ld n) on the library's side stream, started as soon as everything enqueued on `st`
// so far (the factorisation) is done.  The inverse only depends on L, so it overlaps the sampling /
// likelihood stages that follow on `st`; trtri_side_join_f64 makes a stream wait for it (the reverse
// sweep calls it before its first use of W).  T: n x n scratch.  Split-K workspace half 1.
int trtri_side_shard_f64(const double* L, int ldl, double* W, double* T, int n, int rank, int G, cudaStream_t st);

int trtri_side_f64(const double* L, int ldl, double* W, double* T, int n, cudaStream_t st) {
  return trtri_side_shard_f64(L, ldl, W, T, n, 0, 1, st);
}

// rank / G: see trtri_f64 (G = 1: the whole inverse)
int trtri_side_shard_f64(const double* L, int ldl, double* W, double* T, int n, int rank, int G, cudaStream_t st) {
  if (G < 1 || rank < 0 || rank >= G) return GPINV_ERR_ARG;
  LookAhead* la = lookahead_for_current_device();
  if (!la) return GPINV_ERR_CUDA;
  GPINV_TRY(cudaEventRecord(la->evF, st));
  GPINV_TRY(cudaStreamWaitEvent(la->side, la->evF, 0));
  GPINV_TRY(cudaMemsetAsync(W, 0, (size_t)n * n * sizeof(double), la->side));
  int rc = trtri_f64(L, ldl, W, T, n, la->side, 0x7fffffff, 1, 0, rank, G);
  if (rc) return rc;
  GPINV_TRY(cudaEventRecord(la->evW, la->side));
  return GPINV_OK;
}

// ---- early sampling product (sample.cu: sample_u_early_f64) ----
// side_mark_inputs: the inputs of the product exist at this point of `st`.  side_early_begin: the low-priority
// stream waits for that point and for the middle of the most recent factorisation, and is handed out;
// side_early_end / side_early_join: completion event, and a stream waiting for it.
int side_mark_inputs(cudaStream_t st) {
  LookAhead* la = lookahead_for_current_device();
  if (!la) return GPINV_ERR_CUDA;
  GPINV_TRY(cudaEventRecord(la->evIn, st));
  return GPINV_OK;
}
int side_early_begin(cudaStream_t* out) {
  LookAhead* la = lookahead_for_current_device();
  if (!la) return GPINV_ERR_CUDA;
  GPINV_TRY(cudaStreamWaitEvent(la->inv, la->evIn, 0));
  GPINV_TRY(cudaStreamWaitEvent(la->inv, la->evMid, 0));
  *out = la->inv;
  return GPINV_OK;
}
int side_early_end() {
  LookAhead* la = lookahead_for_current_device();
  if (!la) return GPINV_ERR_CUDA;
  GPINV_TRY(cudaEventRecord(la->evU, la->inv));
  return GPINV_OK;
}
int side_early_join(cudaStream_t st) {
  LookAhead* la = lookahead_for_current_device();
  if (!la) return GPINV_ERR_CUDA;
  GPINV_TRY(cudaStreamWaitEvent(st, la->evU, 0));
  return GPINV_OK;
}

// host-side wait for the side stream (used when an inverse buffer is about to be recycled)
int trtri_side_sync_f64() {
  LookAhead* la = lookahead_for_current_device();
  if (!la) return GPINV_ERR_CUDA;
  GPINV_TRY(cudaStreamSynchronize(la->side));
  GPINV_TRY(cudaStreamSynchronize(la->inv));
  return GPINV_OK;
}

int trtri_side_join_f64(cudaStream_t st) {
  LookAhead* la = lookahead_for_current_device();
  if (!la) return GPINV_ERR_CUDA;
  GPINV_TRY(cudaStreamWaitEvent(st, la->evW, 0));
  return GPINV_OK;
}

// Solve X L^T = B in place (B is m x n row-major, L n x n lower): forward substitution over
// columns, i.e. row-wise "L x = b" for every row of B.  Used by the whitened conditional
// (GPinv/conditionals.py:44: A = Lm^{-1} Kmn, applied here to Kmn^T).
__global__ void __launch_bounds__(pcfg::TRSM_ROWS) trsm_rlt_kernel(const double* __restrict__ L,
                                                                   int ldl, double* __restrict__ X,
                                                                   int ldx, int j, int b, int r0,
                                                                   int r1) {
  using namespace pcfg;
  extern __shared__ __align__(16) double sm[];
  double* Lt = sm;                 // [NB][LDT]  Lt[k][c] = D[c][k]
  double* invd = Lt + NB * LDT;    // [NB]
  double* Xs = invd + NB;          // [TRSM_ROWS][LDR]
  const int tid = threadIdx.x;
  const int row_base = r0 + blockIdx.x * TRSM_ROWS;
  const int nrows = max(0, min(TRSM_ROWS, r1 - row_base));
  for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
    const int c = idx >> 6, k = idx & 63;   // read D[c][k] coalesced over k
    double v = (k == c) ? 1.0 : 0.0;
    if (k < b && c < b) v = (k <= c) ? L[(size_t)(j + c) * ldl + j + k] : 0.0;
    Lt[k * LDT + c] = v;
  }
  for (int idx = tid; idx < TRSM_ROWS * NB; idx += TRSM_ROWS) {
    const int r = idx >> 6, c = idx & 63;
    double v = 0.0;
    if (r < nrows && c < b) v = X[(size_t)(row_base + r) * ldx + j + c];
    Xs[r * LDR + c] = v;
  }
  __syncthreads();
  if (tid < NB) invd[tid] = 1.0 / Lt[tid * LDT + tid];
  __syncthreads();
  double* myrow = Xs + tid * LDR;
#pragma unroll 1
  for (int cb = 0; cb < NB / 8; ++cb) {
    const int c0 = cb * 8;
    double s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = myrow[c0 + q];
    for (int k = 0; k < c0; ++k) {
      const double xk = myrow[k];
      const double2* lt = reinterpret_cast<const double2*>(Lt + k * LDT + c0);
      const double2 l01 = lt[0], l23 = lt[1], l45 = lt[2], l67 = lt[3];
      s[0] -= xk * l01.x; s[1] -= xk * l01.y; s[2] -= xk * l23.x; s[3] -= xk * l23.y;
      s[4] -= xk * l45.x; s[5] -= xk * l45.y; s[6] -= xk * l67.x; s[7] -= xk * l67.y;
    }
    double x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double v = s[q];
#pragma unroll
      for (int m = 0; m < q; ++m) v -= x[m] * Lt[(c0 + m) * LDT + c0 + q];
      x[q] = v * invd[c0 + q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) myrow[c0 + q] = x[q];
  }
  __syncthreads();
  for (int idx = tid; idx < TRSM_ROWS * NB; idx += TRSM_ROWS) {
    const int r = idx >> 6, c = idx & 63;
    if (r < nrows && c < b) X[(size_t)(row_base + r) * ldx + j + c] = Xs[r * LDR + c];
  }
}

int trsm_rlt_f64(const double* L, int ldl, double* X, int ldx, int m, int n, cudaStream_t st) {
  using namespace pcfg;
  static bool attr_dev[64] = {};
  bool& attr = *device_flag(attr_dev);
  if (!attr) {
    GPINV_TRY(cudaFuncSetAttribute(trsm_rlt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TRSM_SMEM));
    attr = true;
  }
  if (m <= 0) return GPINV_OK;
  for (int j = 0; j < n; j += NB) {
    const int b = min(NB, n - j);
    trsm_rlt_kernel<<<cdiv(m, TRSM_ROWS), TRSM_ROWS, TRSM_SMEM, st>>>(L, ldl, X, ldx, j, b, 0, m);
    GPINV_CHECK_LAUNCH();
    if (j + b < n) {  // X[:, j+b:] -= X[:, j:j+b] L[j+b:, j:j+b]^T
      int rc = gemm_simple(LK, LK, m, n - j - b, b, -1.0, X + j, ldx, L + (size_t)(j + b) * ldl + j,
                           ldl, 1.0, X + j + b, ldx, false, 0, 1, st);
      if (rc) return rc;
    }
  }
  return GPINV_OK;
}

// ---------------------------------------------------------------------------------------
// Column-block sharded reverse (sample sharding over G GPUs).
//
// The reverse of the factorisation is linear in Lbar and only its contraction with dK/d(theta) is
// needed.  With  G_K = W^T Phi(L^T Lbar) W,  W = L^-1,  column block J of Phi needs column block J of
// Lbar only, and  G_K = sum_J  (W^T Phi[:,J]) W[J,:].  So the ranks reduce-scatter Lbar by column
// blocks (instead of each reversing its own partial Lbar over the whole matrix), every rank runs the
// three products of its blocks -- all with one dimension equal to the block width -- contracts its
// partial G_K with dK/d(theta), and the scalars are summed with the rest of the gradient.
// Blocks are dealt in pairs (b, 2G-1-b) of width w = n / 2G, which evens out the triangular work.
// ---------------------------------------------------------------------------------------

// out: G chunks of (n + w) w doubles; chunk r = [tril-masked Lbar[r w :, block r] ; Lbar[(2G-1-r) w :, block 2G-1-r]]
__global__ void __launch_bounds__(256) lbar_pack_kernel(const double* __restrict__ Lbar, int ldlb, int n, int w,
                                                        int G, double* __restrict__ out) {
  const int i = blockIdx.x;
  const int ncols = min(n, (i / w + 1) * w);
  const size_t chunk = (size_t)(n + w) * w;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    const int b = c / w, j = c - b * w;
    const int r = (b < G) ? b : 2 * G - 1 - b;
    const size_t off = (size_t)r * chunk + ((b < G) ? 0 : (size_t)(n - r * w) * w) + (size_t)(i - b * w) * w + j;
    out[off] = (c <= i) ? Lbar[(size_t)i * ldlb + c] : 0.0;
  }
}

int lbar_pack_f64(const double* Lbar, int ldlb, int n, int w, int G, double* out, cudaStream_t st) {
  if (n <= 0 || w <= 0 || G <= 0 || 2 * G * w != n || ldlb < n) return GPINV_ERR_ARG;
  lbar_pack_kernel<<<n, 256, 0, st>>>(Lbar, ldlb, n, w, G, out);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// lower triangle (diagonal included) of a row-major matrix <-> packed rows [i (i + 1) / 2 + c]
__global__ void __launch_bounds__(256) tri_compact_kernel(const double* __restrict__ src, int ld, int n,
                                                          double* __restrict__ dst) {
  const int i = blockIdx.x;
  const size_t o = (size_t)i * (i + 1) / 2;
  for (int c = threadIdx.x; c <= i; c += blockDim.x) dst[o + c] = src[(size_t)i * ld + c];
}
__global__ void __launch_bounds__(256) tri_expand_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                         int ld, int n) {
  const int i = blockIdx.x;
  const size_t o = (size_t)i * (i + 1) / 2;
  for (int c = threadIdx.x; c <= i; c += blockDim.x) dst[(size_t)i * ld + c] = src[o + c];
}
int tri_compact_f64(const double* src, int ld, int n, double* dst, cudaStream_t st) {
  if (n <= 0 || ld < n) return GPINV_ERR_ARG;
  tri_compact_kernel<<<n, 256, 0, st>>>(src, ld, n, dst);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}
int tri_expand_f64(const double* src, double* dst, int ld, int n, cudaStream_t st) {
  if (n <= 0 || ld < n) return GPINV_ERR_ARG;
  tri_expand_kernel<<<n, 256, 0, st>>>(src, dst, ld, n);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// Gacc (n x n, ld n) <- sum over this rank's blocks of (W^T Phi_J) W[J,:].  Dsh: the reduced, tril-masked
// column blocks of Lbar, block b stored as (n - c0s[b]) x w right behind block b-1.  P: n x w, A: n x w scratch.
int potrf_bwd_shard_f64(const double* L, int ldl, const double* W, int ldw, int n, int w, const int* c0s,
                        int nblk, const double* Dsh, double* Gacc, double* P, double* A, cudaStream_t st) {
  if (n <= 0 || w <= 0 || (w & 1) || nblk <= 0) return GPINV_ERR_ARG;
  GPINV_TRY(cudaMemsetAsync(Gacc, 0, (size_t)n * n * sizeof(double), st));
  const double* D = Dsh;
  dim3 blk(32, 8), grd(cdiv(w, 32), cdiv(w, 8));
  for (int b = 0; b < nblk; ++b) {
    const int c0 = c0s[b], m = n - c0;
    if (c0 < 0 || c0 + w > n || (c0 & 1)) return GPINV_ERR_ARG;
    // P = L[c0:, c0:]^T D   (k >= row), then Phi on the w x w square at the top: tril, diagonal halved
    int rc = gemm_ws(LMN, LMN, EF_BETA, m, w, m, 1.0, L + (size_t)c0 * ldl + c0, ldl, D, w, 0.0, nullptr, 0, P, w,
                     4, st);
    if (rc) return rc;
    zero_upper_kernel<<<grd, blk, 0, st>>>(P, w, w);
    GPINV_CHECK_LAUNCH();
    scale_diag_kernel<<<cdiv(w, 256), 256, 0, st>>>(P, w, w, 0.5);
    GPINV_CHECK_LAUNCH();
    // A = W[c0:, :]^T Phi_J : rows above the block are dense in k, rows from c0 on have k >= row
    if (c0 > 0) {
      rc = gemm_ws(LMN, LMN, EF_BETA, c0, w, m, 1.0, W + (size_t)c0 * ldw, ldw, P, w, 0.0, nullptr, 0, A, w, 0, st);
      if (rc) return rc;
    }
    rc = gemm_ws(LMN, LMN, EF_BETA, m, w, m, 1.0, W + (size_t)c0 * ldw + c0, ldw, P, w, 0.0, nullptr, 0,
                 A + (size_t)c0 * w, w, 4, st);
    if (rc) return rc;
    // Gacc[:, :c0+w] += A W[J, :c0+w]     (W is lower triangular with an exactly zero upper part)
    rc = gemm_ws(LK, LMN, EF_BETA, n, c0 + w, w, 1.0, A, w, W + (size_t)c0 * ldw, ldw, 1.0, Gacc, n, Gacc, n, 0, st);
    if (rc) return rc;
    D += (size_t)m * w;
  }
  return GPINV_OK;
}

}  // namespace gpinv
