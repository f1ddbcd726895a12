// Kernel-matrix build (T1-T3) and its lengthscale gradient.
//   core(x,x') = exp(-D/2)                                   RBF       GPinv/kernels.py:105-107
//              = exp(-D(|x|,|x'|)/2) + exp(-D(|x|,-|x'|)/2)  RBF_csym  GPinv/kernels.py:115-120
//              = exp(-D(|x|,|x'|)/2) - exp(-D(|x|,-|x'|)/2)  RBF_casym GPinv/kernels.py:140-145
// with D in GPflow's expansion form  D_ij = (-2 <x_i/l, x'_j/l> + |x_i/l|^2) + |x'_j/l|^2
// (Stationary.square_dist), and jitter added to the diagonal of the *core*
// (GPinv/kernels.py:57-58).  HBM-bound: 8 n^2 bytes written forward, 8 n^2 read backward.
#include "common.cuh"

namespace gpinv {

// xs[i*D+d] = (mode ? |x| : x)/ls_d ; sq[i] = sum_d xs^2
__global__ void kprep_kernel(const double* __restrict__ X, int n, int D, int ldx,
                             const double* __restrict__ ls, int ard, int mode,
                             double* __restrict__ xs, double* __restrict__ sq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    double x = X[(size_t)i * ldx + d];
    if (mode) x = fabs(x);
    const double v = x / ls[ard ? d : 0];
    xs[(size_t)i * D + d] = v;
    s = __dadd_rn(s, __dmul_rn(v, v));
  }
  sq[i] = s;
}

template <int MODE>
__device__ __forceinline__ double kcore_eval(double dot, double sqi, double sqj) {
  // expansion form, additions in the reference's order, no FMA contraction
  const double d1 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), sqi), sqj);
  const double e1 = exp(-0.5 * d1);
  if (MODE == 0) return e1;
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(2.0, dot), sqi), sqj);
  const double e2 = exp(-0.5 * d2);
  return (MODE == 1) ? e1 + e2 : e1 - e2;
}

// One thread computes 2 adjacent columns x 4 rows; 128-bit stores when aligned.
template <int MODE>
__global__ void __launch_bounds__(256) kbuild_kernel(const double* __restrict__ xs1,
                                                     const double* __restrict__ sq1,
                                                     const double* __restrict__ xs2,
                                                     const double* __restrict__ sq2, int n, int n2,
                                                     int D, double jitter, int add_jitter,
                                                     double* __restrict__ K, int ldk, int vec,
                                                     int lower_only) {
  const int c = 2 * (blockIdx.x * 64 + (threadIdx.x & 63));
  const int rb = blockIdx.y * 16 + (threadIdx.x >> 6) * 4;
  if (c >= n2) return;
  const bool two = (c + 1 < n2);
  const double sj0 = sq2[c], sj1 = two ? sq2[c + 1] : 0.0;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = rb + rr;
    if (r >= n) break;
    if (lower_only && c > r) {   // strictly above the diagonal: exact zeros, no exp
      double* o = K + (size_t)r * ldk + c;
      if (two && vec) *reinterpret_cast<double2*>(o) = make_double2(0.0, 0.0);
      else {
        o[0] = 0.0;
        if (two) o[1] = 0.0;
      }
      continue;
    }
    double dot0 = 0.0, dot1 = 0.0;
    for (int d = 0; d < D; ++d) {
      const double xi = xs1[(size_t)r * D + d];
      dot0 = __dadd_rn(dot0, __dmul_rn(xi, xs2[(size_t)c * D + d]));
      if (two) dot1 = __dadd_rn(dot1, __dmul_rn(xi, xs2[(size_t)(c + 1) * D + d]));
    }
    const double si = sq1[r];
    double v0 = kcore_eval<MODE>(dot0, si, sj0);
    double v1 = two ? kcore_eval<MODE>(dot1, si, sj1) : 0.0;
    if (add_jitter) {
      if (r == c) v0 += jitter;
      if (r == c + 1) v1 += jitter;
    }
    if (lower_only && c + 1 > r) v1 = 0.0;
    double* o = K + (size_t)r * ldk + c;
    if (two && vec) *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
    else {
      o[0] = v0;
      if (two) o[1] = v1;
    }
  }
}

// ls_bar[dsel or 0] += (1/l) sum_ij Kbar_ij (e1 D1_d +/- e2 D2_d), D_d summed over the
// selected dimension(s): dsel >= 0 picks one ARD dimension, dsel < 0 sums all (shared l).
template <int MODE>
__global__ void __launch_bounds__(256) kbuild_bwd_kernel(const double* __restrict__ xs1,
                                                         const double* __restrict__ sq1,
                                                         const double* __restrict__ xs2,
                                                         const double* __restrict__ sq2, int n,
                                                         int n2, int D, int dsel,
                                                         const double* __restrict__ ls,
                                                         const double* __restrict__ Kbar, int ldkb,
                                                         double* __restrict__ ls_bar, int lower_only) {
  __shared__ double red[32];
  double acc = 0.0;
  const int c = 2 * (blockIdx.x * 64 + (threadIdx.x & 63));
  const int rb = blockIdx.y * 16 + (threadIdx.x >> 6) * 4;
  if (c < n2) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = rb + rr;
      if (r >= n) break;
      for (int e = 0; e < 2; ++e) {
        const int cc = c + e;
        if (cc >= n2) break;
        if (lower_only && cc > r) break;   // Kbar holds a lower triangle (strict part pre-doubled)
        double dot = 0.0, dsel_dot = 0.0, dsel_sq = 0.0;
        for (int d = 0; d < D; ++d) {
          const double xi = xs1[(size_t)r * D + d], xj = xs2[(size_t)cc * D + d];
          const double pr = __dmul_rn(xi, xj);
          dot = __dadd_rn(dot, pr);
          if (dsel < 0 || d == dsel) {
            dsel_dot += pr;
            dsel_sq += xi * xi + xj * xj;
          }
        }
        const double si = sq1[r], sj = sq2[cc];
        const double d1 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), si), sj);
        double g = exp(-0.5 * d1) * (dsel_sq - 2.0 * dsel_dot);
        if (MODE != 0) {
          const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(2.0, dot), si), sj);
          const double g2 = exp(-0.5 * d2) * (dsel_sq + 2.0 * dsel_dot);
          g = (MODE == 1) ? g + g2 : g - g2;
        }
        acc += Kbar[(size_t)r * ldkb + cc] * g;
      }
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && acc != 0.0) {
    const int li = dsel < 0 ? 0 : dsel;
    atomicAdd(ls_bar + li, acc / ls[li]);
  }
}

size_t kbuild_ws_bytes(int n, int n2, int D) { return (size_t)(n + n2) * (D + 1) * sizeof(double); }

static void kprep(const double* X, int n, int D, int ldx, const double* ls, int ard, int mode,
                  double* xs, double* sq, cudaStream_t st) {
  kprep_kernel<<<cdiv(n, 256), 256, 0, st>>>(X, n, D, ldx, ls, ard, mode, xs, sq);
}

int kbuild_ex_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
                  const double* ls, int ard, int mode, double jitter, double* K, int ldk, double* ws,
                  int lower_only, cudaStream_t st);

int kbuild_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
               const double* ls, int ard, int mode, double jitter, double* K, int ldk, double* ws,
               cudaStream_t st) {
  return kbuild_ex_f64(X, X2, n, n2, D, ldx, ldx2, ls, ard, mode, jitter, K, ldk, ws, 0, st);
}

// lower_only (X2 == NULL): the strict upper triangle is written as exact zeros and never evaluated
int kbuild_ex_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
                  const double* ls, int ard, int mode, double jitter, double* K, int ldk, double* ws,
                  int lower_only, cudaStream_t st) {
  if (n <= 0 || D <= 0 || mode < 0 || mode > 2) return GPINV_ERR_ARG;
  double* xs1 = ws;
  double* sq1 = xs1 + (size_t)n * D;
  double *xs2 = xs1, *sq2 = sq1;
  kprep(X, n, D, ldx, ls, ard, mode, xs1, sq1, st);
  if (X2) {
    xs2 = sq1 + n;
    sq2 = xs2 + (size_t)n2 * D;
    kprep(X2, n2, D, ldx2, ls, ard, mode, xs2, sq2, st);
  } else {
    n2 = n;
  }
  const int add_j = (X2 == nullptr) && (jitter != 0.0);
  const int vec = ((reinterpret_cast<uintptr_t>(K) & 15u) == 0) && (ldk % 2 == 0);
  dim3 grid(cdiv(n2, 128), cdiv(n, 16));
  if (X2) lower_only = 0;
  if (mode == 0) kbuild_kernel<0><<<grid, 256, 0, st>>>(xs1, sq1, xs2, sq2, n, n2, D, jitter, add_j, K, ldk, vec, lower_only);
  else if (mode == 1) kbuild_kernel<1><<<grid, 256, 0, st>>>(xs1, sq1, xs2, sq2, n, n2, D, jitter, add_j, K, ldk, vec, lower_only);
  else kbuild_kernel<2><<<grid, 256, 0, st>>>(xs1, sq1, xs2, sq2, n, n2, D, jitter, add_j, K, ldk, vec, lower_only);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int kbuild_bwd_ex_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
                      const double* ls, int ard, int mode, const double* Kbar, int ldkb,
                      double* ls_bar, double* ws, int lower_only, cudaStream_t st);

int kbuild_bwd_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
                   const double* ls, int ard, int mode, const double* Kbar, int ldkb,
                   double* ls_bar, double* ws, cudaStream_t st) {
  return kbuild_bwd_ex_f64(X, X2, n, n2, D, ldx, ldx2, ls, ard, mode, Kbar, ldkb, ls_bar, ws, 0, st);
}

// lower_only (X2 == NULL): Kbar is a lower triangle whose strict part already carries the factor
// 2 of the symmetric pair (the blocked Cholesky reverse sweep leaves it in that convention)
int kbuild_bwd_ex_f64(const double* X, const double* X2, int n, int n2, int D, int ldx, int ldx2,
                      const double* ls, int ard, int mode, const double* Kbar, int ldkb,
                      double* ls_bar, double* ws, int lower_only, cudaStream_t st) {
  if (n <= 0 || D <= 0 || mode < 0 || mode > 2) return GPINV_ERR_ARG;
  double* xs1 = ws;
  double* sq1 = xs1 + (size_t)n * D;
  double *xs2 = xs1, *sq2 = sq1;
  kprep(X, n, D, ldx, ls, ard, mode, xs1, sq1, st);
  if (X2) {
    xs2 = sq1 + n;
    sq2 = xs2 + (size_t)n2 * D;
    kprep(X2, n2, D, ldx2, ls, ard, mode, xs2, sq2, st);
  } else {
    n2 = n;
  }
  const int nls = ard ? D : 1;
  GPINV_TRY(cudaMemsetAsync(ls_bar, 0, nls * sizeof(double), st));
  dim3 grid(cdiv(n2, 128), cdiv(n, 16));
  for (int l = 0; l < nls; ++l) {
    const int dsel = ard ? l : -1;
    const int lo = X2 ? 0 : lower_only;
    if (mode == 0) kbuild_bwd_kernel<0><<<grid, 256, 0, st>>>(xs1, sq1, xs2, sq2, n, n2, D, dsel, ls, Kbar, ldkb, ls_bar, lo);
    else if (mode == 1) kbuild_bwd_kernel<1><<<grid, 256, 0, st>>>(xs1, sq1, xs2, sq2, n, n2, D, dsel, ls, Kbar, ldkb, ls_bar, lo);
    else kbuild_bwd_kernel<2><<<grid, 256, 0, st>>>(xs1, sq1, xs2, sq2, n, n2, D, dsel, ls, Kbar, ldkb, ls_bar, lo);
    GPINV_CHECK_LAUNCH();
  }
  return GPINV_OK;
}

}  // namespace gpinv
