// The callers either side of the hot path, kept on the device:
//
//  * free-state <-> parameter transform and gradient packing (GPflow Param.get_free_state /
//    set_state with transforms.positive = softplus + 1e-6, the x that GPinv/model.py:64-75 hands to
//    the objective): ONE kernel unpacks every block of x into its constrained value, ONE kernel
//    applies the chain rule y' = sigmoid(x) and writes the packed gradient;
//  * the analytic whitened KL of GPinv/stvgp.py:187-195 (GPflow gauss_kl_white[_diag]) and its
//    gradient, reading q_sqrt in the reference's [n,n,R] layout (lower triangle only);
//  * the optimiser step of the training loop (GPinv/model.py:206-240 drives a tf.train optimiser;
//    testing/test_stvgp.py:31-35 uses tf.train.AdamOptimizer): TensorFlow's Adam update applied in
//    place to the device-resident state, graph-capturable (the step counter lives on the device).
//
// All HBM-bound streaming kernels (16-byte accesses where the block is aligned).
#include "common.cuh"

namespace gpinv {

constexpr int MAX_BLOCKS = 32;
struct PackDesc {
  int nblocks;
  long long off[MAX_BLOCKS + 1];   // block b covers x[off[b], off[b+1])
  int transform[MAX_BLOCKS];       // 0 identity, 1 positive (softplus + 1e-6)
  double* dst[MAX_BLOCKS];         // unpack: constrained values;  pack: incoming dObj/d(value)
};

__device__ __forceinline__ double softplus_pos(double x) {
  // log(1 + exp(x)) + 1e-6, stable for large |x| (matches torch.nn.functional.softplus, threshold off)
  return (x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x))) + 1e-6;
}
__device__ __forceinline__ double sigmoid(double x) {
  return x >= 0.0 ? 1.0 / (1.0 + exp(-x)) : exp(x) / (1.0 + exp(x));
}

// grid.y = block index
__global__ void __launch_bounds__(256) unpack_kernel(const double* __restrict__ x, const PackDesc d) {
  const int b = blockIdx.y;
  const long long lo = d.off[b], cnt = d.off[b + 1] - lo;
  const double* src = x + lo;
  double* dst = d.dst[b];
  const int tr = d.transform[b];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
    const long long n2 = cnt / 2;
    for (long long k = i0; k < n2; k += stride) {
      double2 v = reinterpret_cast<const double2*>(src)[k];
      if (tr) { v.x = softplus_pos(v.x); v.y = softplus_pos(v.y); }
      reinterpret_cast<double2*>(dst)[k] = v;
    }
    if (i0 == 0 && (cnt & 1)) dst[cnt - 1] = tr ? softplus_pos(src[cnt - 1]) : src[cnt - 1];
  } else {
    for (long long k = i0; k < cnt; k += stride) dst[k] = tr ? softplus_pos(src[k]) : src[k];
  }
}

// gx[off[b] + k] = scale * gval_b[k] * (positive ? sigmoid(x) : 1);  a NULL gval block gives zeros
__global__ void __launch_bounds__(256) pack_kernel(const double* __restrict__ x, const PackDesc d, double scale,
                                                   double* __restrict__ gx) {
  const int b = blockIdx.y;
  const long long lo = d.off[b], cnt = d.off[b + 1] - lo;
  const double* xs = x + lo;
  const double* g = d.dst[b];
  double* out = gx + lo;
  const int tr = d.transform[b];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g && !tr && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    const long long n2 = cnt / 2;
    for (long long k = i0; k < n2; k += stride) {
      double2 v = reinterpret_cast<const double2*>(g)[k];
      v.x *= scale; v.y *= scale;
      reinterpret_cast<double2*>(out)[k] = v;
    }
    if (i0 == 0 && (cnt & 1)) out[cnt - 1] = scale * g[cnt - 1];
    return;
  }
  for (long long k = i0; k < cnt; k += stride) {
    double v = g ? scale * g[k] : 0.0;
    if (tr && g) v *= sigmoid(xs[k]);
    out[k] = v;
  }
}

static int pack_grid(const PackDesc& d) {
  long long mx = 1;
  for (int b = 0; b < d.nblocks; ++b) mx = (d.off[b + 1] - d.off[b] > mx) ? d.off[b + 1] - d.off[b] : mx;
  const long long g = (mx / 2 + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

int unpack_f64(const double* x, int nblocks, const long long* off, const int* transform, double* const* dst,
               cudaStream_t st) {
  if (nblocks <= 0 || nblocks > MAX_BLOCKS) return GPINV_ERR_ARG;
  PackDesc d;
  d.nblocks = nblocks;
  for (int b = 0; b < nblocks; ++b) { d.off[b] = off[b]; d.transform[b] = transform[b]; d.dst[b] = dst[b]; }
  d.off[nblocks] = off[nblocks];
  dim3 grid(pack_grid(d), nblocks);
  unpack_kernel<<<grid, 256, 0, st>>>(x, d);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int pack_f64(const double* x, int nblocks, const long long* off, const int* transform, const double* const* gval,
             double scale, double* gx, cudaStream_t st) {
  if (nblocks <= 0 || nblocks > MAX_BLOCKS) return GPINV_ERR_ARG;
  PackDesc d;
  d.nblocks = nblocks;
  for (int b = 0; b < nblocks; ++b) {
    d.off[b] = off[b]; d.transform[b] = transform[b]; d.dst[b] = const_cast<double*>(gval[b]);
  }
  d.off[nblocks] = off[nblocks];
  dim3 grid(pack_grid(d), nblocks);
  pack_kernel<<<grid, 256, 0, st>>>(x, d, scale, gx);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// ------------------------------ analytic whitened KL ------------------------------
// out[0] += 1/2 sum m^2 ; out[1] += sum_{r,i} log Q_ii^2 ; out[2] += sum over the lower triangle of Q^2
// q_sqrt: [n,n,R] (full, element (i,j,r) at (i*n + j)*R + r) or [n,R] (diag)
__global__ void __launch_bounds__(256) kl_white_kernel(const double* __restrict__ q_mu, const double* __restrict__ q_sqrt,
                                                       int n, int R, int q_diag, double* __restrict__ out) {
  __shared__ double red[32];
  double s_m = 0.0, s_log = 0.0, s_q = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nm = (long long)n * R;
  for (long long k = i0; k < nm; k += stride) {
    const double m = q_mu[k];
    s_m += m * m;
    if (q_diag) {
      const double q = q_sqrt[k];
      s_log += log(q * q);
      s_q += q * q;
    }
  }
  if (!q_diag) {
    const long long rowlen = (long long)n * R;
    const long long total = (long long)n * rowlen;
    for (long long k = i0; k < total; k += stride) {
      const long long i = k / rowlen, rem = k - i * rowlen;
      const long long j = rem / R;
      if (j > i) continue;
      const double q = q_sqrt[k];
      s_q += q * q;
      if (j == i) s_log += log(q * q);
    }
  }
  s_m = block_sum(s_m, red);
  if (threadIdx.x == 0) atomicAdd(out + 0, 0.5 * s_m);
  s_log = block_sum(s_log, red);
  if (threadIdx.x == 0) atomicAdd(out + 1, s_log);
  s_q = block_sum(s_q, red);
  if (threadIdx.x == 0) atomicAdd(out + 2, s_q);
}

// KL = 1/2 sum m^2 - 1/2 nR - 1/2 sum log Q_ii^2 + 1/2 sum Q^2
__global__ void kl_white_finish_kernel(const double* __restrict__ parts, double half_count, double* __restrict__ kl) {
  kl[0] = parts[0] - half_count - 0.5 * parts[1] + 0.5 * parts[2];
}

// dKL/dq_mu = m ; dKL/dQ = Q - diag(1/Q_ii) on the lower triangle, 0 above; both scaled by gbar[0]
__global__ void __launch_bounds__(256) kl_white_bwd_kernel(const double* __restrict__ q_mu, const double* __restrict__ q_sqrt,
                                                           int n, int R, int q_diag, const double* __restrict__ gbar,
                                                           double* __restrict__ mu_bar, double* __restrict__ q_bar) {
  const double g = gbar[0];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nm = (long long)n * R;
  for (long long k = i0; k < nm; k += stride) {
    mu_bar[k] = g * q_mu[k];
    if (q_diag) {
      const double q = q_sqrt[k];
      q_bar[k] = g * (q - 1.0 / q);
    }
  }
  if (!q_diag) {
    const long long rowlen = (long long)n * R;
    const long long total = (long long)n * rowlen;
    for (long long k = i0; k < total; k += stride) {
      const long long i = k / rowlen, rem = k - i * rowlen;
      const long long j = rem / R;
      double v = 0.0;
      if (j <= i) {
        const double q = q_sqrt[k];
        v = g * ((j == i) ? q - 1.0 / q : q);
      }
      q_bar[k] = v;
    }
  }
}

int kl_white_f64(const double* q_mu, const double* q_sqrt, int n, int R, int q_diag, double* kl_out /* [4] */,
                 cudaStream_t st) {
  if (n <= 0 || R <= 0) return GPINV_ERR_ARG;
  GPINV_TRY(cudaMemsetAsync(kl_out, 0, 4 * sizeof(double), st));
  const long long total = q_diag ? (long long)n * R : (long long)n * n * R;
  long long g = (total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  kl_white_kernel<<<(int)g, 256, 0, st>>>(q_mu, q_sqrt, n, R, q_diag, kl_out + 1);
  GPINV_CHECK_LAUNCH();
  kl_white_finish_kernel<<<1, 1, 0, st>>>(kl_out + 1, 0.5 * (double)n * R, kl_out);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

int kl_white_bwd_f64(const double* q_mu, const double* q_sqrt, int n, int R, int q_diag, const double* gbar,
                     double* mu_bar, double* q_bar, cudaStream_t st) {
  if (n <= 0 || R <= 0) return GPINV_ERR_ARG;
  const long long total = q_diag ? (long long)n * R : (long long)n * n * R;
  long long g = (total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  kl_white_bwd_kernel<<<(int)g, 256, 0, st>>>(q_mu, q_sqrt, n, R, q_diag, gbar, mu_bar, q_bar);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// ------------------------------------ Adam ------------------------------------
// TensorFlow's AdamOptimizer update (the optimiser of testing/test_stvgp.py:31-35):
//   t <- t + 1;  lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;
//   x <- x - lr_t m / (sqrt(v) + eps)
// `step` is a device counter (so that the launch is identical from one CUDA-graph replay to the next);
// it is advanced by adam_bump_kernel after every thread has read it.
__global__ void __launch_bounds__(256) adam_step_kernel(double* __restrict__ x, double* __restrict__ m,
                                                        double* __restrict__ v, const double* __restrict__ g,
                                                        long long count, double lr, double b1, double b2,
                                                        double eps, const long long* __restrict__ step) {
  const double t = (double)(step[0] + 1);
  const double lr_t = lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n2 = count / 2;
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* m2 = reinterpret_cast<double2*>(m);
  double2* v2 = reinterpret_cast<double2*>(v);
  const double2* g2 = reinterpret_cast<const double2*>(g);
  for (long long k = i0; k < n2; k += stride) {
    const double2 gg = g2[k];
    double2 mm = m2[k], vv = v2[k], xx = x2[k];
    mm.x = b1 * mm.x + (1.0 - b1) * gg.x;  mm.y = b1 * mm.y + (1.0 - b1) * gg.y;
    vv.x = b2 * vv.x + (1.0 - b2) * gg.x * gg.x;  vv.y = b2 * vv.y + (1.0 - b2) * gg.y * gg.y;
    xx.x -= lr_t * mm.x / (sqrt(vv.x) + eps);  xx.y -= lr_t * mm.y / (sqrt(vv.y) + eps);
    m2[k] = mm; v2[k] = vv; x2[k] = xx;
  }
  if (i0 == 0 && (count & 1)) {
    const long long k = count - 1;
    const double gg = g[k];
    const double mm = b1 * m[k] + (1.0 - b1) * gg, vv = b2 * v[k] + (1.0 - b2) * gg * gg;
    m[k] = mm; v[k] = vv;
    x[k] -= lr_t * mm / (sqrt(vv) + eps);
  }
}
__global__ void adam_bump_kernel(long long* step) { step[0] += 1; }

// x, m, v, g: [count], 16-byte aligned
int adam_step_f64(double* x, double* m, double* v, const double* g, long long count, double lr, double b1,
                  double b2, double eps, long long* step, cudaStream_t st) {
  if (count <= 0) return GPINV_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(g)) & 15u) {
    set_last_error("adam_step: x, m, v, g must be 16-byte aligned");
    return GPINV_ERR_ARG;
  }
  long long grid = (count / 2 + 255) / 256;
  if (grid < 1) grid = 1;
  if (grid > 148 * 8) grid = 148 * 8;
  adam_step_kernel<<<(int)grid, 256, 0, st>>>(x, m, v, g, count, lr, b1, b2, eps, step);
  GPINV_CHECK_LAUNCH();
  adam_bump_kernel<<<1, 1, 0, st>>>(step);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// ---------------------------------------------------------------------------------------
// Gaussian log-density summed over `count` observations from the residual sum of squares (GPflow
// densities.gaussian behind GPinv/likelihoods.py:70-76 and the notebooks' likelihoods):
//   out = -count/2 (log 2 pi + log var) - ss / (2 var)
// and its reverse  ss_bar = -g / (2 var),  var_bar = g (-count / (2 var) + ss / (2 var^2)).
// One thread each: these replace ~15 scalar elementwise launches of the ELBO assembly.
// ---------------------------------------------------------------------------------------
__global__ void gauss_logp_kernel(const double* __restrict__ ss, const double* __restrict__ var, double count,
                                  double* __restrict__ out) {
  const double v = var[0];
  out[0] = -0.5 * count * (1.8378770664093453 + log(v)) - 0.5 * ss[0] / v;
}
__global__ void gauss_logp_bwd_kernel(const double* __restrict__ ss, const double* __restrict__ var, double count,
                                      const double* __restrict__ g, double* __restrict__ ss_bar,
                                      double* __restrict__ var_bar) {
  const double v = var[0], gg = g[0];
  ss_bar[0] = -0.5 * gg / v;
  var_bar[0] = gg * (-0.5 * count / v + 0.5 * ss[0] / (v * v));
}
int gauss_logp_f64(const double* ss, const double* var, double count, double* out, cudaStream_t st) {
  gauss_logp_kernel<<<1, 1, 0, st>>>(ss, var, count, out);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}
int gauss_logp_bwd_f64(const double* ss, const double* var, double count, const double* g, double* ss_bar,
                       double* var_bar, cudaStream_t st) {
  gauss_logp_bwd_kernel<<<1, 1, 0, st>>>(ss, var, count, g, ss_bar, var_bar);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

}  // namespace gpinv
