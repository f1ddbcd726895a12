// Shared helpers for the gpinv_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GPINV_OK 0
#define GPINV_ERR_ARG (-1)
#define GPINV_ERR_CUDA (-2)
#define GPINV_ERR_WS (-3)

namespace gpinv {

// FP64 tensor-core MMA on sm_100a: mma.sync m8n8k4 lowers to one DMMA.8x8x4 (SASS).
// Fragment ownership (lane = 4*g + t):  A[g][t]  B[t][g]  C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared (LDGSTS), zero-filling bytes beyond src_bytes.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
// 8-byte async copy (for operands whose leading dimension / base is not 16-byte aligned).
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}

// 1/sqrt(v): SFU single-precision seed + two Newton steps in fp64 (<= ~1 ulp for v inside the
// float range; falls back to the library routine outside it, and for v <= 0 returns NaN/inf like
// rsqrt()).  About half the latency of rsqrt(double) on the Cholesky pivot chain.
__device__ __forceinline__ double fast_rsqrt(double v) {
  if (!(v > 1e-30 && v < 1e30)) return rsqrt(v);
  // single-precision seed (relative error e0 ~ 2^-22) + ONE cubically convergent (Halley) step:
  // y = y0 (1 + e/2 + 3 e^2/8), e = 1 - v y0^2; remaining error O(e^3) ~ 2^-64.  Four dependent
  // fp64 operations instead of the six of two Newton steps: this sits on the pivot chain.
  const double y0 = (double)rsqrtf((float)v);
  const double t = v * y0;
  const double e = fma(-t, y0, 1.0);
  const double ye = y0 * e;
  const double p = fma(0.375, e, 0.5);
  return fma(ye, p, y0);
}
// sqrt(v) from y ~ 1/sqrt(v) with one correction step
__device__ __forceinline__ double sqrt_from_rsqrt(double v, double y) {
  double l = v * y;
  return fma(fma(-l, l, v), 0.5 * y, l);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace gpinv

#define GPINV_CHECK_LAUNCH()                                  \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) {                                 \
      gpinv::set_last_error(cudaGetErrorString(e__));         \
      return GPINV_ERR_CUDA;                                  \
    }                                                         \
  } while (0)

namespace gpinv {
void set_last_error(const char* msg);
}

// Run a CUDA runtime call; on failure record its error string and return GPINV_ERR_CUDA.
#define GPINV_TRY(call)                                       \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) {                                 \
      gpinv::set_last_error(cudaGetErrorString(e__));         \
      return GPINV_ERR_CUDA;                                  \
    }                                                         \
  } while (0)
