// Host-side dispatch for the DMMA GEMM template (see gemm.cuh).
#include "gemm_tma.cuh"

#include <cstdio>
#include <cstdlib>

namespace gpinv {

namespace {

template <class Cfg, int LA, int LB, unsigned EF>
int launch_persistent(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set[64] = {};     // the attribute is per device
  auto kern = gemm_dmma_persistent_kernel<Cfg, LA, LB, EF>;
  int dev = 0;
  GPINV_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    GPINV_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

template <class Cfg, int LA, int LB, unsigned EF>
int launch_inst(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set[64] = {};     // the attribute is per device
  auto kern = gemm_dmma_kernel<Cfg, LA, LB, EF>;
  int dev = 0;
  GPINV_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    GPINV_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

// split-K workspace registry: [int counters[CNT_SLOTS]] [double partial tiles ...]
constexpr int CNT_SLOTS = 16384;
struct WsReg { void* ptr = nullptr; size_t bytes = 0; };
WsReg g_ws[64];
inline WsReg* ws_for_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return g_ws[dev].ptr ? &g_ws[dev] : nullptr;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int tiles_of(int M, int N, int bm, bool lower) {
  const int tm = cdiv(M, bm), tn = cdiv(N, bm);
  return lower ? tm * (tm + 1) / 2 : tm * tn;
}

// Small tiles when the big-tile grid cannot fill the machine or K is too short to amortise a
// 128x128 tile's fill / epilogue.
inline bool use_small(const GemmParams& p, bool lower) {
  if (p.force_cfg == 1 || p.force_cfg == 3) return false;
  if (p.force_cfg == 2) return true;
  const int batch = p.batch > 1 ? p.batch : 1;
  const long long big_tiles = (long long)tiles_of(p.M, p.N, 128, lower) * batch * (p.splitk > 1 ? p.splitk : 1);
  if (big_tiles < 120) return true;
  if (p.K <= 128 && big_tiles < 600) return true;
  return false;
}

template <class Cfg>
int dispatch(int la, int lb, unsigned ef, GemmParams p, cudaStream_t stream);

}  // namespace

bool gemm_workspace_for_device(void** ptr, size_t* bytes) {
  WsReg* w = ws_for_device();
  if (!w) return false;
  *ptr = w->ptr;
  *bytes = w->bytes;
  return true;
}

// Tile configuration for the TMA / stream-K path: the 128x128 tile whenever the problem holds enough
// k-tile work units to give every SM a useful share (stream-K fills the machine regardless of the
// tile count), the 64x64 tile (3 CTAs/SM) for short-K or tiny problems.
static bool use_small_tma(const GemmParams& p, bool lower) {
  static const char* force = getenv("GPINV_TMA_FORCE");     // debugging: big | small
  if (force && force[0] == 'b') return false;
  if (force && force[0] == 's') return true;
  if (p.force_cfg == 1 || p.force_cfg == 3) return false;
  if (p.force_cfg == 2) return true;
  const int batch = p.batch > 1 ? p.batch : 1;
  const long long tiles = (long long)tiles_of(p.M, p.N, 128, lower) * batch;
  const long long units = tiles * cdiv(p.K, 16) / ((p.kmode != 0) ? 2 : 1);
  if (p.K <= 128) return tiles < 600;
  return units < 148 * 16;
}

void gemm_set_workspace(void* ptr, size_t bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  g_ws[dev].ptr = ptr;
  g_ws[dev].bytes = bytes;
}

int gemm_tile_count(int M, int N, int K, bool lower, int batch) {
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.splitk = 1;
  const int bm = use_small(p, lower) ? 64 : 128;
  return tiles_of(M, N, bm, lower) * (batch > 1 ? batch : 1);
}

#define GPINV_GEMM_CASE(LA_, LB_, EF_)                                   \
  if (la == (LA_) && lb == (LB_) && ef == (unsigned)(EF_))               \
    return launch_inst<Cfg, LA_, LB_, (unsigned)(EF_)>(p, grid, stream);

#define GPINV_GEMM_GENERIC(LA_, LB_)                      \
  GPINV_GEMM_CASE(LA_, LB_, EF_BETA)                      \
  GPINV_GEMM_CASE(LA_, LB_, EF_BETA | EF_LOWER)           \
  GPINV_GEMM_CASE(LA_, LB_, EF_ATOMIC)                    \
  GPINV_GEMM_CASE(LA_, LB_, EF_ATOMIC | EF_LOWER)

namespace {
template <class Cfg>
int dispatch(int la, int lb, unsigned ef, GemmParams p, cudaStream_t stream) {
  p.tiles_m = cdiv(p.M, Cfg::BM);
  p.tiles_n = cdiv(p.N, Cfg::BN);
  if ((ef & EF_LOWER) && p.tiles_m != p.tiles_n) {
    set_last_error("gemm: lower-triangular output must be square");
    return GPINV_ERR_ARG;
  }
  const int ntiles = (ef & EF_LOWER) ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  dim3 grid(ntiles, p.splitk, p.batch);
  p.ntiles = ntiles;
  if (p.max_ctas > 0 && p.splitk == 1 && p.batch == 1 && ntiles > p.max_ctas) {
    // persistent, CTA-capped instantiations (only the combinations that run beside the
    // Cholesky panel chain exist)
    grid.x = p.max_ctas;
    if (la == LK && lb == LK && ef == (EF_BETA | EF_LOWER))
      return launch_persistent<Cfg, LK, LK, EF_BETA | EF_LOWER>(p, grid, stream);
    if (la == LK && lb == LK && ef == (EF_COLVEC | EF_SUMSQ))
      return launch_persistent<Cfg, LK, LK, EF_COLVEC | EF_SUMSQ>(p, grid, stream);
    grid.x = ntiles;
  }

  GPINV_GEMM_GENERIC(LK, LK)
  GPINV_GEMM_GENERIC(LK, LMN)
  GPINV_GEMM_GENERIC(LMN, LMN)
  GPINV_GEMM_GENERIC(LMN, LK)
  // fused epilogues of the StVGP sample / Abel stages
  GPINV_GEMM_CASE(LK, LK, EF_COLVEC | EF_SUMSQ)               // U = E Q^T + m ; R = Y - X A^T
  GPINV_GEMM_CASE(LK, LK, EF_COLVEC)                          // F = sv U Lc^T + mu
  GPINV_GEMM_CASE(LK, LK, EF_COLVEC | EF_EXP2)                // ... and exp(F)
  GPINV_GEMM_CASE(LK, LMN, EF_BETA | EF_EXP2)                   // Kronecker apply: F = sv L1 T + mean and exp(F)
  GPINV_GEMM_CASE(LK, LMN, EF_MULMAT)                         // Fbar = X o (R A) c
  GPINV_GEMM_CASE(LK, LMN, EF_BETA | EF_COLSUM)               // Ubar = sv Fbar Lc + klbar U
  GPINV_GEMM_CASE(LMN, LMN, EF_LOWER)                         // Qbar = tril(Ubar^T E)
  GPINV_GEMM_CASE(LMN, LMN, EF_LOWER | EF_BETA | EF_DOT)      // Lbar += tril(Fbar^T U), <.,Lc>
  set_last_error("gemm: no kernel instantiated for this layout / epilogue combination");
  return GPINV_ERR_ARG;
}
}  // namespace

int gemm_launch(int la, int lb, unsigned ef, GemmParams p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0) return GPINV_OK;
  if (p.splitk < 1) p.splitk = 1;
  if (p.splitk > 1 && !(ef & EF_ATOMIC)) {
    set_last_error("gemm: split-K requires the atomic epilogue");
    return GPINV_ERR_ARG;
  }
  p.vecA = aligned16(p.A) && (p.lda % 2 == 0);
  p.vecB = aligned16(p.B) && (p.ldb % 2 == 0);
  p.vecO = (p.out == nullptr) || (aligned16(p.out) && (p.ldo % 2 == 0));
  if (p.batch < 1) p.batch = 1;
  if (p.batch > 1) {
    p.vecA = p.vecA && (p.strideA % 2 == 0);
    p.vecB = p.vecB && (p.strideB % 2 == 0);
    p.vecO = p.vecO && (p.strideO % 2 == 0);
  }
  p.vecO2 = (p.out2 != nullptr) && aligned16(p.out2) && (p.ldo2 % 2 == 0);
  const bool lower = (ef & EF_LOWER) != 0;
  {
    // Blackwell path: TMA-staged operands, mbarrier ring, persistent stream-K partition
    int rc_tma = GPINV_OK;
    GemmParams pt = p;
    if (gemm_tma_try_launch(la, lb, ef, pt, use_small_tma(p, lower), stream, &rc_tma)) return rc_tma;
  }
  const bool small = use_small(p, lower);
  // ---- k-split planning (big tiles only): list-schedule the CTAs in launch order on the SMs
  // with a per-CTA time model and keep the split factor with the smallest makespan.  Fixes wave
  // quantisation (uniform K) and the long-tile imbalance of the triangular products.
  static const bool no_split = (getenv("GPINV_NO_SPLIT") != nullptr);
  if (p.max_ctas > 0) p.allow_split = 0;
  if (p.allow_split && !no_split && p.splitk == 1 && !(ef & EF_ATOMIC) && !small) {
    WsReg* w = ws_for_device();
    const int tm_ = cdiv(p.M, 128), tn_ = cdiv(p.N, 128);
    const int T1 = lower ? tm_ * (tm_ + 1) / 2 : tm_ * tn_;
    const long long T = (long long)T1 * p.batch;
    // the plan only depends on the problem geometry: cache it (the simulation costs ~1 ms of host time)
    struct PlanKey { int M, N, K, kmode, lower, batch; };
    static thread_local PlanKey cache_key[64];
    static thread_local int cache_val[64];
    static thread_local int cache_from[64];
    static thread_local int cache_n = 0;
    int cached = -1, cached_from = 0;
    for (int i = 0; i < cache_n; ++i) {
      const PlanKey& c = cache_key[i];
      if (c.M == p.M && c.N == p.N && c.K == p.K && c.kmode == p.kmode && c.lower == (int)lower && c.batch == p.batch) {
        cached = cache_val[i];
        cached_from = cache_from[i];
        break;
      }
    }
    if (w && cached > 1) {
      p.splitk = cached;
      p.split_from = cached_from;
      p.ws_cnt = static_cast<int*>(w->ptr);
      p.ws = reinterpret_cast<double*>(static_cast<char*>(w->ptr) + CNT_SLOTS * sizeof(int));
    } else if (w && cached < 0 && T <= CNT_SLOTS / 2 && T1 <= 4096) {
      const size_t tile_bytes = 128 * 128 * sizeof(double);
      const size_t cap = (w->bytes > CNT_SLOTS * sizeof(int)) ? (w->bytes - CNT_SLOTS * sizeof(int)) / tile_bytes / 2 : 0;
      static thread_local int kt[4096];
      for (int i = 0; i < T1; ++i) {
        int tm, tn, kb, ke;
        gemm_decode_tile(i, tm_, tn_, p.kmode, lower, tm, tn);
        gemm_krange(p.kmode, p.K, tm * 128, tn * 128, 128, 128, 16, kb, ke);
        kt[i] = (ke > kb) ? cdiv(ke - kb, 16) : 0;
      }
      const int sms = 148;
      auto makespan = [&](int s_, int from) {
        double busy[148];
        for (int i = 0; i < sms; ++i) busy[i] = 0.0;
        // launch order: blockIdx.x fastest, then the slice index, then the batch; tiles before
        // `from` are not split (their extra slice CTAs exit immediately)
        for (int z = 0; z < p.batch; ++z)
          for (int y = 0; y < s_; ++y)
            for (int i = 0; i < T1; ++i) {
              const int si = (i >= from) ? s_ : 1;
              if (y >= si) continue;
              const int per = cdiv(kt[i], si);
              const int mine = (kt[i] - y * per < per ? kt[i] - y * per : per);
              const double t = (mine > 0 ? mine : 0) * 2.1 + 3.0 + (si > 1 ? 6.0 : 0.0);
              int best = 0;
              for (int q = 1; q < sms; ++q) if (busy[q] < busy[best]) best = q;
              busy[best] += t;
            }
        double mx = 0.0;
        for (int i = 0; i < sms; ++i) if (busy[i] > mx) mx = busy[i];
        return mx;
      };
      const double base = makespan(1, 0);
      double best = base;
      int best_s = 1, best_from = 0;
      int kt_max = 0;
      for (int i = 0; i < T1; ++i) if (kt[i] > kt_max) kt_max = kt[i];
      for (int s_ = 2; s_ <= 6; ++s_) {
        if (kt_max / s_ < 12 || (size_t)(T * s_) > cap) break;
        const double c = makespan(s_, 0);
        if (c < 0.97 * best) { best = c; best_s = s_; best_from = 0; }
        // tail-wave split: with uniform tiles only the last (T mod 148) tiles need splitting;
        // same makespan, a fraction of the slice exchange through the workspace
        if (p.kmode == 0 && !lower && p.batch == 1 && T1 > sms && T1 % sms != 0) {
          const int from = (T1 / sms) * sms;
          const double ct = makespan(s_, from);
          if (ct < 0.97 * base && ct <= 1.005 * best) {
            best = (ct < best ? ct : best);
            best_s = s_;
            best_from = from;
          }
        }
      }
      if ((size_t)(T * best_s) > cap) best_s = 1;
      if (cache_n < 64) {
        cache_key[cache_n] = PlanKey{p.M, p.N, p.K, p.kmode, (int)lower, p.batch};
        cache_from[cache_n] = best_from;
        cache_val[cache_n++] = best_s;
      }
      if (best_s > 1) {
        p.splitk = best_s;
        p.split_from = best_from;
        p.ws_cnt = static_cast<int*>(w->ptr);
        p.ws = reinterpret_cast<double*>(static_cast<char*>(w->ptr) + CNT_SLOTS * sizeof(int));
      }
    }
  }
  // ---- k-split for the 64x64 configuration: a few hundred short tiles cannot fill 148 SMs x 3
  // resident CTAs (e.g. the 1024^3 triangular products of the Cholesky reverse leaves: 136 tiles)
  if (p.allow_split && !no_split && p.splitk == 1 && !(ef & EF_ATOMIC) && small) {
    WsReg* w = ws_for_device();
    const int tm_ = cdiv(p.M, 64), tn_ = cdiv(p.N, 64);
    const long long T = (long long)(lower ? tm_ * (tm_ + 1) / 2 : tm_ * tn_) * p.batch;
    const int ktiles = cdiv(p.K, 16);
    const int slots = 148 * 3;
    int s_ = (int)(slots / (T > 0 ? T : 1));
    s_ = min(s_, ktiles / 12);          // >= 12 k-tiles (192 columns of K) per slice
    s_ = min(s_, 4);
    const size_t tile_bytes = 64 * 64 * sizeof(double);
    const size_t cap = (w && w->bytes > CNT_SLOTS * sizeof(int)) ? (w->bytes - CNT_SLOTS * sizeof(int)) / tile_bytes / 2 : 0;
    if (w && s_ >= 2 && T <= CNT_SLOTS / 2 && (size_t)(T * s_) <= cap) {
      p.splitk = s_;
      p.ws_cnt = static_cast<int*>(w->ptr);
      p.ws = reinterpret_cast<double*>(static_cast<char*>(w->ptr) + CNT_SLOTS * sizeof(int));
    }
  }
  if (p.splitk > 1 && p.ws && p.ws_slot == 1) {   // second half of the workspace
    WsReg* w = ws_for_device();
    p.ws_cnt += CNT_SLOTS / 2;
    p.ws += ((w->bytes - CNT_SLOTS * sizeof(int)) / sizeof(double)) / 2;
  }
  static const bool log_gemm = (getenv("GPINV_GEMM_LOG") != nullptr);
  if (log_gemm)
    fprintf(stderr, "gemm-log la=%d lb=%d ef=%u M=%d N=%d K=%d kmode=%d splitk=%d batch=%d cfg=%s from=%d\n", la, lb, ef,
            p.M, p.N, p.K, p.kmode, p.splitk, p.batch, small ? "small" : "big", p.split_from);
  if (small) return dispatch<CfgSmall>(la, lb, ef, p, stream);
  static const int big_bk = getenv("GPINV_BIG_BK") ? atoi(getenv("GPINV_BIG_BK")) : 16;
  if (big_bk == 32 || p.force_cfg == 3) return dispatch<CfgBig32>(la, lb, ef, p, stream);
  return dispatch<CfgBig>(la, lb, ef, p, stream);
}

int gemm_simple(int la, int lb, int M, int N, int K, double alpha, const double* A, int lda,
                const double* B, int ldb, double beta, double* C, int ldc, bool lower, int kmode,
                int splitk, cudaStream_t stream) {
  GemmParams p = {};
  p.A = A; p.B = B; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb;
  p.alpha = alpha; p.beta = beta; p.Cin = C; p.ldcin = ldc; p.out = C; p.ldo = ldc;
  p.kmode = kmode; p.splitk = splitk;
  unsigned ef = (splitk > 1) ? EF_ATOMIC : EF_BETA;
  if (lower) ef |= EF_LOWER;
  if (splitk > 1 && beta != 1.0) {
    set_last_error("gemm_simple: split-K needs beta == 1");
    return GPINV_ERR_ARG;
  }
  return gemm_launch(la, lb, ef, p, stream);
}

int gemm_batched(int la, int lb, int M, int N, int K, double alpha, const double* A, int lda,
                 long long strideA, const double* B, int ldb, long long strideB, double beta,
                 double* C, int ldc, long long strideC, int batch, bool lower, int kmode,
                 cudaStream_t stream, int allow_split, int ws_slot) {
  if (batch <= 0) return GPINV_OK;
  GemmParams p = {};
  p.A = A; p.B = B; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb;
  p.alpha = alpha; p.beta = beta; p.Cin = C; p.ldcin = ldc; p.out = C; p.ldo = ldc;
  p.kmode = kmode; p.splitk = 1;
  p.batch = batch; p.strideA = strideA; p.strideB = strideB; p.strideO = strideC;
  p.strideCin = strideC;
  p.allow_split = allow_split; p.ws_slot = ws_slot;
  unsigned ef = EF_BETA;
  if (lower) ef |= EF_LOWER;
  return gemm_launch(la, lb, ef, p, stream);
}

}  // namespace gpinv
