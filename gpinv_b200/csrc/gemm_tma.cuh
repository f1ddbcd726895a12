// Blackwell-native data movement for the FP64 DMMA GEMM (sm_100a): operands are staged by TMA
// (cp.async.bulk.tensor, 128-byte hardware swizzle, out-of-bounds zero fill) into an mbarrier-tracked
// ring by a dedicated producer warp; the DMMA warps never issue a load or take part in a CTA-wide
// barrier: they wait on a stage's "full" mbarrier, read their fragments with conflict-free LDS.64,
// and release the stage with one arrive per warp on its "empty" mbarrier.
//
// The kernel is PERSISTENT with a stream-K work partition: the host cuts the ordered list of
// (tile, k-tile) work units into one contiguous range per CTA (equal work, so the triangular
// products -- whose tiles differ 32x in length -- and the last partial wave of uniform problems are
// balanced).  A tile whose k range is shared by several CTAs is combined through the registered
// workspace: every contributor stores its partial accumulators, takes a ticket, and the LAST
// arriver sums the slices in fixed CTA order (bitwise reproducible) and runs the fused epilogue.
// No CTA ever waits on another.  Because the producer warp walks the same range independently, it
// prefetches the next tile's first stages while the DMMA warps are still in the epilogue.
//
// tcgen05 / TMEM have no f64 kind, so the math stays mma.sync m8n8k4 (SASS DMMA.8x8x4); what is
// Blackwell/Hopper-native here is the data path: SASS shows UTMALDG + SYNCS (mbarrier) and no
// LDGSTS / BAR in the main loop.
//
// Shared-memory layouts (both dense, 16 KB per 128-row operand stage):
//   LK  operand (k contiguous):   one TMA box [rows][16 k]; row pitch 128 B; element (r,k) at
//        r*128 + (((k>>1) ^ (r&7)) << 4) + (k&1)*8                       (128B swizzle, key = row)
//   LMN operand (row index contiguous): rows/16 boxes [16 k][16 rows] of 2 KB; element (r,k) at
//        (r>>4)*2048 + k*128 + ((((r&15)>>1) ^ (k&7)) << 4) + (r&1)*8    (128B swizzle, key = k)
// Fragment ownership is permuted so that every LDS.64 of a half-warp hits 16 distinct bank pairs:
// within each group of 16 rows two DMMA tiles (i even / odd) are interleaved,
//   LK : row = 2 g + (i&1)                       LMN: row = (g&1) + 2 (i&1) + 4 (g>>2) + 8 ((g>>1)&1)
// (lane = 4 g + t, k = 4 kk + t).  The epilogue uses the same maps for the output coordinates.
#pragma once
#include <cuda.h>

#include "gemm.cuh"

namespace gpinv {

constexpr int TMA_MAX_CTAS = 448;

// Work partition of one launch: CTA c owns the units [start[c], start[c+1]) of the ordered tile
// list; a unit position is (global tile index << 16) | k-tile offset inside the tile.
// `perm` (use_perm != 0; uniform, non-triangular, un-batched products with <= TMA_MAX_PERM tiles): list
// position -> tile index of the problem.  The list is ordered so that the positions the CTAs work on AT THE
// SAME TIME (same offset from the start of their ranges) are neighbours in the tile grid -- one wave of 148
// CTAs then covers all row panels of ~18 column panels, every operand panel is fetched from DRAM once and
// shared through L2.  With the plain order (CTA c owns tiles 3.46 c ...) each CTA walked the row panels of
// "its" column panel one after the other and the column panels were re-read from DRAM for every row panel:
// 2.49 GB per Abel-transform launch against 0.37 GB algorithmic (ncu, profiles/r2_abel_gemm_ncu_full.txt).
constexpr int TMA_MAX_PERM = 1024;
struct TmaSched {
  int nctas;
  int total_tiles;
  int use_perm;
  unsigned start[TMA_MAX_CTAS + 1];
  unsigned short perm[TMA_MAX_PERM];
};

template <int TM_, int TN_, int WM_, int WN_, int STAGES_, int MINB_, int REG_CONS_, int REG_PROD_>
struct TmaCfg {
  static constexpr int TM = TM_, TN = TN_, WM = WM_, WN = WN_;
  static constexpr int BM = WM * TM * 8, BN = WN * TN * 8, BK = 16;
  static constexpr int CONSUMERS = WM * WN * 32;
  // + one producer WARPGROUP (4 warps, one of which issues the TMA loads): registers are allocated
  // per 4 warps, and setmaxnreg -- which moves the producer's registers to the DMMA warps, whose
  // 64x32 accumulator tile alone is 128 registers -- works on whole warpgroups
  static constexpr int THREADS = CONSUMERS + 128;
  static constexpr int REG_CONS = REG_CONS_, REG_PROD = REG_PROD_;
  static constexpr int STAGES = STAGES_;
  static constexpr int MINB = MINB_;
  static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // [stages][A|B]  [full barriers][empty barriers]  [ticket]   (+1 KB slack for the 1024-byte alignment)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16 + 1024;
};
// 128x128 tile, 8 DMMA warps (warp tile 64x32) + producer warpgroup, 6 stages x 32 KB, 1 CTA/SM:
// 384 threads start with 168 registers; the DMMA warps take 232, the producer warpgroup keeps 40
using TmaBig = TmaCfg<8, 4, 2, 4, 6, 1, 232, 40>;
// 64x64 tile, 4 DMMA warps (warp tile 32x32) + producer warpgroup, 4 stages x 16 KB, 3 CTAs/SM.
// Not instantiated: small / short-K problems (a few k-tiles per tile) gain nothing from the ring and
// measured slower than the cp.async kernels of gemm.cuh (1024^3: 9.9 vs 16.9 TF/s), which keep them.
using TmaSmall = TmaCfg<4, 4, 2, 2, 4, 3, 136, 24>;

// ------------------------------------ PTX wrappers ------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Wait for the phase with the given parity.  A wait that lasts ~2 s is a protocol bug: trap (the
// launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (done) return;
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// 3-D tiled TMA load (coordinates innermost first) completing on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// one tile's share of work for this CTA
struct TmaWork {
  int gtile;        // global tile index (batch * ntiles + tile)
  int tile_x, z;    // tile inside the problem, batch index
  int m0, n0;
  int kb;           // first k of the tile's k range
  int ke;           // k limit (exclusive) of the tile's k range
  int ktiles;       // k-tiles of the whole tile
  int k_first, k_last;   // this CTA's k-tiles [k_first, k_last)
};

template <class Cfg, unsigned EF>
__device__ __forceinline__ void tma_decode(const GemmParams& p, const TmaSched& sched, int gtile, TmaWork& w) {
  w.gtile = gtile;
  w.z = gtile / p.ntiles;
  w.tile_x = gtile - w.z * p.ntiles;
  if (sched.use_perm) w.tile_x = sched.perm[gtile];      // (un-batched: gtile == list position)
  int tm, tn;
  gemm_decode_tile(w.tile_x, p.tiles_m, p.tiles_n, p.kmode, (EF & EF_LOWER) != 0, tm, tn);
  w.m0 = tm * Cfg::BM;
  w.n0 = tn * Cfg::BN;
  gemm_krange(p.kmode, p.K, w.m0, w.n0, Cfg::BM, Cfg::BN, Cfg::BK, w.kb, w.ke);
  w.ktiles = (w.ke > w.kb) ? (w.ke - w.kb + Cfg::BK - 1) / Cfg::BK : 0;
}

// Output coordinates of accumulator element (i, j, e) of this lane (see the header comment).
template <int LA>
__device__ __forceinline__ int tma_row_of(int i, int g) {
  if (LA == LK) return (i >> 1) * 16 + 2 * g + (i & 1);
  return (i >> 1) * 16 + (g & 1) + 2 * (i & 1) + 4 * (g >> 2) + 8 * ((g >> 1) & 1);
}

template <class Cfg, int LA, int LB, unsigned EF>
__device__ __forceinline__ void tma_epilogue(const GemmParams& p, const TmaWork& w,
                                             double (&acc)[Cfg::TM][Cfg::TN][2], int warp, int lane) {
  constexpr int TM = Cfg::TM, TN = Cfg::TN;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
  const int m0 = w.m0, n0 = w.n0;
  const double* Cin = p.Cin;
  double* out = p.out;
  double* out2 = p.out2;
  if (p.batch > 1) {
    if (out) out += (long long)w.z * p.strideO;
    if (out2) out2 += (long long)w.z * p.strideO;
    if (Cin) Cin += (long long)w.z * p.strideCin;
  }
  double alpha = p.alpha;
  if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
  double beta = 0.0;
  if (EF & EF_BETA) {
    beta = p.beta;
    if (p.beta_dev) beta *= __ldg(p.beta_dev);
  }
  double r_sumsq = 0.0, r_dot = 0.0;
  double r_col[TN][2];
#pragma unroll
  for (int P = 0; P < TN; ++P) r_col[P][0] = r_col[P][1] = 0.0;

  // "Pair" P of a row holds two ADJACENT output columns (col0, col0+1):
  //   LB == LMN: the two elements e = 0,1 of DMMA tile j = P
  //   LB == LK : element e = P&1 of the two interleaved DMMA tiles j = 2 (P>>1) and 2 (P>>1) + 1
#define GPINV_TMA_ACC(i_, P_, q_) ((LB == LK) ? acc[i_][2 * ((P_) >> 1) + (q_)][(P_)&1] : acc[i_][P_][q_])
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + wm * TM * 8 + tma_row_of<LA>(i, g);
    bool ok[TN][2], upper[TN][2];
    double cin[TN][2], mulv[TN][2], wv[TN][2];
    int col0s[TN];
#pragma unroll
    for (int P = 0; P < TN; ++P) {
      const int col0 = n0 + wn * TN * 8 +
                       ((LB == LK) ? ((P >> 1) * 16 + 4 * t + 2 * (P & 1))
                                   : ((P >> 1) * 16 + 2 * (P & 1) + 4 * (t >> 1) + 8 * (t & 1)));
      col0s[P] = col0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = col0 + q;
        bool o = (row < p.M) && (c < p.N);
        upper[P][q] = false;
        if (EF & EF_LOWER) {
          upper[P][q] = o && (c > row);
          o = o && (c <= row);
        }
        ok[P][q] = o;
        cin[P][q] = 0.0; mulv[P][q] = 1.0; wv[P][q] = 0.0;
        if (o) {
          if ((EF & EF_BETA) && beta != 0.0) cin[P][q] = __ldcg(Cin + (size_t)row * p.ldcin + c);
          if (EF & EF_MULMAT) mulv[P][q] = __ldg(p.mul + (size_t)row * p.ldmul + c);
          if (EF & EF_DOT) wv[P][q] = __ldg(p.W + (size_t)row * p.ldw + c);
        }
      }
    }
#pragma unroll
    for (int P = 0; P < TN; ++P) {
      const int col = col0s[P];
      double v[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = col + q;
        double x = alpha * GPINV_TMA_ACC(i, P, q);
        if (ok[P][q]) {
          if (EF & EF_DOT) r_dot += x * wv[P][q];
          if (EF & EF_MULMAT) x *= mulv[P][q];
          if (EF & EF_BETA) x += beta * cin[P][q];
          if (EF & EF_COLVEC) x += __ldg(p.colvec + c);
          if (EF & EF_SUMSQ) r_sumsq += x * x;
          if (EF & EF_COLSUM) r_col[P][q] += x;
        }
        v[q] = x;
        if ((EF & EF_LOWER) && !(EF & EF_ATOMIC) && upper[P][q] && out) {
          v[q] = 0.0;
          out[(size_t)row * p.ldo + c] = 0.0;
        }
      }
      if (EF & EF_ATOMIC) {
        if (ok[P][0]) atomicAdd(out + (size_t)row * p.ldo + col, v[0]);
        if (ok[P][1]) atomicAdd(out + (size_t)row * p.ldo + col + 1, v[1]);
      } else {
        if (out) {
          double* o = out + (size_t)row * p.ldo + col;
          if (ok[P][0] && ok[P][1] && p.vecO) {
            *reinterpret_cast<double2*>(o) = make_double2(v[0], v[1]);
          } else {
            if (ok[P][0]) o[0] = v[0];
            if (ok[P][1]) o[1] = v[1];
          }
        }
        if (EF & EF_EXP2) {
          double* o = out2 + (size_t)row * p.ldo2 + col;
          if (ok[P][0] && ok[P][1] && p.vecO2) {
            *reinterpret_cast<double2*>(o) = make_double2(exp(v[0]), exp(v[1]));
          } else {
            if (ok[P][0]) o[0] = exp(v[0]);
            if (ok[P][1]) o[1] = exp(v[1]);
          }
        }
      }
    }
  }
#undef GPINV_TMA_ACC
  if (EF & EF_SUMSQ) {
    r_sumsq = warp_sum(r_sumsq);
    if (lane == 0) atomicAdd(p.sumsq, r_sumsq);
  }
  if (EF & EF_DOT) {
    r_dot = warp_sum(r_dot);
    if (lane == 0) atomicAdd(p.dot, r_dot);
  }
  if (EF & EF_COLSUM) {
#pragma unroll
    for (int P = 0; P < TN; ++P)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double s = r_col[P][q];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        const int c = n0 + wn * TN * 8 + q +
                      ((LB == LK) ? ((P >> 1) * 16 + 4 * t + 2 * (P & 1))
                                  : ((P >> 1) * 16 + 2 * (P & 1) + 4 * (t >> 1) + 8 * (t & 1)));
        if (g == 0 && c < p.N) atomicAdd(p.colsum + c, s);
      }
  }
}

template <class Cfg, int LA, int LB, unsigned EF>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const GemmParams p, const __grid_constant__ TmaSched sched) {
  constexpr int TM = Cfg::TM, TN = Cfg::TN, BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
  constexpr int STAGES = Cfg::STAGES, NCONS = Cfg::CONSUMERS;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: the 128B swizzle pattern is a function of the shared-memory address
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem0 + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  int* s_ticket = reinterpret_cast<int*>(smem_raw + (bar_empty + STAGES * 8 - smem_u32(smem_raw)));

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);                 // one arrive.expect_tx by the producer
      mbar_init(bar_empty + 8 * s, NCONS / 32);       // one arrive per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int cta = blockIdx.x;
  const unsigned u_begin = sched.start[cta], u_end = sched.start[cta + 1];
  const int t_begin = (int)(u_begin >> 16), k_begin = (int)(u_begin & 0xffffu);
  const int t_end = (int)(u_end >> 16), k_end = (int)(u_end & 0xffffu);
  // tiles touched: t_begin .. t_end (t_end only when k_end > 0)
  const int t_stop = (k_end > 0) ? t_end + 1 : t_end;
  // Processing order (uniform products, use_perm): a range that starts in the middle of a tile takes that
  // partial piece LAST.  All CTAs then start their whole tiles together and sweep k in lock-step, so the
  // CTAs that share an operand panel ask L2 for the same k-slice at the same time; the partial pieces, whose
  // k offsets are staggered by construction, come at the end.
  const int t_cnt = t_stop - t_begin;
  const bool rotate = sched.use_perm && k_begin > 0 && t_cnt > 1;

  if (warp >= NCONS / 32) {
    // =============================== TMA producer ===============================
    if (Cfg::REG_PROD > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REG_PROD > 0 ? Cfg::REG_PROD : 24));
    if (warp == NCONS / 32 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = 0; ti < t_cnt; ++ti) {
        const int gt = rotate ? ((ti + 1 < t_cnt) ? t_begin + 1 + ti : t_begin) : t_begin + ti;
        TmaWork w;
        tma_decode<Cfg, EF>(p, sched, gt, w);
        const int kf = (gt == t_begin) ? k_begin : 0;
        const int kl = (gt == t_end) ? k_end : w.ktiles;
        const int zA = p.strideA ? w.z : 0, zB = p.strideB ? w.z : 0;
        for (int kt = kf; kt < kl; ++kt) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, Cfg::STAGE_BYTES);
          const uint32_t sA = smem0 + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          const int k0 = w.kb + kt * BK;
          if (LA == LK) {
            tma_load_3d(sA, &mapA, k0, w.m0, zA, full);
          } else {
#pragma unroll
            for (int q = 0; q < BM / 16; ++q) tma_load_3d(sA + q * 2048, &mapA, w.m0 + 16 * q, k0, zA, full);
          }
          if (LB == LK) {
            tma_load_3d(sB, &mapB, k0, w.n0, zB, full);
          } else {
#pragma unroll
            for (int q = 0; q < BN / 16; ++q) tma_load_3d(sB + q * 2048, &mapB, w.n0 + 16 * q, k0, zB, full);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // ================================= DMMA consumers =================================
  if (Cfg::REG_CONS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REG_CONS > 0 ? Cfg::REG_CONS : 24));
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
  // per-lane fragment address pieces (bytes inside a stage); see the header comment
  uint32_t a_lane, b_lane;      // lane-constant part
  uint32_t a_x, b_x;            // swizzle key material
  if (LA == LK) {
    a_lane = (wm * TM * 8) * 128 + g * 256 + (t & 1) * 8;
    a_x = ((g & 3) << 1) | (t >> 1);
  } else {
    a_lane = (wm * TM / 2) * 2048 + t * 128 + (g & 1) * 8;
    a_x = ((((g >> 2) & 1) << 1) | (((g >> 1) & 1) << 2)) ^ t;
  }
  if (LB == LK) {
    b_lane = (wn * TN * 8) * 128 + g * 256 + (t & 1) * 8;
    b_x = ((g & 3) << 1) | (t >> 1);
  } else {
    b_lane = (wn * TN / 2) * 2048 + t * 128 + (g & 1) * 8;
    b_x = ((((g >> 2) & 1) << 1) | (((g >> 1) & 1) << 2)) ^ t;
  }

  int stage = 0;
  uint32_t phase = 0;
  for (int ti = 0; ti < t_cnt; ++ti) {
    const int gt = rotate ? ((ti + 1 < t_cnt) ? t_begin + 1 + ti : t_begin) : t_begin + ti;
    TmaWork w;
    tma_decode<Cfg, EF>(p, sched, gt, w);
    w.k_first = (gt == t_begin) ? k_begin : 0;
    w.k_last = (gt == t_end) ? k_end : w.ktiles;

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = w.k_first; kt < w.k_last; ++kt) {
      mbar_wait(bar_full + 8 * stage, phase);
      const uint32_t sA = smem0 + stage * Cfg::STAGE_BYTES + a_lane;
      const uint32_t sB = smem0 + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES + b_lane;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[TM], b[TN];
        // even / odd DMMA tile of each 16-row group
        uint32_t a0, a1, b0, b1;
        if (LA == LK) {
          a0 = sA + (((2 * kk) ^ a_x) << 4);
          a1 = sA + ((((2 * kk) ^ a_x) ^ 1) << 4) + 128;
        } else {
          a0 = sA + kk * 512 + ((a_x ^ ((kk & 1) << 2)) << 4);
          a1 = sA + kk * 512 + (((a_x ^ ((kk & 1) << 2)) ^ 1) << 4);
        }
        if (LB == LK) {
          b0 = sB + (((2 * kk) ^ b_x) << 4);
          b1 = sB + ((((2 * kk) ^ b_x) ^ 1) << 4) + 128;
        } else {
          b0 = sB + kk * 512 + ((b_x ^ ((kk & 1) << 2)) << 4);
          b1 = sB + kk * 512 + (((b_x ^ ((kk & 1) << 2)) ^ 1) << 4);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = lds_f64(((i & 1) ? a1 : a0) + (i >> 1) * 2048);
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = lds_f64(((j & 1) ? b1 : b0) + (j >> 1) * 2048);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      // Release the stage only after this warp's fragment loads have been PERFORMED.  ptxas schedules
      // the arrive right behind the ISSUE of the last LDS (it has no register dependence on the loaded
      // values), and on B200 the SYNCS arrive can take effect while that LDS is still queued: the
      // producer, blocked on this barrier, then refills the stage under the load and the tile gets a
      // wrong partial sum -- sporadically (tests/gpu_gemm_chain.py: up to 39 of 40 repetitions at ragged
      // sizes, 1 of 59 at aligned ones).  A CTA-scope fence orders the loads before the arrive; it
      // costs ~40 cycles per 4096-cycle k-tile.
      __threadfence_block();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }

    const bool whole = (w.k_first == 0 && w.k_last == w.ktiles);
    if ((EF & EF_ATOMIC) || whole) {
      if (!(EF & EF_ATOMIC) || w.k_last > w.k_first || w.ktiles == 0)
        tma_epilogue<Cfg, LA, LB, EF>(p, w, acc, warp, lane);
      continue;
    }
    // ---- this tile's k range is shared with other CTAs: exchange through the workspace ----
    int c_first = cta, c_last = cta;
    while (c_first > 0 && sched.start[c_first] > ((unsigned)gt << 16)) --c_first;
    while (c_last + 1 < sched.nctas && sched.start[c_last + 1] < ((unsigned)(gt + 1) << 16)) ++c_last;
    const int nsl = c_last - c_first + 1;
    constexpr int TILE = BM * BN;
    double* mine = p.ws + (size_t)(2 * cta + ((t_begin == gt) ? 0 : 1)) * TILE;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) __stcg(mine + ((i * TN + j) * 2 + e) * NCONS + tid, acc[i][j][e]);
    __threadfence();
    consumer_bar_sync(NCONS);
    if (tid == 0) *s_ticket = atomicAdd(p.ws_cnt + c_first, 1);
    consumer_bar_sync(NCONS);
    const int ticket = *s_ticket;
    consumer_bar_sync(NCONS);              // everyone has read the ticket before it is reused
    if (ticket != nsl - 1) continue;
    if (tid == 0) p.ws_cnt[c_first] = 0;   // ready for the next launch (stream ordered)
    __threadfence();
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int c = c_first; c <= c_last; ++c) {   // fixed order: bitwise reproducible
      const int slot = ((int)(sched.start[c] >> 16) == gt) ? 0 : 1;
      const double* src = p.ws + (size_t)(2 * c + slot) * TILE + tid;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        double tmp[TN][2];
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) tmp[j][e] = __ldcg(src + ((i * TN + j) * 2 + e) * NCONS);
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) acc[i][j][e] += tmp[j][e];
      }
    }
    tma_epilogue<Cfg, LA, LB, EF>(p, w, acc, warp, lane);
  }
}

// Host side (gemm.cu): true when the TMA path was taken; rc receives the launch status.
bool gemm_tma_try_launch(int la, int lb, unsigned ef, GemmParams& p, bool small, cudaStream_t stream, int* rc);

}  // namespace gpinv
