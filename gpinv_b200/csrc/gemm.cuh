// FP64 tensor-core (DMMA.8x8x4) GEMM with fused epilogues, triangular k-range skipping,
// lower-triangular output and split-K.  One kernel template serves every GEMM-shaped stage
// of the StVGP/GPMC hot path (sample products T10/T12, Abel transform T15, their reverse
// products, and the trailing updates of the blocked Cholesky T4 and its reverse sweep).
//
//   out(r,c) = epilogue( alpha * sum_k A(r,k) * B(c,k) ),  r < M, c < N, k in [k_begin,k_end)
//
// Operand layouts:  LK  : element (r,k) at p[r*ld + k]   (k contiguous)
//                   LMN : element (r,k) at p[k*ld + r]   (row index contiguous)
// so BLAS "C = A B^T" is <LK,LK>, "C = A B" is <LK,LMN>, "C = A^T B" is <LMN,LMN>.
//
// Two tile configurations: CfgBig = CTA tile 128x128x16, 8 warps (2x4), warp tile 64x32 = 8x4 DMMA
// tiles, accumulators in registers (128 regs), 4-stage cp.async (LDGSTS) pipeline, 1 CTA/SM;
// CfgSmall = 64x64x16, 4 warps, 3 stages, 3 CTAs/SM for small or short-K problems where one
// 128x128 tile is a whole SM's worth of serial work.  Padded conflict-free smem in both.
#pragma once
#include "common.cuh"

namespace gpinv {

enum Layout { LK = 0, LMN = 1 };

enum : unsigned {
  EF_BETA = 1u,     // v += beta * Cin(r,c)
  EF_COLVEC = 2u,   // v += colvec[c]
  EF_MULMAT = 4u,   // v *= Mul(r,c)
  EF_EXP2 = 8u,     // out2(r,c) = exp(v)
  EF_SUMSQ = 16u,   // *sumsq += sum v^2
  EF_COLSUM = 32u,  // colsum[c] += sum_r v
  EF_DOT = 64u,     // *dot += sum (alpha*acc) * W(r,c)
  EF_LOWER = 128u,  // square output, only c <= r computed / stored
  EF_ATOMIC = 256u  // split-K: out(r,c) += alpha*acc via red.add.f64
};

struct GemmParams {
  const double* A;
  const double* B;
  int M, N, K;
  int lda, ldb;
  double alpha;
  const double* alpha_dev;  // optional device scalar multiplied into alpha
  double beta;
  const double* beta_dev;  // optional device scalar multiplied into beta
  const double* Cin;
  int ldcin;
  double* out;  // may be null (then only out2 / reductions are produced)
  int ldo;
  double* out2;
  int ldo2;
  const double* colvec;
  const double* mul;
  int ldmul;
  const double* W;
  int ldw;
  double* sumsq;
  double* colsum;
  double* dot;
  int kmode;   // 0 full | 1: k < n0+BN | 2: k >= n0 | 3: k < m0+BM | 4: k >= m0 | 5: n0 <= k < m0+BM
  int splitk;  // >= 1
  int tiles_m, tiles_n;
  int force_cfg;  // 0 auto | 1 big (128x128) | 2 small (64x64) | 3 big with BK=32
  // Workspace split-K (wave-quantisation fix): when splitk > 1 and the epilogue is not the
  // atomic one, every k-slice CTA stores its partial tile to `ws`, takes a ticket on the tile's
  // counter, and the LAST arriver sums the slices in fixed order (deterministic) and runs the
  // fused epilogue.  No CTA ever waits on another.  Enabled per call with allow_split.
  int allow_split;
  int split_from;   // k-split applies to tiles with index >= split_from only (tail-wave split)
  double* ws;
  int* ws_cnt;
  int vecA, vecB, vecO;  // 16-byte alignment of operands / output rows
  int vecO2;             // ... of the second output (EF_EXP2)
  // batching over blockIdx.z (element strides; 0 = shared operand)
  int batch;
  long long strideA, strideB, strideO, strideCin;
  // Persistent mode (max_ctas > 0, splitk == 1, batch == 1): the grid is capped at max_ctas and
  // every CTA walks the tile list with stride gridDim.x.  Used for GEMMs that run BESIDE a
  // latency-critical kernel chain on another stream (the look-ahead update of the Cholesky, the
  // first sampling product): they must leave SMs free instead of flooding the machine.
  int max_ctas;
  int ntiles;   // filled by the launcher
  // The registered split-K workspace is used in two halves so that two k-split GEMMs may run
  // concurrently on two streams (Cholesky reverse sweep): ws_slot 0 / 1 selects the half.
  int ws_slot;
};

template <int TM_, int TN_, int WM_, int WN_, int STAGES_, int MINB_, int BK_ = 16>
struct GemmCfg {
  static constexpr int TM = TM_, TN = TN_, WM = WM_, WN = WN_;   // DMMA tiles per warp, warps per CTA
  static constexpr int BM = WM * TM * 8, BN = WN * TN * 8, BK = BK_;
  static constexpr int THREADS = WM * WN * 32;
  static constexpr int STAGES = STAGES_;
  static constexpr int MINB = MINB_;
  static constexpr int LDK = BK + 4;                 // LK  tile: [rows][20]
  static constexpr int LDMA = BM + 4, LDMB = BN + 4; // LMN tile: [16][rows+4]
  static constexpr int A_ELEMS = (BM * LDK > BK * LDMA) ? BM * LDK : BK * LDMA;
  static constexpr int B_ELEMS = (BN * LDK > BK * LDMB) ? BN * LDK : BK * LDMB;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_ELEMS * 8;
};
// 128x128 tile, 8 warps, warp tile 64x32, 4 stages, 160 KB smem, 1 CTA/SM: the throughput config
using CfgBig = GemmCfg<8, 4, 2, 4, 4, 1>;
// same tile with BK = 32 and 3 stages (216 KB smem): half as many barriers per flop
using CfgBig32 = GemmCfg<8, 4, 2, 4, 3, 1, 32>;
// (A 16-warp variant of the 128x128 tile -- warp tile 32x32, 4 warps per scheduler -- was measured
// and is 0.5-4 % slower on every shape of the evaluation: GemmCfg<4, 4, 4, 4, 4, 1>.)
// 64x64 tile, 4 warps, warp tile 32x32, 3 stages, 60 KB smem, 3 CTAs/SM: small / short-K problems
using CfgSmall = GemmCfg<4, 4, 2, 2, 3, 3>;

// Tile decode (heavy tiles first so the hardware block scheduler balances the tail) and k-range
// of a tile.  Shared by the kernel and by the host-side split planner, which must agree.
__host__ __device__ __forceinline__ void gemm_decode_tile(int idx, int tiles_m, int tiles_n, int kmode,
                                                          bool lower, int& tm, int& tn) {
  if (lower && kmode == 5) {
    // work grows with the distance from the diagonal: enumerate far diagonals first
    int d = tiles_m - 1, rem = idx;
    while (rem >= tiles_m - d) { rem -= tiles_m - d; --d; }
    tn = rem;
    tm = rem + d;
  } else if (lower) {
    int r = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= idx) ++r;
    while (r * (r + 1) / 2 > idx) --r;
    tm = r;
    tn = idx - r * (r + 1) / 2;
  } else if (kmode == 1) {
    tn = tiles_n - 1 - idx / tiles_m;
    tm = idx % tiles_m;
  } else if (kmode == 3) {
    tm = tiles_m - 1 - idx / tiles_n;
    tn = idx % tiles_n;
  } else if (kmode == 4) {
    tm = idx / tiles_n;
    tn = idx % tiles_n;
  } else {
    tn = idx / tiles_m;
    tm = idx % tiles_m;
  }
}

__host__ __device__ __forceinline__ void gemm_krange(int kmode, int K, int m0, int n0, int BM, int BN,
                                                     int BK, int& kb, int& ke) {
  kb = 0;
  ke = K;
  if (kmode == 1) ke = (K < n0 + BN) ? K : n0 + BN;
  else if (kmode == 2) kb = (n0 / BK) * BK;
  else if (kmode == 3) ke = (K < m0 + BM) ? K : m0 + BM;
  else if (kmode == 4) kb = (m0 / BK) * BK;
  else if (kmode == 5) { kb = (n0 / BK) * BK; ke = (K < m0 + BM) ? K : m0 + BM; }
}

template <int LAYOUT, int ROWS, int THREADS, int BK>
__device__ __forceinline__ void gemm_load_tile(double* s, const double* __restrict__ g, int ld,
                                               int r0, int k0, int Rmax, int Kmax, int vec,
                                               int tid) {
  constexpr int LDK = BK + 4;
  constexpr int LDM = ROWS + 4;
  if (LAYOUT == LK) {
    if (vec) {
#pragma unroll
      for (int c = 0; c < ROWS * (BK / 2) / THREADS; ++c) {
        const int id = tid + c * THREADS;
        const int row = id / (BK / 2), kc = (id % (BK / 2)) * 2;
        const int gr = r0 + row, gk = k0 + kc;
        int valid = (gr < Rmax) ? min(max(Kmax - gk, 0), 2) : 0;
        const double* src = valid ? g + (size_t)gr * ld + gk : g;
        cp_async16(s + row * LDK + kc, src, valid * 8);
      }
    } else {
#pragma unroll
      for (int c = 0; c < ROWS * BK / THREADS; ++c) {
        const int id = tid + c * THREADS;
        const int row = id / BK, k = id % BK;
        const int gr = r0 + row, gk = k0 + k;
        const bool ok = (gr < Rmax) && (gk < Kmax);
        const double* src = ok ? g + (size_t)gr * ld + gk : g;
        cp_async8(s + row * LDK + k, src, ok ? 8 : 0);
      }
    }
  } else {
    if (vec) {
#pragma unroll
      for (int c = 0; c < (ROWS / 2) * BK / THREADS; ++c) {
        const int id = tid + c * THREADS;
        const int k = id / (ROWS / 2), rc = (id % (ROWS / 2)) * 2;
        const int gr = r0 + rc, gk = k0 + k;
        int valid = (gk < Kmax) ? min(max(Rmax - gr, 0), 2) : 0;
        const double* src = valid ? g + (size_t)gk * ld + gr : g;
        cp_async16(s + k * LDM + rc, src, valid * 8);
      }
    } else {
#pragma unroll
      for (int c = 0; c < ROWS * BK / THREADS; ++c) {
        const int id = tid + c * THREADS;
        const int k = id / ROWS, r = id % ROWS;
        const int gr = r0 + r, gk = k0 + k;
        const bool ok = (gr < Rmax) && (gk < Kmax);
        const double* src = ok ? g + (size_t)gk * ld + gr : g;
        cp_async8(s + k * LDM + r, src, ok ? 8 : 0);
      }
    }
  }
}

// Per-thread precomputed state for the 16-byte cp.async path: pointers, shared-memory offsets and
// row validity are fixed for the whole k loop; per k-tile only the k offset and the k-tail clamp
// change.  (The 8-byte fallback for unaligned operands keeps using gemm_load_tile.)
template <int LAYOUT, int ROWS, int THREADS, int BK>
struct TileLoader {
  static constexpr int CPR = BK / 2;                        // 16-byte chunks per k-contiguous row
  static constexpr int NCH = ROWS * CPR / THREADS;          // chunks per thread per tile
  static constexpr int LDK = BK + 4, LDM = ROWS + 4;
  const double* ptr;    // first chunk of this thread at k = 0
  long long step;       // element stride between this thread's consecutive chunks
  int soff, sstep;      // shared-memory offset of the first chunk / stride between chunks
  int kfix;             // LK: k offset inside the tile (kc); LMN: first k row (k)
  int rbytes;           // LMN: valid bytes along the row index (0, 8, 16)
  unsigned rowmask;     // LK: bit c set when chunk c's row is inside the matrix
  int ld;

  __device__ __forceinline__ void init(const double* g, int ld_, int r0, int Rmax, int tid) {
    ld = ld_;
    if (LAYOUT == LK) {
      const int row = tid / CPR, kc = (tid % CPR) * 2;
      constexpr int RPP = THREADS / CPR;                   // rows covered per pass
      ptr = g + (size_t)(r0 + row) * ld + kc;
      step = (long long)RPP * ld;
      soff = row * LDK + kc;
      sstep = RPP * LDK;
      kfix = kc;
      rowmask = 0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) rowmask |= (r0 + row + c * RPP < Rmax) ? (1u << c) : 0u;
      rbytes = 0;
    } else {
      constexpr int HALF = ROWS / 2;
      constexpr int KPP = THREADS / HALF;                  // k rows covered per pass
      const int k = tid / HALF, rc = (tid % HALF) * 2;
      ptr = g + (size_t)k * ld + r0 + rc;
      step = (long long)KPP * ld;
      soff = k * LDM + rc;
      sstep = KPP * LDM;
      kfix = k;
      rbytes = min(max(Rmax - (r0 + rc), 0), 2) * 8;
      rowmask = 0;
    }
  }

  // copy the tile whose first k index is k0 (global k limit ke) into stage buffer s
  __device__ __forceinline__ void load(double* s, const double* g, int k0, int ke) const {
    if (LAYOUT == LK) {
      const int kbytes = min(max(ke - (k0 + kfix), 0), 2) * 8;
      const double* p0 = ptr + k0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int bytes = ((rowmask >> c) & 1u) ? kbytes : 0;
        cp_async16(s + soff + c * sstep, bytes ? p0 + c * step : g, bytes);
      }
    } else {
      constexpr int KPP = THREADS / (ROWS / 2);
      const double* p0 = ptr + (size_t)k0 * ld;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int bytes = (k0 + kfix + c * KPP < ke) ? rbytes : 0;
        cp_async16(s + soff + c * sstep, bytes ? p0 + c * step : g, bytes);
      }
    }
  }
};

template <class Cfg, int LA, int LB, unsigned EF>
__device__ __forceinline__ void gemm_tile_body(const GemmParams& pin, const int tile_x, double* smem) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, LDK = Cfg::LDK, LDMA = Cfg::LDMA, LDMB = Cfg::LDMB;
  constexpr int A_ELEMS = Cfg::A_ELEMS, STAGE_ELEMS = Cfg::STAGE_ELEMS;
  GemmParams p = pin;
  if (p.batch > 1) {
    const long long z = blockIdx.z;
    p.A += z * p.strideA;
    p.B += z * p.strideB;
    if (p.out) p.out += z * p.strideO;
    if (p.out2) p.out2 += z * p.strideO;      // the second output follows the first through the batch
    if (p.Cin) p.Cin += z * p.strideCin;
  }

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;

  // ---- tile decode and k range ----
  int tm, tn;
  gemm_decode_tile(tile_x, p.tiles_m, p.tiles_n, p.kmode, (EF & EF_LOWER) != 0, tm, tn);
  const int m0 = tm * BM, n0 = tn * BN;
  int kb, ke;
  gemm_krange(p.kmode, p.K, m0, n0, BM, BN, BK, kb, ke);
  int ktiles = (ke > kb) ? (ke - kb + BK - 1) / BK : 0;
  // slices of THIS tile: tiles before split_from run unsplit (their extra slice CTAs exit at once)
  const int nsl = (p.splitk > 1 && tile_x >= p.split_from) ? p.splitk : 1;
  if ((int)blockIdx.y >= nsl) return;
  if (nsl > 1) {
    const int per = (ktiles + nsl - 1) / nsl;
    const int first = blockIdx.y * per;
    const int last = min(ktiles, first + per);
    kb += first * BK;
    ktiles = max(last - first, 0);
    if (ktiles == 0 && (EF & EF_ATOMIC)) return;
  }

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // ---- pipeline prologue ----
  TileLoader<LA, BM, THREADS, BK> ldA;
  TileLoader<LB, BN, THREADS, BK> ldB;
  ldA.init(p.A, p.lda, m0, p.M, tid);
  ldB.init(p.B, p.ldb, n0, p.N, tid);
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) {
      double* sA = smem + s * STAGE_ELEMS;
      double* sB = sA + A_ELEMS;
      if (p.vecA) ldA.load(sA, p.A, kb + s * BK, ke);
      else gemm_load_tile<LA, BM, THREADS, BK>(sA, p.A, p.lda, m0, kb + s * BK, p.M, ke, 0, tid);
      if (p.vecB) ldB.load(sB, p.B, kb + s * BK, ke);
      else gemm_load_tile<LB, BN, THREADS, BK>(sB, p.B, p.ldb, n0, kb + s * BK, p.N, ke, 0, tid);
    }
    cp_async_commit();
  }

  // fragment base offsets inside a tile
  const int aoff = (LA == LK) ? ((wm * TM * 8 + g) * LDK + t) : (t * LDMA + wm * TM * 8 + g);
  const int boff = (LB == LK) ? ((wn * TN * 8 + g) * LDK + t) : (t * LDMB + wn * TN * 8 + g);
  constexpr int A_ISTRIDE = (LA == LK) ? 8 * LDK : 8;   // next 8-row DMMA tile
  constexpr int A_KSTRIDE = (LA == LK) ? 4 : 4 * LDMA;  // next k4 step
  constexpr int B_JSTRIDE = (LB == LK) ? 8 * LDK : 8;
  constexpr int B_KSTRIDE = (LB == LK) ? 4 : 4 * LDMB;

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < ktiles) {
        double* sA = smem + (nk % STAGES) * STAGE_ELEMS;
        double* sB = sA + A_ELEMS;
        if (p.vecA) ldA.load(sA, p.A, kb + nk * BK, ke);
        else gemm_load_tile<LA, BM, THREADS, BK>(sA, p.A, p.lda, m0, kb + nk * BK, p.M, ke, 0, tid);
        if (p.vecB) ldB.load(sB, p.B, kb + nk * BK, ke);
        else gemm_load_tile<LB, BN, THREADS, BK>(sB, p.B, p.ldb, n0, kb + nk * BK, p.N, ke, 0, tid);
      }
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * STAGE_ELEMS + aoff;
    const double* sB = smem + (kt % STAGES) * STAGE_ELEMS + A_ELEMS + boff;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = sA[i * A_ISTRIDE + kk * A_KSTRIDE];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = sB[j * B_JSTRIDE + kk * B_KSTRIDE];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // ---- workspace split-K: combine the k-slices of this tile ----
  if (!(EF & EF_ATOMIC) && nsl > 1) {
    const size_t tile_id = (size_t)blockIdx.z * p.ntiles + tile_x;
    constexpr int TILE = BM * BN;
    double* mine = p.ws + (tile_id * p.splitk + blockIdx.y) * TILE;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          __stcg(mine + ((i * TN + j) * 2 + e) * THREADS + tid, acc[i][j][e]);
    __threadfence();
    __syncthreads();   // all warps are past the main loop: the operand stages are dead
    int* s_ticket = reinterpret_cast<int*>(smem);
    if (tid == 0) *s_ticket = atomicAdd(p.ws_cnt + tile_id, 1);
    __syncthreads();
    if (*s_ticket != nsl - 1) return;
    if (tid == 0) p.ws_cnt[tile_id] = 0;   // ready for the next launch (stream ordered)
    __threadfence();
    const double* base = p.ws + tile_id * p.splitk * TILE + tid;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // slices in fixed order (bitwise reproducible); all loads of one slice are independent
    for (int sl = 0; sl < nsl; ++sl) {
      const double* src = base + (size_t)sl * TILE;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        double tmp[TN][2];
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) tmp[j][e] = __ldcg(src + ((i * TN + j) * 2 + e) * THREADS);
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) acc[i][j][e] += tmp[j][e];
      }
    }
  }

  // ---- epilogue ----
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
  for (int j = 0; j < TN; ++j) r_col[j][0] = r_col[j][1] = 0.0;

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + wm * TM * 8 + i * 8 + g;
    // phase 1: issue every global read of this row group back to back (8 loads in flight per
    // operand) so the epilogue of small-K GEMMs is not a chain of exposed L2 latencies
    bool ok[TN][2], upper[TN][2];
    double cin[TN][2], mulv[TN][2], wv[TN][2];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n0 + wn * TN * 8 + j * 8 + 2 * t + e;
        bool o = (row < p.M) && (c < p.N);
        upper[j][e] = false;
        if (EF & EF_LOWER) {
          upper[j][e] = o && (c > row);
          o = o && (c <= row);
        }
        ok[j][e] = o;
        cin[j][e] = 0.0; mulv[j][e] = 1.0; wv[j][e] = 0.0;
        if (o) {
          if ((EF & EF_BETA) && beta != 0.0) cin[j][e] = __ldcg(p.Cin + (size_t)row * p.ldcin + c);
          if (EF & EF_MULMAT) mulv[j][e] = __ldg(p.mul + (size_t)row * p.ldmul + c);
          if (EF & EF_DOT) wv[j][e] = __ldg(p.W + (size_t)row * p.ldw + c);
        }
      }
    }
    // phase 2: combine, reduce, store
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + wn * TN * 8 + j * 8 + 2 * t;
      double v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = col + e;
        double x = alpha * acc[i][j][e];
        if (ok[j][e]) {
          if (EF & EF_DOT) r_dot += x * wv[j][e];
          if (EF & EF_MULMAT) x *= mulv[j][e];
          if (EF & EF_BETA) x += beta * cin[j][e];
          if (EF & EF_COLVEC) x += __ldg(p.colvec + c);
          if (EF & EF_SUMSQ) r_sumsq += x * x;
          if (EF & EF_COLSUM) r_col[j][e] += x;
        }
        v[e] = x;
        // strict upper part of a diagonal tile: store exact zeros so the result is a clean tril()
        if ((EF & EF_LOWER) && !(EF & EF_ATOMIC) && upper[j][e] && p.out) {
          v[e] = 0.0;
          p.out[(size_t)row * p.ldo + c] = 0.0;
        }
      }
      if (EF & EF_ATOMIC) {
        if (ok[j][0]) atomicAdd(p.out + (size_t)row * p.ldo + col, v[0]);
        if (ok[j][1]) atomicAdd(p.out + (size_t)row * p.ldo + col + 1, v[1]);
      } else {
        if (p.out) {
          double* o = p.out + (size_t)row * p.ldo + col;
          if (ok[j][0] && ok[j][1] && p.vecO) {
            *reinterpret_cast<double2*>(o) = make_double2(v[0], v[1]);
          } else {
            if (ok[j][0]) o[0] = v[0];
            if (ok[j][1]) o[1] = v[1];
          }
        }
        if (EF & EF_EXP2) {
          double* o = p.out2 + (size_t)row * p.ldo2 + col;
          if (ok[j][0]) o[0] = exp(v[0]);
          if (ok[j][1]) o[1] = exp(v[1]);
        }
      }
    }
  }
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
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double s = r_col[j][e];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        const int c = n0 + wn * TN * 8 + j * 8 + 2 * t + e;
        if (g == 0 && c < p.N) atomicAdd(p.colsum + c, s);
      }
  }
}

template <class Cfg, int LA, int LB, unsigned EF>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_dmma_kernel(const GemmParams pin) {
  extern __shared__ __align__(16) double smem[];
  gemm_tile_body<Cfg, LA, LB, EF>(pin, blockIdx.x, smem);
}

// Persistent variant: gridDim.x <= ntiles CTAs, each walking the tile list (same tile order).
template <class Cfg, int LA, int LB, unsigned EF>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_dmma_persistent_kernel(const GemmParams pin) {
  extern __shared__ __align__(16) double smem[];
  for (int tile = blockIdx.x; tile < pin.ntiles; tile += gridDim.x) {
    gemm_tile_body<Cfg, LA, LB, EF>(pin, tile, smem);
    __syncthreads();   // every warp is done with the operand stages before the next tile's prologue
  }
}

// Register / query the split-K workspace (device memory owned by the caller; one per device).
void gemm_set_workspace(void* ptr, size_t bytes);

// Number of CTA tiles the launcher will use for an M x N problem (for split-K heuristics).
int gemm_tile_count(int M, int N, int K, bool lower, int batch);

// Host-side launcher (defined in gemm.cu).  `la`,`lb` are Layout values, `ef` the flag set.
int gemm_launch(int la, int lb, unsigned ef, GemmParams p, cudaStream_t stream);

// Convenience: C = alpha*op(A)op(B) + beta*C with optional lower-only output / k-range mode.
int gemm_simple(int la, int lb, int M, int N, int K, double alpha, const double* A, int lda,
                const double* B, int ldb, double beta, double* C, int ldc, bool lower, int kmode,
                int splitk, cudaStream_t stream);

// Batched variant (blockIdx.z) with element strides between consecutive problems.
int gemm_batched(int la, int lb, int M, int N, int K, double alpha, const double* A, int lda,
                 long long strideA, const double* B, int ldb, long long strideB, double beta,
                 double* C, int ldc, long long strideC, int batch, bool lower, int kmode,
                 cudaStream_t stream, int allow_split = 0, int ws_slot = 0);

}  // namespace gpinv
