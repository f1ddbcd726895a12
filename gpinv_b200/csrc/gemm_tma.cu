// Host side of the TMA / mbarrier / stream-K DMMA GEMM (gemm_tma.cuh): tensor-map construction,
// the equal-work partition of the tile list, and the kernel instantiations.
#include "gemm_tma.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <vector>

namespace gpinv {

namespace {

// ------------------------------- tensor maps -------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// operand with `rows` rows and K columns of the GEMM view, element (r,k) at base[r*ld + k] (LK) or
// base[k*ld + r] (LMN); `batch` problems `stride` elements apart (stride 0: one shared problem)
bool make_operand_map(CUtensorMap* map, int layout, const double* base, int rows, int K, int ld, int batch,
                      long long stride, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const int nb = (batch > 1 && stride != 0) ? batch : 1;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  if (layout == LK) {
    gdim[0] = (cuuint64_t)K; gdim[1] = (cuuint64_t)rows;
    box[0] = 16; box[1] = (cuuint32_t)box_rows;
  } else {
    gdim[0] = (cuuint64_t)rows; gdim[1] = (cuuint64_t)K;
    box[0] = 16; box[1] = 16;
  }
  gdim[2] = (cuuint64_t)nb;
  box[2] = 1;
  gstr[0] = (cuuint64_t)ld * 8;
  gstr[1] = (nb > 1) ? (cuuint64_t)stride * 8 : gstr[0] * gdim[1];
  if (gstr[0] % 16 || gstr[1] % 16 || gstr[0] >= (1ull << 40) || gstr[1] >= (1ull << 40)) return false;
  if (gdim[0] * 8 > gstr[0]) return false;       // rows must not overlap
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------- work partition -------------------------------
struct SchedKey {
  int M, N, K, kmode, lower, batch, small, max_ctas, allow_split, atomic;
  bool operator==(const SchedKey& o) const { return memcmp(this, &o, sizeof(SchedKey)) == 0; }
};
struct SchedEntry {
  SchedKey key;
  TmaSched sched;
  int tiles_m, tiles_n, ntiles;
  bool split;      // some tile is shared between CTAs (needs the workspace)
};

std::mutex g_sched_mutex;
std::vector<SchedEntry*> g_sched_cache;

int sm_count() {
  static int table[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!table[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    table[dev] = n;
  }
  return table[dev];
}

// Cut the ordered tile list into contiguous, (nearly) equal-work ranges.  `split`: a tile's k range may
// be shared between CTAs (stream-K); otherwise whole tiles only.
const SchedEntry* get_sched(const GemmParams& p, bool lower, bool small, bool atomic, bool ws_ok) {
  SchedKey key;
  memset(&key, 0, sizeof(key));
  key.M = p.M; key.N = p.N; key.K = p.K; key.kmode = p.kmode; key.lower = lower; key.batch = p.batch;
  static const bool no_split = (getenv("GPINV_NO_SPLIT") != nullptr);
  key.small = small; key.max_ctas = p.max_ctas; key.allow_split = (p.allow_split && ws_ok && !no_split) ? 1 : 0;
  key.atomic = atomic;
  std::lock_guard<std::mutex> lock(g_sched_mutex);
  for (SchedEntry* e : g_sched_cache)
    if (e->key == key) return e;
  const int bm = small ? 64 : 128, bk = 16;
  SchedEntry* e = new SchedEntry();
  e->key = key;
  e->tiles_m = cdiv(p.M, bm);
  e->tiles_n = cdiv(p.N, bm);
  e->ntiles = lower ? e->tiles_m * (e->tiles_m + 1) / 2 : e->tiles_m * e->tiles_n;
  const long long T = (long long)e->ntiles * p.batch;
  if (T <= 0 || T >= 65535) { delete e; return nullptr; }
  std::vector<int> units((size_t)T), kts((size_t)T);
  long long W = 0;
  int kt_max = 0;
  for (int x = 0; x < e->ntiles; ++x) {
    int tm, tn, kb, ke;
    gemm_decode_tile(x, e->tiles_m, e->tiles_n, p.kmode, lower, tm, tn);
    gemm_krange(p.kmode, p.K, tm * bm, tn * bm, bm, bm, bk, kb, ke);
    const int kt = (ke > kb) ? cdiv(ke - kb, bk) : 0;
    if (kt >= 65535) { delete e; return nullptr; }
    for (int z = 0; z < p.batch; ++z) {
      kts[(size_t)z * e->ntiles + x] = kt;
      units[(size_t)z * e->ntiles + x] = kt > 0 ? kt : 1;
    }
    W += (long long)(kt > 0 ? kt : 1) * p.batch;
    if (kt > kt_max) kt_max = kt;
  }
  // every tile also costs a fixed prologue / epilogue, worth about 6 k-tiles of the big tile
  const int fixed = small ? 3 : 6;
  static const int small_occ = getenv("GPINV_TMA_SMALL_OCC") ? atoi(getenv("GPINV_TMA_SMALL_OCC")) : 3;
  int gmax = sm_count() * (small ? small_occ : 1);
  if (p.max_ctas > 0 && p.max_ctas < gmax) gmax = p.max_ctas;
  if (gmax > TMA_MAX_CTAS) gmax = TMA_MAX_CTAS;
  TmaSched& s = e->sched;
  s.total_tiles = (int)T;

  // (1) whole tiles: greedy contiguous cut at the running average
  auto whole_partition = [&](int G, std::vector<unsigned>& st) -> double {
    st.assign(G + 1, 0);
    long long Wf = W + (long long)fixed * T;
    double mk = 0;
    long long acc = 0, done = 0;
    int c = 0;
    st[0] = 0;
    long long cur = 0;
    for (long long i = 0; i < T; ++i) {
      const long long u = units[(size_t)i] + fixed;
      // close the current range before tile i when that brings it closer to the ideal boundary
      const double ideal = (double)Wf * (c + 1) / G;
      if (c + 1 < G && cur > 0 && (acc + u - ideal) > (ideal - acc)) {
        if (cur > mk) mk = (double)cur;
        ++c;
        st[c] = (unsigned)i << 16;
        cur = 0;
      }
      acc += u;
      cur += u;
      ++done;
    }
    if (cur > mk) mk = (double)cur;
    for (int k = c + 1; k <= G; ++k) st[k] = (unsigned)T << 16;
    return mk;
  };
  int G0 = (int)(T < gmax ? T : gmax);
  std::vector<unsigned> st_whole;
  const double mk_whole = whole_partition(G0, st_whole);
  const double ideal_whole = (double)(W + (long long)fixed * T) / gmax;

  // (2) stream-K: equal COST per CTA, where a tile costs its k-tiles plus the fixed prologue / epilogue
  // (a CTA whose range is made of many short tiles -- the tail of a triangular product -- would
  // otherwise carry up to 40 % more than its share); cuts are snapped away from tile boundaries
  bool use_split = false;
  std::vector<unsigned> st_split;
  int Gs = 0;
  if ((key.allow_split || atomic) && W >= 16) {
    const int min_units = small ? 12 : 8;
    const long long Wf = W + (long long)fixed * T;
    Gs = (int)((W / min_units) < gmax ? (W / min_units) : gmax);
    if (Gs > 24 * T) Gs = (int)(24 * T);     // the last arriver sums the slices serially: bound their number
    if (Gs < 1) Gs = 1;
    const double mk_split = (double)Wf / Gs + fixed + (atomic ? 0 : 4);
    if (Gs > 1 && mk_split < 0.93 * mk_whole) {
      st_split.assign(Gs + 1, 0);
      long long i = 0, before = 0;       // tile i starts at cost offset `before`
      for (int c = 1; c < Gs; ++c) {
        const long long target = Wf * c / Gs;
        while (i < T && before + units[(size_t)i] + fixed <= target) { before += units[(size_t)i] + fixed; ++i; }
        // inside tile i: the first `fixed` cost units are its prologue, the rest its k-tiles
        long long koff = target - before - fixed / 2;
        const int kt = (i < T) ? kts[(size_t)i] : 0;
        if (koff < 3 || kt == 0) koff = 0;                 // snap to the tile's start
        else if (kt - koff < 3) { before += units[(size_t)i] + fixed; ++i; koff = 0; }   // ... or to its end
        unsigned pos = ((unsigned)i << 16) | (unsigned)koff;
        if (pos < st_split[c - 1]) pos = st_split[c - 1];
        st_split[c] = pos;
      }
      st_split[Gs] = (unsigned)T << 16;
      // drop empty ranges (identical consecutive starts)
      std::vector<unsigned> compact;
      compact.push_back(st_split[0]);
      for (int c = 1; c <= Gs; ++c)
        if (st_split[c] != compact.back()) compact.push_back(st_split[c]);
      if (compact.back() != ((unsigned)T << 16)) compact.push_back((unsigned)T << 16);
      st_split = compact;
      Gs = (int)st_split.size() - 1;
      use_split = Gs >= 1;
    }
  }
  (void)ideal_whole;
  (void)kt_max;
  const std::vector<unsigned>& st = use_split ? st_split : st_whole;
  // trim trailing empty CTAs of the whole-tile partition
  int G = use_split ? Gs : G0;
  while (G > 1 && st[G - 1] == st[G]) --G;
  s.nctas = G;
  for (int c = 0; c <= G; ++c) s.start[c] = st[c];
  s.start[G] = (unsigned)T << 16;
  e->split = false;
  for (int c = 1; c < G; ++c)
    if (s.start[c] & 0xffffu) e->split = true;
  // L2-friendly list order for uniform products (see TmaSched): sort the list positions by (index among the
  // whole tiles of the owning CTA, CTA) and hand them the tiles of the grid in enumeration order
  s.use_perm = 0;
  static const bool no_perm = (getenv("GPINV_TMA_NO_PERM") != nullptr);
  if (!no_perm && p.kmode == 0 && !lower && p.batch == 1 && T <= TMA_MAX_PERM && T > G && G > 1) {
    std::vector<std::pair<long long, int>> key((size_t)T);
    int c = 0;
    for (int i = 0; i < (int)T; ++i) {
      while (c + 1 < G && (int)(s.start[c + 1] >> 16) <= i && !((int)(s.start[c + 1] >> 16) == i && (s.start[c + 1] & 0xffffu)))
        ++c;                                     // owner: the CTA that processes the START of tile i
      // (a range that starts inside a tile processes that piece last: its first whole tile is the next one)
      const int off = i - ((int)(s.start[c] >> 16) + ((s.start[c] & 0xffffu) ? 1 : 0));
      key[(size_t)i] = std::make_pair((long long)off * G + c, i);
    }
    std::sort(key.begin(), key.end());
    for (int j = 0; j < (int)T; ++j) s.perm[key[(size_t)j].second] = (unsigned short)j;
    s.use_perm = 1;
  }
  g_sched_cache.push_back(e);
  return e;
}

// per-device "dynamic shared memory attribute already set" flags, one per kernel instantiation
struct DeviceOnce {
  bool done[64] = {};
};

template <class Cfg, int LA, int LB, unsigned EF>
int launch_tma(const CUtensorMap& mA, const CUtensorMap& mB, const GemmParams& p, const TmaSched& sched,
               cudaStream_t stream) {
  static DeviceOnce once;
  auto kern = gemm_tma_kernel<Cfg, LA, LB, EF>;
  int dev = 0;
  GPINV_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!once.done[dev]) {
    GPINV_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    once.done[dev] = true;
  }
  kern<<<sched.nctas, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(mA, mB, p, sched);
  GPINV_CHECK_LAUNCH();
  return GPINV_OK;
}

#define GPINV_TMA_CASE(LA_, LB_, EF_)                                 \
  if (la == (LA_) && lb == (LB_) && ef == (unsigned)(EF_))            \
    return launch_tma<Cfg, LA_, LB_, (unsigned)(EF_)>(mA, mB, p, sched, stream);

#define GPINV_TMA_GENERIC(LA_, LB_)                     \
  GPINV_TMA_CASE(LA_, LB_, EF_BETA)                     \
  GPINV_TMA_CASE(LA_, LB_, EF_BETA | EF_LOWER)          \
  GPINV_TMA_CASE(LA_, LB_, EF_ATOMIC)                   \
  GPINV_TMA_CASE(LA_, LB_, EF_ATOMIC | EF_LOWER)

template <class Cfg>
int dispatch_tma(int la, int lb, unsigned ef, const CUtensorMap& mA, const CUtensorMap& mB, const GemmParams& p,
                 const TmaSched& sched, cudaStream_t stream) {
  GPINV_TMA_GENERIC(LK, LK)
  GPINV_TMA_GENERIC(LK, LMN)
  GPINV_TMA_GENERIC(LMN, LMN)
  GPINV_TMA_GENERIC(LMN, LK)
  GPINV_TMA_CASE(LK, LK, EF_COLVEC | EF_SUMSQ)               // U = E Q^T + m ; R = Y - X A^T
  GPINV_TMA_CASE(LK, LK, EF_COLVEC)                          // F = sv U Lc^T + mu
  GPINV_TMA_CASE(LK, LK, EF_COLVEC | EF_EXP2)                // ... and exp(F)
  GPINV_TMA_CASE(LK, LMN, EF_BETA | EF_EXP2)                   // Kronecker apply: F = sv L1 T + mean and exp(F)
  GPINV_TMA_CASE(LK, LMN, EF_MULMAT)                         // Fbar = X o (R A) c
  GPINV_TMA_CASE(LK, LMN, EF_BETA | EF_COLSUM)               // Ubar = sv Fbar Lc + klbar U
  GPINV_TMA_CASE(LMN, LMN, EF_LOWER)                         // Qbar = tril(Ubar^T E)
  GPINV_TMA_CASE(LMN, LMN, EF_LOWER | EF_BETA | EF_DOT)      // Lbar += tril(Fbar^T U), <.,Lc>
  return 1;   // no instantiation: the caller falls back to the cp.async kernels
}

}  // namespace

// workspace registry lives in gemm.cu
bool gemm_workspace_for_device(void** ptr, size_t* bytes);
constexpr int TMA_CNT_SLOTS = 16384;     // must equal CNT_SLOTS in gemm.cu

bool gemm_tma_try_launch(int la, int lb, unsigned ef, GemmParams& p, bool small, cudaStream_t stream, int* rc) {
  static const bool disabled = (getenv("GPINV_GEMM_LEGACY") != nullptr);
  if (disabled || small) return false;     // 64x64-tile problems: cp.async kernels (see gemm_tma.cuh)
  // TMA needs 16-byte aligned bases and strides; anything else stays on the cp.async path.  Ragged
  // M / N / K are fine: out-of-bounds parts of a box are zero-filled by the hardware.
  if (!aligned16(p.A) || !aligned16(p.B) || (p.lda & 1) || (p.ldb & 1)) return false;
  if (p.batch > 1 && ((p.strideA & 1) || (p.strideB & 1))) return false;
  if (p.splitk > 1) return false;          // caller-imposed atomic split: legacy semantics
  const bool lower = (ef & EF_LOWER) != 0;
  const bool atomic = (ef & EF_ATOMIC) != 0;
  void* wsp = nullptr;
  size_t wsb = 0;
  const bool have_ws = gemm_workspace_for_device(&wsp, &wsb);
  const size_t tile_bytes = (small ? 64 * 64 : 128 * 128) * sizeof(double);
  const int gmax_all = small ? TMA_MAX_CTAS : 160;
  const bool ws_ok = have_ws && wsb > TMA_CNT_SLOTS * sizeof(int) &&
                     (wsb - TMA_CNT_SLOTS * sizeof(int)) / 2 >= (size_t)2 * gmax_all * tile_bytes;
  const SchedEntry* e = get_sched(p, lower, small, atomic, ws_ok);
  if (!e) return false;
  p.tiles_m = e->tiles_m;
  p.tiles_n = e->tiles_n;
  p.ntiles = e->ntiles;
  if (e->split && !atomic) {
    char* base = static_cast<char*>(wsp);
    p.ws_cnt = reinterpret_cast<int*>(base) + (p.ws_slot == 1 ? TMA_CNT_SLOTS / 2 : 0);
    double* w0 = reinterpret_cast<double*>(base + TMA_CNT_SLOTS * sizeof(int));
    p.ws = w0 + (p.ws_slot == 1 ? ((wsb - TMA_CNT_SLOTS * sizeof(int)) / sizeof(double)) / 2 : 0);
  }
  const int rowsA = small ? 64 : 128;
  CUtensorMap mA, mB;
  if (!make_operand_map(&mA, la, p.A, p.M, p.K, p.lda, p.batch, p.strideA, rowsA)) return false;
  if (!make_operand_map(&mB, lb, p.B, p.N, p.K, p.ldb, p.batch, p.strideB, rowsA)) return false;
  static const bool log_gemm = (getenv("GPINV_GEMM_LOG") != nullptr);
  if (log_gemm)
    fprintf(stderr, "gemm-log tma la=%d lb=%d ef=%u M=%d N=%d K=%d kmode=%d batch=%d cfg=%s ctas=%d split=%d\n", la, lb,
            ef, p.M, p.N, p.K, p.kmode, p.batch, small ? "small" : "big", e->sched.nctas, (int)e->split);
  int r = dispatch_tma<TmaBig>(la, lb, ef, mA, mB, p, e->sched, stream);
  if (r == 1) return false;
  *rc = r;
  return true;
}

}  // namespace gpinv
