// Device radix sort of fp32 keys and the 1-D DBSCAN clean ratio built on it.
// BASELINE.json north_star asks for "1-D DBSCAN thresholds on a device sort"; the reference's
// estimate_ratio_dbscan ("# z_score + DBSCAN.py:272-301") only consumes the fraction of non-noise
// points, which in one dimension is fully determined by the sorted order:
//   core(i)   <=>  #{j : |x_j - x_i| <= eps} >= min_samples     (float64 distances, as scikit-learn)
//   noise(i)  <=>  not core(i) and no core point within eps
// LSD radix sort, 4 passes x 8 bits over the order-preserving key (NaN last), stable, original index as optional
// payload, in the "onesweep" form: ONE kernel reads the input once and builds the four global digit histograms, then
// every pass is ONE kernel -- a tile ranks its keys (warp-level match ranking, stable), publishes its 256 digit counts
// to a tile-state array, resolves the counts of all earlier tiles by decoupled look-back (one thread per digit),
// reorders the tile in shared memory and writes digit-contiguous runs.  Traffic per element: 4 B (histogram read) +
// per pass read + write of key (+ payload): 36 B keys only, 64 B with the payload (pass 0 reads the floats and
// generates the index, the last pass writes floats).  The first version (per-tile histograms -> single-CTA scan ->
// scatter, 3 launches per pass, uncoalesced 4-byte scatter) ran at 2 % of the HBM peak.
#include "common.cuh"

namespace sg {
namespace srt {

constexpr int kThreads = 256;          // element-wise kernels
// Tile = 4096 keys: 256 threads x 16 keys when a payload moves along (3 CTAs / SM at 80 registers), 512 threads x 8 keys
// for keys alone -- the best of the shapes measured (profiles/r1e_sort_bench.json).
constexpr int kSortTile = 4096;
constexpr uint32_t kStAgg = 1u << 30, kStIncl = 2u << 30, kStMask = (1u << 30) - 1;   // tile state: flag | count

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// All four digit histograms in one read of the input.  Per-warp shared-memory histograms; each thread keeps the
// (digit, count) run of its last key per pass and only touches shared memory when the digit changes, so clustered
// inputs (constant exponent byte, all-equal vectors) do not serialise on one counter.
__global__ void __launch_bounds__(kThreads) hist4_kernel(const float* __restrict__ v, int64_t n, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t s_h[kThreads / 32][1024];
  for (int i = threadIdx.x; i < (kThreads / 32) * 1024; i += kThreads) (&s_h[0][0])[i] = 0;
  __syncthreads();
  const uint32_t hb = smem_addr_reg(s_h[threadIdx.x >> 5]);
  uint32_t cur[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0};
  stream_f32<4>(v, n, [&](float x, int64_t) {
    const uint32_t k = float_to_key(x);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t d = (k >> (8 * p)) & 255u;
      if (d != cur[p]) {
        if (cnt[p]) red_shared_add(hb + (uint32_t)(p * 256 + cur[p]) * 4u, cnt[p]);
        cur[p] = d;
        cnt[p] = 0;
      }
      ++cnt[p];
    }
  });
#pragma unroll
  for (int p = 0; p < 4; ++p)
    if (cnt[p]) red_shared_add(hb + (uint32_t)(p * 256 + cur[p]) * 4u, cnt[p]);
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += kThreads) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += s_h[w][i];
    if (t) atomicAdd(&ghist[i], t);
  }
}

// exclusive scan of one value per thread over the first 256 threads (8 warps) of the CTA; s_w: 8 words of scratch
__device__ __forceinline__ uint32_t scan256_excl(uint32_t v, uint32_t* s_w) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[w] = x;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  uint32_t base = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) base += (ww < w) ? s_w[ww] : 0u;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  return base + x - v;
}

// One radix pass.  FIRST: keys_in are the caller's floats and the payload is the element index.  LAST: keys are written
// back as floats.  Tiles are claimed from an atomic counter, so every predecessor of a tile is running or done and the
// look-back spin cannot deadlock.  T threads x I keys per thread = one tile.
constexpr int kLook = 8;   // predecessors read per look-back round trip

template <int T, int I, bool PAIRS>
struct SweepCfg {
  static constexpr int kWarps = T / 32;
  static constexpr int kTile = T * I;
  static constexpr int kCntBytes = kWarps * 256 * 2;
  static constexpr int kSmemBytes = kCntBytes + kTile * 4 * (PAIRS ? 2 : 1) + 256 * 4 + 256 * 2 + 64;
  static constexpr int kRegs = I <= 8 ? 64 : 80;
  static constexpr int kMinBlocks = 65536 / (T * kRegs);
};

template <int T, int I, bool FIRST, bool LAST, bool PAIRS>
__global__ void __launch_bounds__(T, (SweepCfg<T, I, PAIRS>::kMinBlocks))
onesweep_kernel(const void* __restrict__ keys_in_, const int32_t* __restrict__ pay_in, uint32_t n, int shift,
                const uint32_t* __restrict__ ghist, uint32_t* __restrict__ tile_state, uint32_t* __restrict__ tile_counter,
                void* __restrict__ keys_out_, int32_t* __restrict__ pay_out) {
  using Cfg = SweepCfg<T, I, PAIRS>;
  static_assert(T >= 256 && Cfg::kTile == kSortTile && Cfg::kTile <= 65535, "one thread per digit; 16-bit local positions");
  extern __shared__ __align__(16) uint8_t smem_sort[];
  uint16_t (*s_cnt)[256] = reinterpret_cast<uint16_t (*)[256]>(smem_sort);   // per-warp digit counts -> offsets in the digit
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(smem_sort + Cfg::kCntBytes);
  int32_t* s_pay = reinterpret_cast<int32_t*>(s_keys + Cfg::kTile);
  uint32_t* s_gbase = reinterpret_cast<uint32_t*>(s_keys + Cfg::kTile * (PAIRS ? 2 : 1));   // global run start - local start
  uint16_t* s_lstart = reinterpret_cast<uint16_t*>(s_gbase + 256);
  uint32_t* s_w = reinterpret_cast<uint32_t*>(s_lstart + 256);
  uint32_t* s_tile = s_w + 8;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_tile = atomicAdd(tile_counter, 1u);
  for (int i = threadIdx.x; i < Cfg::kCntBytes / 4; i += T) reinterpret_cast<uint32_t*>(smem_sort)[i] = 0;
  __syncthreads();
  const uint32_t tile = *s_tile;
  // warp w owns keys [w*32*I, (w+1)*32*I) of the tile, visited in I rounds of 32 in index order (stability)
  const uint32_t wbase = tile * Cfg::kTile + w * (I * 32);
  uint32_t key[I];
  uint32_t rank[I];
#pragma unroll
  for (int j = 0; j < I; ++j) {
    const uint32_t i = wbase + j * 32 + lane;
    if (FIRST) key[j] = i < n ? float_to_key(static_cast<const float*>(keys_in_)[i]) : 0xFFFFFFFFu;
    else key[j] = i < n ? static_cast<const uint32_t*>(keys_in_)[i] : 0xFFFFFFFFu;
  }
#pragma unroll
  for (int j = 0; j < I; ++j) {
    const bool valid = wbase + j * 32 + lane < n;
    const uint32_t d = (key[j] >> shift) & 255u;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    // lanes holding the same digit, from 8 bit-plane votes (measured 10-15 % faster per sort than __match_any_sync)
    unsigned peers = vmask;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const bool bit = (d >> b) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? bal : ~bal;
    }
    const unsigned below = peers & ((1u << lane) - 1u);
    const uint32_t before = valid ? s_cnt[w][d] : 0u;
    rank[j] = before + __popc(below);
    __syncwarp();
    if (valid && below == 0u) s_cnt[w][d] = (uint16_t)(before + __popc(peers));  // the lowest lane of each digit group
    __syncwarp();
  }
  __syncthreads();
  // ---- digit threads: tile counts -> publish, local run starts, first look-back batch in flight ----
  uint32_t run = 0, gstart = 0, lstart = 0, excl = 0;
  uint32_t wd[kLook];
  int64_t t = (int64_t)tile - 1;
  uint32_t* st = tile_state + (size_t)tile * 256 + threadIdx.x;
  if (threadIdx.x < 256) {
    const int d = threadIdx.x;
#pragma unroll
    for (int ww = 0; ww < Cfg::kWarps; ++ww) {
      const uint32_t c = s_cnt[ww][d];
      s_cnt[ww][d] = (uint16_t)run;
      run += c;
    }
    st_relaxed_u32(st, (tile == 0 ? kStIncl : kStAgg) | run);
#pragma unroll
    for (int q = 0; q < kLook; ++q)
      wd[q] = (t - q >= 0) ? ld_relaxed_u32(tile_state + (size_t)(t - q) * 256 + d) : kStIncl;
    lstart = scan256_excl(run, s_w);
    gstart = scan256_excl(ghist[d], s_w);
    s_lstart[d] = (uint16_t)lstart;
  }
  __syncthreads();
  // ---- all threads: reorder the tile in shared memory (needs only the local starts) while the look-back loads fly ----
  int32_t pay[PAIRS ? I : 1];
  if (PAIRS && !FIRST) {
#pragma unroll
    for (int j = 0; j < I; ++j) {
      const uint32_t i = wbase + j * 32 + lane;
      pay[j] = i < n ? pay_in[i] : 0;
    }
  }
#pragma unroll
  for (int j = 0; j < I; ++j) {
    const uint32_t i = wbase + j * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[j] >> shift) & 255u;
      const uint32_t pos = (uint32_t)s_lstart[d] + s_cnt[w][d] + rank[j];
      s_keys[pos] = key[j];
      if (PAIRS) s_pay[pos] = FIRST ? (int32_t)i : pay[j];
    }
  }
  // ---- digit threads: finish the look-back.  kLook predecessors per round trip: with ~300 tiles in flight the nearest
  // tile that knows its inclusive prefix is typically 10-25 tiles back, and one dependent L2 access per tile made
  // that walk the dominant cost of a pass (profiles/r1e_sort_full.txt: barrier + long-scoreboard stalls).
  if (threadIdx.x < 256) {
    const int d = threadIdx.x;
    bool done = t < 0;
    while (!done) {
#pragma unroll
      for (int q = 0; q < kLook; ++q) {
        if (done) break;
        const uint32_t f = wd[q] >> 30;
        if (f == 0u) break;            // not published yet: re-read from this tile on
        excl += wd[q] & kStMask;
        --t;
        done = (f == 2u);
      }
      if (done) break;
#pragma unroll
      for (int q = 0; q < kLook; ++q)
        wd[q] = (t - q >= 0) ? ld_relaxed_u32(tile_state + (size_t)(t - q) * 256 + d) : kStIncl;
    }
    if (tile != 0) st_relaxed_u32(st, kStIncl | (excl + run));
    s_gbase[d] = gstart + excl - lstart;
  }
  __syncthreads();
  const uint32_t tbase = tile * Cfg::kTile;
  const uint32_t count = n - tbase < (uint32_t)Cfg::kTile ? n - tbase : (uint32_t)Cfg::kTile;
  for (uint32_t p = threadIdx.x; p < count; p += T) {
    const uint32_t k = s_keys[p];
    const uint32_t g = s_gbase[(k >> shift) & 255u] + p;
    if (LAST) static_cast<float*>(keys_out_)[g] = key_to_float(k);
    else static_cast<uint32_t*>(keys_out_)[g] = k;
    if (PAIRS) pay_out[g] = s_pay[p];
  }
}

// Neighbourhood [lo, hi] of sorted point i: |s_j - s_i| <= eps evaluated in float64 (as scikit-learn does).  The window
// is found by galloping outwards from i (1, 2, 4, ... steps, then a binary search inside the last interval): O(log
// window) neighbouring, cache-resident loads instead of 2 x log2(n) dependent loads across the whole array.
__device__ __forceinline__ void dbscan_window(const float* __restrict__ s, int64_t n, int64_t i, double eps, int64_t& lo, int64_t& hi) {
  const double x = (double)s[i];
  int64_t a, b = i;                       // first j in [0, i] with x - s[j] <= eps (b: known inside, or i itself)
  for (int64_t step = 1;; step <<= 1) {
    a = b - step;
    if (a < 0) { a = 0; break; }
    if (x - (double)s[a] <= eps) b = a; else { a = a + 1; break; }
  }
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    if (x - (double)s[m] <= eps) b = m; else a = m + 1;
  }
  lo = a;
  a = i;                                  // last j in [i, n-1] with s[j] - x <= eps
  for (int64_t step = 1;; step <<= 1) {
    b = a + step;
    if (b > n - 1) { b = n - 1; break; }
    if ((double)s[b] - x <= eps) a = b; else { b = b - 1; break; }
  }
  while (a < b) {
    const int64_t m = (a + b + 1) >> 1;
    if ((double)s[m] - x <= eps) a = m; else b = m - 1;
  }
  hi = a;
}

__global__ void __launch_bounds__(kThreads) dbscan_core_kernel(const float* __restrict__ s, int64_t n, double eps,
                                                               int min_samples, uint8_t* __restrict__ core) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo, hi;
    dbscan_window(s, n, i, eps, lo, hi);
    core[i] = (hi - lo + 1 >= min_samples) ? 1 : 0;
  }
}

// noise(i) = not core and no core point inside its window.  A non-core point has fewer than min_samples points in its
// window, so the window is simply walked (no compaction of the core positions, no second search).
__global__ void __launch_bounds__(kThreads) dbscan_noise_kernel(const float* __restrict__ s, const uint8_t* __restrict__ core,
                                                                const int32_t* __restrict__ order, int64_t n, double eps,
                                                                uint8_t* __restrict__ noise_out,
                                                                unsigned long long* __restrict__ clean_count) {
  unsigned local = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    bool clean = core[i] != 0;
    if (!clean) {
      int64_t lo, hi;
      dbscan_window(s, n, i, eps, lo, hi);
      for (int64_t j = lo; j <= hi && !clean; ++j) clean = core[j] != 0;
    }
    if (noise_out) noise_out[order[i]] = clean ? 0 : 1;
    local += clean ? 1u : 0u;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(clean_count, (unsigned long long)local);
}

struct SortWs {
  uint32_t* keys[2];
  int32_t* idx[2];
  uint32_t* zero_begin;   // [ghist 4 x 256 | tile counters 4 (+ pad) | tile state 4 x tiles x 256], cleared per sort
  size_t zero_bytes;
  uint32_t* ghist;
  uint32_t* counters;
  uint32_t* state;
  size_t total;
};
static SortWs carve(void* ws, int64_t n) {
  SortWs s;
  uint8_t* p = static_cast<uint8_t*>(ws);
  const size_t nb = align_up((size_t)(n > 0 ? n : 1) * 4, 256);
  const int64_t tiles = ceil_div(n > 0 ? n : 1, kSortTile);
  s.keys[0] = reinterpret_cast<uint32_t*>(p); p += nb;
  s.keys[1] = reinterpret_cast<uint32_t*>(p); p += nb;
  s.idx[0] = reinterpret_cast<int32_t*>(p); p += nb;
  s.idx[1] = reinterpret_cast<int32_t*>(p); p += nb;
  s.zero_begin = reinterpret_cast<uint32_t*>(p);
  s.ghist = s.zero_begin;
  s.counters = s.ghist + 1024;
  s.state = s.counters + 64;
  s.zero_bytes = (size_t)(1024 + 64 + 4 * tiles * 256) * 4;
  p += align_up(s.zero_bytes, 256);
  s.total = (size_t)(p - static_cast<uint8_t*>(ws));
  return s;
}

static int grid1d(int64_t n) {
  int64_t b = ceil_div(n, kThreads * 4);
  const int64_t cap = (int64_t)state().sm_count * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

struct PassArgs {
  const void* kin; const int32_t* pin; uint32_t n; int shift; const uint32_t* ghist; uint32_t* state; uint32_t* counter;
  void* kout; int32_t* pout;
};
template <bool FIRST, bool LAST>
static void launch_pass(bool pairs, int tiles, cudaStream_t st, const PassArgs& a) {
  if (pairs)
    onesweep_kernel<256, 16, FIRST, LAST, true><<<tiles, 256, SweepCfg<256, 16, true>::kSmemBytes, st>>>(
        a.kin, a.pin, a.n, a.shift, a.ghist, a.state, a.counter, a.kout, a.pout);
  else
    onesweep_kernel<512, 8, FIRST, LAST, false><<<tiles, 512, SweepCfg<512, 8, false>::kSmemBytes, st>>>(
        a.kin, a.pin, a.n, a.shift, a.ghist, a.state, a.counter, a.kout, a.pout);
}

template <bool FIRST, bool LAST>
static int set_attrs() {
  SG_CUDA(cudaFuncSetAttribute(onesweep_kernel<256, 16, FIRST, LAST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               SweepCfg<256, 16, true>::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(onesweep_kernel<512, 8, FIRST, LAST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               SweepCfg<512, 8, false>::kSmemBytes));
  return SG_OK;
}

// Sorts v ascending (NaN last, stable).  sorted_out (floats) and / or order_out (original indices) may be null; they
// must not alias the workspace buffers keys[0] / idx[0] (the last pass reads those).
static int sort_keys(const float* v, int64_t n, SortWs& s, cudaStream_t st, float* sorted_out, int32_t* order_out) {
  const int tiles = (int)ceil_div(n, kSortTile);
  const bool pairs = order_out != nullptr;
  SG_CUDA(cudaMemsetAsync(s.zero_begin, 0, s.zero_bytes, st));
  int hb = (int)ceil_div(n, kThreads * 16);
  const int hcap = state().sm_count * 4;
  if (hb > hcap) hb = hcap;
  hist4_kernel<<<hb < 1 ? 1 : hb, kThreads, 0, st>>>(v, n, s.ghist);
  SG_LAUNCH_CHECK();
  const uint32_t un = (uint32_t)n;
  const size_t ts = (size_t)tiles * 256;
  launch_pass<true, false>(pairs, tiles, st, PassArgs{v, nullptr, un, 0, s.ghist, s.state, s.counters, s.keys[0], s.idx[0]});
  launch_pass<false, false>(pairs, tiles, st, PassArgs{s.keys[0], s.idx[0], un, 8, s.ghist + 256, s.state + ts, s.counters + 1, s.keys[1], s.idx[1]});
  launch_pass<false, false>(pairs, tiles, st, PassArgs{s.keys[1], s.idx[1], un, 16, s.ghist + 512, s.state + 2 * ts, s.counters + 2, s.keys[0], s.idx[0]});
  if (sorted_out)
    launch_pass<false, true>(pairs, tiles, st, PassArgs{s.keys[0], s.idx[0], un, 24, s.ghist + 768, s.state + 3 * ts, s.counters + 3, sorted_out, order_out});
  else
    launch_pass<false, false>(pairs, tiles, st, PassArgs{s.keys[0], s.idx[0], un, 24, s.ghist + 768, s.state + 3 * ts, s.counters + 3, s.keys[1], order_out});
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // namespace srt
}  // namespace sg

extern "C" {

int sg_sort_init_attributes() {
  using namespace sg::srt;
  int r = set_attrs<true, false>();
  if (r == SG_OK) r = set_attrs<false, false>();
  if (r == SG_OK) r = set_attrs<false, true>();
  return r;
}

size_t sg_sort_workspace_bytes(int64_t n) {
  sg::srt::SortWs s = sg::srt::carve(nullptr, n);
  return s.total + 256;
}

int sg_sort_f32(const float* v, int64_t n, float* sorted_out, int32_t* order_out, void* workspace, void* stream) {
  using namespace sg::srt;
  SG_READY();
  SG_REQUIRE(n >= 0 && n <= ((int64_t)1 << 30) && workspace, "arguments (n <= 2^30)");
  if (n == 0) return SG_OK;
  SG_REQUIRE(v != nullptr, "v");
  SG_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  SortWs s = carve(workspace, n);
  return sort_keys(v, n, s, sg::as_stream(stream), sorted_out, order_out);
}

size_t sg_dbscan1d_workspace_bytes(int64_t n) {
  const size_t nn = (size_t)(n > 0 ? n : 1);
  return sg_sort_workspace_bytes(n) + sg::align_up(nn * 4, 256) + sg::align_up(nn, 256) + 256;
}

int sg_dbscan1d(const float* v, int64_t n, double eps, int min_samples, int64_t* counts_out, uint8_t* noise_out,
                void* workspace, void* stream) {
  using namespace sg::srt;
  SG_READY();
  SG_REQUIRE(v && counts_out && workspace, "null pointer");
  SG_REQUIRE(n >= 1 && n <= ((int64_t)1 << 30), "n (1 .. 2^30)");
  SG_REQUIRE(eps >= 0.0 && min_samples >= 1, "eps/min_samples");
  SG_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t st = sg::as_stream(stream);
  SortWs s = carve(workspace, n);
  uint8_t* p = static_cast<uint8_t*>(workspace) + sg::align_up(s.total, 256);
  float* sorted = reinterpret_cast<float*>(p); p += sg::align_up((size_t)n * 4, 256);
  uint8_t* core = p; p += sg::align_up((size_t)n, 256);
  unsigned long long* clean = reinterpret_cast<unsigned long long*>(p);
  int32_t* order = noise_out ? s.idx[1] : nullptr;   // per-sample noise flags need the original positions
  int r = sort_keys(v, n, s, st, sorted, order);
  if (r != SG_OK) return r;
  SG_CUDA(cudaMemsetAsync(clean, 0, 8, st));
  dbscan_core_kernel<<<grid1d(n), kThreads, 0, st>>>(sorted, n, eps, min_samples, core);
  dbscan_noise_kernel<<<grid1d(n), kThreads, 0, st>>>(sorted, core, order, n, eps, noise_out, clean);
  SG_LAUNCH_CHECK();
  SG_CUDA(cudaMemcpyAsync(counts_out, clean, 8, cudaMemcpyDeviceToDevice, st));
  return SG_OK;
}

}  // extern "C"
