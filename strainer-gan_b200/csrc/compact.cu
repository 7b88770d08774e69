// Stream compaction.
// Replaces np.where(losses < thr)[0] ("#strainer gan.py:384"), the boolean gathers real_cpu[mask] /
// real_cpu[~mask] and torch.cat ("# 상위 10% 제거해서 fake image에 concate.py:247-249,268") and the pool
// gather ("# strainer gan + concate.py:623-627").  The reference does nonzero + index on the host
// (one device sync per gather); here one single-pass kernel (decoupled look-back scan over dynamic
// tile ids) produces ascending indices / stable destinations, and a vectorised row mover copies the
// 48 KiB image rows with 128-bit streaming loads/stores.
#include "common.cuh"

namespace sg {
namespace cmp {

constexpr int kThreads = 256;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;  // 4096 elements per tile

// tile status word: [63:62] flag (0 invalid, 1 aggregate, 2 inclusive prefix) | [61:0] value
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagIncl = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

struct ScanWs {
  unsigned int tile_counter;
  unsigned int pad[3];
  // followed by tile status words
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Exclusive prefix of this tile's aggregate over all earlier tiles (decoupled look-back, warp 0).
// Tiles are claimed from an atomic counter, so every predecessor is already running or done.
__device__ __forceinline__ unsigned long long lookback(unsigned long long* status, int tile, unsigned long long agg,
                                                       unsigned long long* s_excl) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    if (lane == 0) st_release_u64(&status[tile], (tile == 0 ? kFlagIncl : kFlagAgg) | agg);
    unsigned long long excl = 0;
    int pos = tile - 1;
    while (pos >= 0) {
      const int idx = pos - lane;
      unsigned long long w = kFlagIncl;  // virtual tiles before 0: inclusive prefix 0
      if (idx >= 0) {
        do { w = ld_acquire_u64(&status[idx]); } while ((w >> 62) == 0ull);
      }
      const unsigned incl_mask = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
      const int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;  // nearest tile with a full prefix
      unsigned long long contrib = (lane <= first_incl) ? (w & kValMask) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
      excl += contrib;
      if (incl_mask) break;
      pos -= 32;
    }
    if (lane == 0) {
      if (tile != 0) st_release_u64(&status[tile], kFlagIncl | (excl + agg));
      *s_excl = excl;
    }
  }
  __syncthreads();
  return *s_excl;
}

// Block-wide exclusive scan of one small count per thread; returns the thread's offset, *total = sum.
__device__ __forceinline__ int block_excl_scan(int c, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = (lane < kThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < kThreads / 32) s_warp[lane] = s;
  }
  __syncthreads();
  *total = s_warp[kThreads / 32 - 1];
  return x - c + (w ? s_warp[w - 1] : 0);
}

// v[i] CMP thr -> ascending global indices (int64) + optional byte mask.
// One CTA per tile of kBigTile = 16384 elements; tile ids are claimed from an atomic counter in launch
// order (every predecessor is running or done -> the look-back spin cannot deadlock).  The tile is
// large on purpose: a look-back window resolves at most 32 tiles per L2 round trip, which with
// 4096-element tiles capped the kernel at ~45 % of HBM peak (ncu: 23 of 29 stall cycles on the
// barrier behind the look-back).
// Layout inside a tile is STRIPED per warp: warp w owns 2048 consecutive elements, slot j of lane l is
// element w*2048 + j*32 + l.  One __ballot_sync per slot gives the keep-mask of 32 consecutive
// elements, so a lane's output position is popc(mask & lanes-below) past a running base and the kept
// lanes of a slot write one contiguous run of int64: no shared-memory staging, no serial loop.
// Each lane keeps only its own 64 keep-bits; the ballots are re-formed in the store phase.
constexpr int kBigItems = 64;
constexpr int kCiThreads = 256;
constexpr int kBigTile = kCiThreads * kBigItems;  // 16384

template <int CMP>
__device__ __forceinline__ bool cmp_static(float v, float thr) {
  bool r;
  if ((CMP & 3) == SG_LT) r = v < thr;
  else if ((CMP & 3) == SG_LE) r = v <= thr;
  else if ((CMP & 3) == SG_GE) r = v >= thr;
  else r = v > thr;
  return (CMP & SG_NOT) ? !r : r;
}
// predicated 64-bit store (no branch around it: the SASS is one @P STG instead of BSSY/BRA/STG/BSYNC)
__device__ __forceinline__ void st_if_u64(int64_t* p, int64_t v, bool pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u64 [%0], %1;\n\t}"
               :: "l"(p), "l"(v), "r"((uint32_t)pred) : "memory");
}

// keep-bits of the 64 slots of one lane (slot j = element wbase + j*32 + lane)
template <int CMP, bool FULL>
__device__ __forceinline__ unsigned long long ci_keep_bits(const float* __restrict__ v, int64_t wbase, int64_t n,
                                                           float thr, int lane) {
  const float* p = v + wbase + lane;
  uint32_t lo = 0u, hi = 0u;
#pragma unroll
  for (int g = 0; g < kBigItems / 16; ++g) {
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int s = g * 16 + j;
      x[j] = (FULL || wbase + s * 32 + lane < n) ? __ldg(p + s * 32) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int s = g * 16 + j;
      const bool keep = (FULL || wbase + s * 32 + lane < n) && cmp_static<CMP>(x[j], thr);
      if (s < 32) lo |= (uint32_t)keep << s;
      else hi |= (uint32_t)keep << (s - 32);
    }
  }
  return ((unsigned long long)hi << 32) | lo;
}

template <int CMP, bool MASK>
__global__ void __launch_bounds__(kCiThreads) compact_indices_kernel(const float* __restrict__ v, int64_t n,
                                                                   const float* __restrict__ thr_p,
                                                                   int64_t index_base, int64_t* __restrict__ idx_out,
                                                                   int64_t* __restrict__ count_out,
                                                                   uint8_t* __restrict__ mask_out, ScanWs* ws,
                                                                   int num_tiles) {
  __shared__ int s_tile;
  __shared__ int s_wtot[kCiThreads / 32];
  __shared__ unsigned long long s_excl;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(ws + 1);
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ws->tile_counter, 1u);
  const float thr = *thr_p;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  __syncthreads();
  const int tile = s_tile;
  const int64_t wbase = (int64_t)tile * kBigTile + w * (kBigItems * 32);
  const bool full = wbase + kBigItems * 32 <= n;
  const unsigned long long bits = full ? ci_keep_bits<CMP, true>(v, wbase, n, thr, lane)
                                       : ci_keep_bits<CMP, false>(v, wbase, n, thr, lane);
  int wtotal = __popcll(bits);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wtotal += __shfl_xor_sync(0xffffffffu, wtotal, o);
  if (lane == 0) s_wtot[w] = wtotal;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int ww = 0; ww < kCiThreads / 32; ++ww) {
    const int c = s_wtot[ww];
    if (ww < w) woff += c;
    total += c;
  }
  const unsigned long long excl = lookback(status, tile, (unsigned long long)total, &s_excl);
  int64_t* dst = idx_out + excl + woff;
  const int64_t gidx = index_base + wbase + lane;
  const uint32_t blo = (uint32_t)bits, bhi = (uint32_t)(bits >> 32);
  int o = 0;
#pragma unroll
  for (int j = 0; j < kBigItems; ++j) {
    const bool keep = (((j < 32) ? blo : bhi) >> (j & 31)) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    st_if_u64(dst + (o + __popc(m & lt)), gidx + j * 32, keep);
    o += __popc(m);
  }
  if (MASK) {
#pragma unroll 16
    for (int j = 0; j < kBigItems; ++j) {
      const int64_t i = wbase + j * 32 + lane;
      if (full || i < n) mask_out[i] = (uint8_t)((bits >> j) & 1ull);
    }
  }
  if (tile == num_tiles - 1 && threadIdx.x == 0) *count_out = (int64_t)(excl + total);
}

template <int CMP>
static void launch_compact_indices(int grid, cudaStream_t st, const float* v, int64_t n, const float* thr,
                                   int64_t index_base, int64_t* idx_out, int64_t* count_out, uint8_t* mask_out,
                                   ScanWs* ws, int num_tiles) {
  if (mask_out)
    compact_indices_kernel<CMP, true><<<grid, kCiThreads, 0, st>>>(v, n, thr, index_base, idx_out, count_out, mask_out, ws,
                                                                 num_tiles);
  else
    compact_indices_kernel<CMP, false><<<grid, kCiThreads, 0, st>>>(v, n, thr, index_base, idx_out, count_out, mask_out, ws,
                                                                  num_tiles);
}

// Stable two-way partition destinations from a byte mask: dest[i] = rank among kept rows (mask != 0)
// or -(rank among dropped rows) - 1.  counts_out = {#kept, #dropped}.
__global__ void __launch_bounds__(kThreads) partition_dest_kernel(const uint8_t* __restrict__ mask, int64_t n,
                                                                  int64_t* __restrict__ dest,
                                                                  int64_t* __restrict__ counts_out, ScanWs* ws,
                                                                  int num_tiles) {
  __shared__ int s_tile;
  __shared__ int s_warp[kThreads / 32];
  __shared__ unsigned long long s_excl;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(ws + 1);
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ws->tile_counter, 1u);
  __syncthreads();
  {
    const int tile = s_tile;
    const int64_t base = (int64_t)tile * kTile + (int64_t)threadIdx.x * kItems;
    unsigned flags = 0;
#pragma unroll
    for (int j = 0; j < kItems; ++j)
      if (base + j < n && mask[base + j] != 0) flags |= 1u << j;
    const int c = __popc(flags);
    int total;
    const int toff = block_excl_scan(c, s_warp, &total);
    const unsigned long long excl = lookback(status, tile, (unsigned long long)total, &s_excl);
    int64_t kept_rank = (int64_t)excl + toff;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      const int64_t i = base + j;
      if (i < n) {
        if (flags & (1u << j)) dest[i] = kept_rank++;
        else dest[i] = -(i - kept_rank) - 1;  // dropped rank = i - (#kept before i)
      }
    }
    if (tile == num_tiles - 1 && threadIdx.x == 0) {
      counts_out[0] = (int64_t)(excl + total);
      counts_out[1] = n - (int64_t)(excl + total);
    }
  }
}

// One CTA per row: 128-bit streaming copy of row i to kept[dest] or dropped[-dest-1].
__global__ void __launch_bounds__(256) move_rows_kernel(const int4* __restrict__ rows, int64_t row_vec,
                                                        const int64_t* __restrict__ dest, int4* __restrict__ kept,
                                                        int4* __restrict__ dropped) {
  const int64_t i = blockIdx.x;
  const int64_t d = dest[i];
  int4* out = (d >= 0) ? kept : dropped;
  if (out == nullptr) return;
  const int64_t r = (d >= 0) ? d : (-d - 1);
  const int4* src = rows + i * row_vec;
  int4* dst = out + r * row_vec;
  int64_t j = threadIdx.x;
  for (; j + 3 * 256 < row_vec; j += 4 * 256) {
    const int4 a = ldg_stream_i4(src + j), b = ldg_stream_i4(src + j + 256);
    const int4 c = ldg_stream_i4(src + j + 512), e = ldg_stream_i4(src + j + 768);
    stg_stream_i4(dst + j, a); stg_stream_i4(dst + j + 256, b);
    stg_stream_i4(dst + j + 512, c); stg_stream_i4(dst + j + 768, e);
  }
  for (; j < row_vec; j += 256) stg_stream_i4(dst + j, ldg_stream_i4(src + j));
}

// out[i] = a[i] for i < na, out[na + j] = b[j] behind them: one CTA per row (the concat of the strain block; when the
// b rows were written in place by sg_strain_rows_concat only the first na CTAs are launched).
__global__ void __launch_bounds__(256) concat_rows_kernel(const int4* __restrict__ a, int64_t na, const int4* __restrict__ b,
                                                          int64_t row_vec, int4* __restrict__ out) {
  const int64_t i = blockIdx.x;
  const int4* src = (i < na) ? a + i * row_vec : b + (i - na) * row_vec;
  int4* dst = out + i * row_vec;
  int64_t j = threadIdx.x;
  for (; j + 3 * 256 < row_vec; j += 4 * 256) {
    const int4 p = ldg_stream_i4(src + j), q = ldg_stream_i4(src + j + 256);
    const int4 r = ldg_stream_i4(src + j + 512), t = ldg_stream_i4(src + j + 768);
    stg_stream_i4(dst + j, p); stg_stream_i4(dst + j + 256, q);
    stg_stream_i4(dst + j + 512, r); stg_stream_i4(dst + j + 768, t);
  }
  for (; j < row_vec; j += 256) stg_stream_i4(dst + j, ldg_stream_i4(src + j));
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const int4* __restrict__ rows, int64_t row_vec,
                                                          const int64_t* __restrict__ idx,
                                                          const int64_t* __restrict__ count_dev,
                                                          int4* __restrict__ out) {
  const int64_t i = blockIdx.x;
  if (count_dev && i >= *count_dev) return;
  const int4* src = rows + idx[i] * row_vec;
  int4* dst = out + i * row_vec;
  int64_t j = threadIdx.x;
  for (; j + 3 * 256 < row_vec; j += 4 * 256) {
    const int4 a = ldg_stream_i4(src + j), b = ldg_stream_i4(src + j + 256);
    const int4 c = ldg_stream_i4(src + j + 512), e = ldg_stream_i4(src + j + 768);
    stg_stream_i4(dst + j, a); stg_stream_i4(dst + j + 256, b);
    stg_stream_i4(dst + j + 512, c); stg_stream_i4(dst + j + 768, e);
  }
  for (; j < row_vec; j += 256) stg_stream_i4(dst + j, ldg_stream_i4(src + j));
}

// ------------------------------------------------------------------------------------------
// The whole selection half of the in-batch strain block in ONE CTA
// ("# 상위 10% 제거해서 fake image에 concate.py:246-249": thr = torch.quantile(scores, q); mask = scores >= thr;
// real[mask], real[~mask]) for B <= 2048 scores: bitonic sort of the radix keys in shared memory, the
// library's interpolation rule, the mask, and the stable two-way partition ranks -- the reference spends
// a sort kernel, a lerp, a compare, two nonzero() (each a device sync) and two index_select launches here.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) strain_small_kernel(const float* __restrict__ v, int n, int pow2, int k0, int k1,
                                                            float w, int lerp_kind, int cmp, float* __restrict__ thr_out,
                                                            uint8_t* __restrict__ mask_out, int64_t* __restrict__ dest,
                                                            int64_t* __restrict__ counts_out, int tail) {
  extern __shared__ uint32_t s_keys[];   // pow2 keys
  __shared__ int s_nan;
  __shared__ float s_thr;
  __shared__ int s_wsum[32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (t == 0) s_nan = 0;
  __syncthreads();
  for (int i = t; i < pow2; i += blockDim.x) {
    uint32_t key = 0xFFFFFFFFu;
    if (i < n) {
      key = float_to_key(v[i]);
      if (key == 0xFFFFFFFFu) s_nan = 1;
    }
    s_keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = t; i < (pow2 >> 1); i += blockDim.x) {
        const int lo = ((i / stride) * stride * 2) + (i % stride);
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const uint32_t a = s_keys[lo], b = s_keys[hi];
        if ((a > b) == up) { s_keys[lo] = b; s_keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (t == 0) {
    float r = __uint_as_float(0x7FC00000u);
    if (!s_nan) r = lerp_rule(key_to_float(s_keys[k0]), key_to_float(s_keys[k1]), w, lerp_kind);
    s_thr = r;
    thr_out[0] = r;
  }
  __syncthreads();
  const float thr = s_thr;
  // two consecutive elements per thread (n <= 2048), block exclusive scan of the keep flags
  const int i0 = 2 * t, i1 = 2 * t + 1;
  const bool k0f = i0 < n && cmp_apply(v[i0], thr, cmp);
  const bool k1f = i1 < n && cmp_apply(v[i1], thr, cmp);
  const int c = (int)k0f + (int)k1f;
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_wsum[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int sacc = (lane < (int)(blockDim.x >> 5)) ? s_wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, sacc, o);
      if (lane >= o) sacc += y;
    }
    s_wsum[lane] = sacc;
  }
  __syncthreads();
  const int kept_before = x - c + (wid ? s_wsum[wid - 1] : 0);
  const int total = s_wsum[(blockDim.x >> 5) - 1];
  // tail: the dropped rows go to rows [total, n) of ONE n-row buffer -- right behind the `total` rows the generator
  // will write into it ("# 상위 10% 제거해서 fake image에 concate.py:268": cat([fake, filtered_fake]) without the copy)
  const int dbase = tail ? total : 0;
  if (i0 < n) {
    if (mask_out) mask_out[i0] = (uint8_t)k0f;
    dest[i0] = k0f ? (int64_t)kept_before : -(int64_t)(dbase + i0 - kept_before) - 1;
  }
  if (i1 < n) {
    const int kb = kept_before + (int)k0f;
    if (mask_out) mask_out[i1] = (uint8_t)k1f;
    dest[i1] = k1f ? (int64_t)kb : -(int64_t)(dbase + i1 - kb) - 1;
  }
  if (t == 0) { counts_out[0] = total; counts_out[1] = n - total; }
}

static size_t scan_ws_bytes(int64_t n) {
  const int64_t tiles = ceil_div(n > 0 ? n : 1, kTile);
  return align_up(sizeof(ScanWs) + (size_t)tiles * 8, 256);
}

}  // namespace cmp
}  // namespace sg

extern "C" {

size_t sg_compact_workspace_bytes(int64_t n) {
  // scan state + one int64 destination per element (row partition)
  return sg::cmp::scan_ws_bytes(n) + sg::align_up((size_t)(n > 0 ? n : 1) * 8, 256);
}

int sg_compact_indices(const float* v, int64_t n, const float* thr, int cmp, int64_t index_base, int64_t* idx_out,
                       int64_t* count_out, uint8_t* mask_out, void* workspace, void* stream) {
  using namespace sg::cmp;
  SG_READY();
  SG_REQUIRE(n >= 0 && thr && count_out && workspace, "arguments");
  SG_REQUIRE(n == 0 || (v && idx_out), "v/idx_out");
  SG_REQUIRE(cmp >= 0 && cmp <= (SG_GT | SG_NOT), "cmp");
  SG_REQUIRE(n < ((int64_t)1 << 42), "n too large");
  cudaStream_t st = sg::as_stream(stream);
  if (n == 0) {
    SG_CUDA(cudaMemsetAsync(count_out, 0, 8, st));
    return SG_OK;
  }
  const int num_tiles = (int)sg::ceil_div(n, kBigTile);
  SG_CUDA(cudaMemsetAsync(workspace, 0, sizeof(ScanWs) + (size_t)num_tiles * 8, st));
  const int grid = num_tiles;  // one CTA per tile; ids come from the atomic counter in launch order
  ScanWs* sws = static_cast<ScanWs*>(workspace);
  switch (cmp) {
#define SG_CI_CASE(C) case C: launch_compact_indices<C>(grid, st, v, n, thr, index_base, idx_out, count_out, mask_out, sws, num_tiles); break;
    SG_CI_CASE(0) SG_CI_CASE(1) SG_CI_CASE(2) SG_CI_CASE(3) SG_CI_CASE(4) SG_CI_CASE(5) SG_CI_CASE(6) SG_CI_CASE(7)
#undef SG_CI_CASE
    default: break;
  }
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_compact_rows(const void* rows, int64_t n, int64_t row_bytes, const uint8_t* mask, void* kept, void* dropped,
                    int64_t* counts_out, void* workspace, void* stream) {
  using namespace sg::cmp;
  SG_READY();
  SG_REQUIRE(n >= 0 && counts_out && workspace, "arguments");
  SG_REQUIRE(n == 0 || (rows && mask), "rows/mask");
  SG_REQUIRE(row_bytes > 0 && (row_bytes & 15) == 0, "row_bytes must be a positive multiple of 16");
  SG_REQUIRE(n <= 0x7FFFFFFF, "n too large for one launch");
  SG_REQUIRE(((uintptr_t)rows & 15) == 0 && ((uintptr_t)kept & 15) == 0 && ((uintptr_t)dropped & 15) == 0,
             "row buffers must be 16-byte aligned");
  cudaStream_t st = sg::as_stream(stream);
  if (n == 0) {
    SG_CUDA(cudaMemsetAsync(counts_out, 0, 16, st));
    return SG_OK;
  }
  const int num_tiles = (int)sg::ceil_div(n, kTile);
  const size_t sbytes = scan_ws_bytes(n);
  SG_CUDA(cudaMemsetAsync(workspace, 0, sbytes, st));
  int64_t* dest = reinterpret_cast<int64_t*>(static_cast<uint8_t*>(workspace) + sbytes);
  const int grid = num_tiles;
  partition_dest_kernel<<<grid, kThreads, 0, st>>>(mask, n, dest, counts_out, static_cast<ScanWs*>(workspace), num_tiles);
  SG_LAUNCH_CHECK();
  move_rows_kernel<<<(unsigned)n, 256, 0, st>>>(static_cast<const int4*>(rows), row_bytes / 16, dest,
                                                static_cast<int4*>(kept), static_cast<int4*>(dropped));
  SG_LAUNCH_CHECK();
  return SG_OK;
}

static int strain_rows_impl(const float* scores, int64_t n, int k0, int k1, float weight, int lerp_kind, int cmp,
                            const void* rows, int64_t row_bytes, void* kept, void* dropped, uint8_t* mask_out, float* thr_out,
                            int64_t* counts_out, void* workspace, int tail, void* stream) {
  using namespace sg::cmp;
  SG_READY();
  SG_REQUIRE(scores && thr_out && counts_out && workspace, "null pointer");
  SG_REQUIRE(n >= 1 && n <= 2048, "n must be in [1, 2048] (one batch)");
  SG_REQUIRE(k0 >= 0 && k0 < n && k1 >= k0 && k1 < n, "need 0 <= k0 <= k1 < n");
  SG_REQUIRE(lerp_kind == SG_LERP_NUMPY || lerp_kind == SG_LERP_TORCH, "lerp_kind");
  SG_REQUIRE(cmp >= 0 && cmp <= (SG_GT | SG_NOT), "cmp");
  SG_REQUIRE(rows == nullptr || (row_bytes > 0 && (row_bytes & 15) == 0), "row_bytes must be a positive multiple of 16");
  SG_REQUIRE(((uintptr_t)rows & 15) == 0 && ((uintptr_t)kept & 15) == 0 && ((uintptr_t)dropped & 15) == 0,
             "row buffers must be 16-byte aligned");
  cudaStream_t st = sg::as_stream(stream);
  int pow2 = 2;
  while (pow2 < n) pow2 <<= 1;
  int threads = pow2 / 2;
  if (threads < 32) threads = 32;
  int64_t* dest = static_cast<int64_t*>(workspace);
  strain_small_kernel<<<1, threads, pow2 * sizeof(uint32_t), st>>>(scores, (int)n, pow2, k0, k1, weight, lerp_kind, cmp,
                                                                  thr_out, mask_out, dest, counts_out, tail);
  SG_LAUNCH_CHECK();
  if (rows != nullptr) {
    move_rows_kernel<<<(unsigned)n, 256, 0, st>>>(static_cast<const int4*>(rows), row_bytes / 16, dest,
                                                  static_cast<int4*>(kept), static_cast<int4*>(dropped));
    SG_LAUNCH_CHECK();
  }
  return SG_OK;
}

int sg_strain_rows(const float* scores, int64_t n, int k0, int k1, float weight, int lerp_kind, int cmp, const void* rows,
                   int64_t row_bytes, void* kept, void* dropped, uint8_t* mask_out, float* thr_out, int64_t* counts_out,
                   void* workspace, void* stream) {
  return strain_rows_impl(scores, n, k0, k1, weight, lerp_kind, cmp, rows, row_bytes, kept, dropped, mask_out, thr_out,
                          counts_out, workspace, 0, stream);
}

int sg_strain_rows_concat(const float* scores, int64_t n, int k0, int k1, float weight, int lerp_kind, int cmp,
                          const void* rows, int64_t row_bytes, void* kept, void* concat, uint8_t* mask_out, float* thr_out,
                          int64_t* counts_out, void* workspace, void* stream) {
  SG_REQUIRE(rows != nullptr && concat != nullptr, "rows / concat buffer");
  return strain_rows_impl(scores, n, k0, k1, weight, lerp_kind, cmp, rows, row_bytes, kept, concat, mask_out, thr_out,
                          counts_out, workspace, 1, stream);
}

int sg_concat_rows(const void* a, int64_t na, const void* b, int64_t nb, int64_t row_bytes, void* out, void* stream) {
  SG_READY();
  SG_REQUIRE(na >= 0 && nb >= 0 && na + nb <= 0x7FFFFFFF, "row counts");
  SG_REQUIRE(row_bytes > 0 && (row_bytes & 15) == 0, "row_bytes must be a positive multiple of 16");
  if (na + nb == 0) return SG_OK;
  SG_REQUIRE(out && (na == 0 || a) && (nb == 0 || b), "null pointer");
  SG_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)out & 15) == 0,
             "row buffers must be 16-byte aligned");
  const bool b_in_place = nb == 0 || b == static_cast<const uint8_t*>(out) + (size_t)na * row_bytes;
  const int64_t grid = na + (b_in_place ? 0 : nb);
  if (grid == 0) return SG_OK;
  sg::cmp::concat_rows_kernel<<<(unsigned)grid, 256, 0, sg::as_stream(stream)>>>(
      static_cast<const int4*>(a), na, static_cast<const int4*>(b), row_bytes / 16, static_cast<int4*>(out));
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_gather_rows(const void* rows, int64_t row_bytes, const int64_t* idx, int64_t count, const int64_t* count_dev,
                   void* out, void* stream) {
  SG_READY();
  SG_REQUIRE(count >= 0 && count <= 0x7FFFFFFF, "count");
  SG_REQUIRE(row_bytes > 0 && (row_bytes & 15) == 0, "row_bytes must be a positive multiple of 16");
  if (count == 0) return SG_OK;
  SG_REQUIRE(rows && idx && out, "null pointer");
  SG_REQUIRE(((uintptr_t)rows & 15) == 0 && ((uintptr_t)out & 15) == 0, "row buffers must be 16-byte aligned");
  sg::cmp::gather_rows_kernel<<<(unsigned)count, 256, 0, sg::as_stream(stream)>>>(
      static_cast<const int4*>(rows), row_bytes / 16, idx, count_dev, static_cast<int4*>(out));
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
