// Order statistics and thresholds.
// Replaces np.percentile ("#strainer gan.py:381", "# 종합 loss.py:288-292") and torch.quantile
// ("# 상위 10% 제거해서 fake image에 concate.py:246", "# z_score + DBSCAN.py:323"): the reference
// partitions / fully sorts on the CPU; here a 3-pass MSD radix select (11/11/10 bits) finds x_(k),
// one more pass finds x_(k+1) as the smallest key above it, and a one-thread kernel applies the
// library's interpolation rule with the same rounding steps.  Integer histograms make the result
// independent of summation order, hence bit-identical under any sharding.
#include <math.h>

#include "common.cuh"

namespace sg {
namespace sel {

// internal workspace words (after the public ones)
constexpr int W_PREFIX = 260;      // selected key bits so far (8, 16, 24, 32 bits)
constexpr int W_KREM_LO = 262;     // remaining rank inside the current bucket (u64)
constexpr int W_SELKEY = 264;      // final key of x_(k)
constexpr int W_NEEDNEXT = 265;    // 1: x_(k+1) has a larger key than x_(k)
constexpr int W_NEXTIN = 266;     // smallest populated last-pass digit above the selected one (or ~0)
// one-pass (sampled) selection state
constexpr int W_LO_F = 270;        // float bits of the lower pivot (canonical, -inf when open)
constexpr int W_HI_F = 271;        // float bits of the upper pivot (+inf when open)
constexpr int W_BELOW = 272;       // u64: elements < lower pivot
constexpr int W_CAND = 274;        // u64: elements in [lower, upper] = append cursor of the candidate buffer
constexpr int W_NANF = 276;        // NaNs seen by the filter pass
constexpr int W_OVERFLOW = 277;    // candidate buffer overflowed
constexpr int W_TICKET = 278;      // CTAs of the filter pass that are done
constexpr int W_USECAND = 279;     // 1: the radix passes run over the candidate buffer with the reduced rank
constexpr int W_TICKET_SAMPLE = 284; // CTAs of the sampling kernel that are done
constexpr int kSampleCount = 32768;
constexpr int kSampleThreads = 1024;
constexpr int kSampleSmem = (kSampleCount + 2 * 256 * 32) * 4;   // keys + lane-private histograms of both pivots

// Pass plan: 4 x 8 bits.  Loss vectors are heavily skewed in their top bits (a handful of exponents),
// so a shared histogram serialises on same-address atomics.  With 256 bins the histogram fits in 32
// lane-private copies (bin*32 + lane: every lane owns a bank), which makes every shared atomic
// conflict-free in EVERY pass; passes 1..3 only count the elements inside the selected bucket.
// The last pass also tracks the smallest key ABOVE the 24-bit bucket, so x_(k+1) needs no extra read.
__global__ void begin_kernel(uint32_t* ws, unsigned long long k) {
  for (int i = threadIdx.x; i < SG_SELECT_WS_WORDS; i += blockDim.x) ws[i] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    ws[SG_SELECT_WS_MINABOVE] = 0xFFFFFFFFu;
    ws[W_NEXTIN] = 0xFFFFFFFFu;
    *reinterpret_cast<unsigned long long*>(ws + W_KREM_LO) = k;
  }
}

template <int PASS>
__global__ void __launch_bounds__(512, 4) hist_kernel(const float* __restrict__ v, int64_t n, uint32_t* __restrict__ ws) {
  __shared__ uint32_t s_hist[256 * 32];
  __shared__ uint32_t s_nan, s_min;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) s_hist[i] = 0u;
  if (threadIdx.x == 0) { s_nan = 0u; s_min = 0xFFFFFFFFu; }
  __syncthreads();
  const uint32_t prefix = ws[W_PREFIX];
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t hbase = smem_addr_reg(s_hist) + lane * 4u;   // lane-private column of the [256][32] histogram
  constexpr int kShift = 24 - 8 * PASS;   // digit position
  uint32_t nan_local = 0, min_local = 0xFFFFFFFFu;
  stream_f32<4>(v, n, [&](float f, int64_t) {
    const uint32_t key = float_to_key(f);
    if constexpr (PASS == 0) {
      nan_local += (key == 0xFFFFFFFFu);
      red_shared_add(hbase + ((key >> 24) << 7), 1u);
    } else {
      const uint32_t hi = key >> (kShift + 8);
      if (hi == prefix) red_shared_add(hbase + (((key >> kShift) & 255u) << 7), 1u);
      if (PASS == 3 && hi > prefix) min_local = min(min_local, key);
    }
  });
  if (PASS == 0 && nan_local) atomicAdd(&s_nan, nan_local);
  if (PASS == 3 && min_local != 0xFFFFFFFFu) atomicMin(&s_min, min_local);
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) c += s_hist[i * 32 + ((k + i) & 31)];
    if (c) atomicAdd(&ws[SG_SELECT_WS_HIST + i], c);
  }
  if (PASS == 0 && threadIdx.x == 0 && s_nan) atomicAdd(&ws[SG_SELECT_WS_NANCOUNT], s_nan);
  if (PASS == 3 && threadIdx.x == 0 && s_min != 0xFFFFFFFFu) atomicMin(&ws[SG_SELECT_WS_MINABOVE], s_min);
}

// ---- single-device radix phases: ONE cooperative kernel ---------------------------------------------
// All four 8-bit passes in one launch (cudaLaunchCooperativeKernel: every CTA is resident, so a grid barrier
// may spin).  CTA b owns a contiguous slice of the source; the first kCacheKeys keys of the slice are kept in
// shared memory, so a source that fits (the candidate buffer of the one-pass filter, or any n <= grid * 40960)
// is read ONCE and passes 1..3 never touch global memory.  Per pass: lane-private histogram -> RED into the
// pass's own 256 global bins -> grid barrier -> every CTA derives the bucket redundantly from those bins.
constexpr int kTailThreads = 1024;
constexpr int kCacheKeys = 40960;                                  // 160 KB
constexpr int kTailSmem = (kCacheKeys + 256 * 32) * 4;             // + 32 KB lane-private histogram
constexpr int W_BAR = 288;         // grid-barrier arrivals
constexpr int W_ERR = 289;         // 1: a grid barrier timed out (protocol bug guard)
constexpr int W_HIST4 = 512;       // [4][256] per-pass digit histograms of the cooperative kernel

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t gtimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void grid_barrier(uint32_t* ws, uint32_t target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ws + W_BAR, 1u);
    const uint64_t t0 = gtimer_ns();
    while (ld_acquire_u32(ws + W_BAR) < target) {
      if (gtimer_ns() - t0 > 2000000000ull) { ws[W_ERR] = 1u; break; }   // never hang the GPU
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kTailThreads, 1) radix_phases_kernel(const float* __restrict__ v, int64_t n,
                                                                       const float* __restrict__ cand,
                                                                       uint32_t* __restrict__ ws, float* __restrict__ out2) {
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_keys = s_dyn;
  uint32_t* s_hist = s_dyn + kCacheKeys;
  __shared__ unsigned long long s_warp[8];
  __shared__ unsigned long long s_before, s_cnt;
  __shared__ int s_bucket;
  __shared__ uint32_t s_nan, s_min, s_next;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  // candidate mode: every key lies in [klo, khi]; a pass whose digit (and everything above it) is the same in
  // both bounds has a single populated bin and is skipped (typically the exponent byte, often two passes)
  uint32_t klo = 0u, khi = 0xFFFFFFFFu;
  if (cand != nullptr && ws[W_USECAND] != 0u) {
    v = cand;
    n = (int64_t)*reinterpret_cast<const unsigned long long*>(ws + W_CAND);
    klo = float_to_key(__uint_as_float(ws[W_LO_F]));
    khi = float_to_key(__uint_as_float(ws[W_HI_F]));
  }
  const int64_t per = ((n + gridDim.x - 1) / gridDim.x + 3) & ~(int64_t)3;
  const int64_t beg = min((int64_t)blockIdx.x * per, n), end = min(beg + per, n);
  const int ncache = (int)min((int64_t)kCacheKeys, end - beg);
  if ((reinterpret_cast<uintptr_t>(v + beg) & 15) == 0) {   // beg is a multiple of 4: aligned whenever v is
    const float4* v4 = reinterpret_cast<const float4*>(v + beg);
    for (int i = t; i < (ncache >> 2); i += kTailThreads) {
      const float4 q = ldg_stream4(v4 + i);
      s_keys[4 * i] = float_to_key(q.x); s_keys[4 * i + 1] = float_to_key(q.y);
      s_keys[4 * i + 2] = float_to_key(q.z); s_keys[4 * i + 3] = float_to_key(q.w);
    }
    for (int i = (ncache & ~3) + t; i < ncache; i += kTailThreads) s_keys[i] = float_to_key(__ldg(v + beg + i));
  } else {
    for (int i = t; i < ncache; i += kTailThreads) s_keys[i] = float_to_key(__ldg(v + beg + i));
  }
  const float* rest = v + beg + ncache;          // streamed again in every pass (only when the slice exceeds the cache)
  const int64_t nrest = end - beg - ncache;
  unsigned long long krem = *reinterpret_cast<const unsigned long long*>(ws + W_KREM_LO);
  uint32_t prefix = 0u, arrivals = 0u;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {   // skippable passes (the last one always runs: it resolves x_(k+1))
    if ((klo >> (24 - 8 * pass)) != (khi >> (24 - 8 * pass))) break;
    prefix = klo >> (24 - 8 * pass);
  }
#pragma unroll 1
  for (int pass = (prefix == 0u && (klo >> 24) != (khi >> 24)) ? 0 : ((klo >> 16) != (khi >> 16)) ? 1 : ((klo >> 8) != (khi >> 8)) ? 2 : 3;
       pass < 4; ++pass) {
    for (int i = t; i < 256 * 32; i += kTailThreads) s_hist[i] = 0u;
    if (t == 0) { s_nan = 0u; s_min = 0xFFFFFFFFu; s_bucket = -1; s_next = 0xFFFFFFFFu; }
    __syncthreads();
    const int shift = 24 - 8 * pass;
    uint32_t nan_local = 0, min_local = 0xFFFFFFFFu;
    auto visit = [&](uint32_t key) {
      const uint32_t hi = (pass == 0) ? 0u : (key >> (shift + 8));
      if (hi == prefix) atomicAdd(&s_hist[(((key >> shift) & 255u) << 5) + lane], 1u);
      if (pass == 0) nan_local += (key == 0xFFFFFFFFu);
      if (pass == 3 && hi > prefix) min_local = min(min_local, key);
    };
#pragma unroll 4
    for (int i = t; i < ncache; i += kTailThreads) visit(s_keys[i]);
    if (nrest > 0)
      stream_f32_grid<4>(rest, nrest, (int64_t)t, (int64_t)kTailThreads, [&](float f, int64_t) { visit(float_to_key(f)); });
    if (nan_local) atomicAdd(&s_nan, nan_local);
    if (min_local != 0xFFFFFFFFu) atomicMin(&s_min, min_local);
    __syncthreads();
    uint32_t* gh = ws + W_HIST4 + pass * 256;
    if (t < 256) {
      uint32_t c = 0;
#pragma unroll
      for (int k = 0; k < 32; ++k) c += s_hist[t * 32 + ((k + t) & 31)];
      if (c) atomicAdd(gh + t, c);
    }
    if (t == 0) {
      if (s_nan) atomicAdd(&ws[SG_SELECT_WS_NANCOUNT], s_nan);
      if (s_min != 0xFFFFFFFFu) atomicMin(&ws[SG_SELECT_WS_MINABOVE], s_min);
    }
    arrivals += gridDim.x;
    grid_barrier(ws, arrivals);
    // every CTA: bucket of rank krem among the 256 bins (thread t < 256 owns bin t)
    const bool on = t < 256;
    const unsigned long long c = on ? (unsigned long long)__ldcg(gh + t) : 0ull;
    unsigned long long x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (on && lane == 31) s_warp[w] = x;
    __syncthreads();
    unsigned long long run = x - c;
    if (on) for (int ww = 0; ww < w; ++ww) run += s_warp[ww];
    if (on && krem >= run && krem < run + c) { s_bucket = t; s_before = run; s_cnt = c; }
    __syncthreads();
    int b = s_bucket;
    if (pass == 3 && on && b >= 0 && t > b && c != 0ull) atomicMin(&s_next, (uint32_t)t);
    __syncthreads();
    unsigned long long before = s_before, cnt = s_cnt;
    if (b < 0) { b = 255; before = 0; cnt = 0; }
    prefix = (prefix << 8) | (uint32_t)b;
    krem -= before;
    if (pass == 3 && blockIdx.x == 0 && t == 0) {
      const float nanv = __uint_as_float(0x7FC00000u);
      if (__ldcg(ws + SG_SELECT_WS_NANCOUNT) != 0u) {
        out2[0] = nanv; out2[1] = nanv;
      } else if (__ldcg(ws + W_ERR) != 0u) {
        // a grid barrier gave up (possible only when the device is time-sliced under the cooperative launch): never
        // hand out statistics built from incomplete histograms; sg_select_check reports the condition
        out2[0] = out2[1] = __uint_as_float(0x7FC00000u);
      } else {
        const float a = key_to_float(prefix);
        float bb = a;
        if (krem + 1 >= cnt) {   // x_(k+1) has a larger key
          const uint32_t above = __ldcg(ws + SG_SELECT_WS_MINABOVE);
          if (s_next != 0xFFFFFFFFu) bb = key_to_float((prefix & ~0xFFu) | s_next);
          else if (above != 0xFFFFFFFFu) bb = key_to_float(above);
        }
        out2[0] = a; out2[1] = bb;
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void step_body(uint32_t* ws, int pass);
__device__ __forceinline__ void finish_body(const uint32_t* ws, float* out2);

// ---- one-pass selection ---------------------------------------------------------------------------
// SURVEY §8(d) counts ONE 4-byte read per element for the global select.  A radix select needs the bucket of
// pass p before it can run pass p+1, i.e. four full reads.  Instead: (1) one CTA draws kSampleCount samples
// at pseudo-random offsets of equal strides and radix-selects two pivots lo <= x_(k) <= x_(k+1) <= hi from them
// (ranks k*S/n -+ 4.5 sigma of the binomial spread), (2) ONE streaming pass counts the elements below lo and
// appends the elements of [lo, hi] (~3 % of n) to a candidate buffer (warp-private shared staging, one global
// atomic per 128 candidates), (3) the last CTA verifies that both order statistics lie inside the candidates
// and switches the four radix passes to that (L2-resident) buffer with the reduced rank.  If the check fails
// (pivots missed, heavy ties overflowing the buffer) the radix passes run over the full input: always exact.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(kSampleThreads) sample_pivot_kernel(const float* __restrict__ v, int64_t n,
                                                                      uint32_t* __restrict__ ws, uint32_t* __restrict__ skeys,
                                                                      uint32_t* __restrict__ ticket, unsigned long long k,
                                                                      int r_lo, int r_hi) {
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_keys = s_dyn;                        // kSampleCount keys
  uint32_t* s_h = s_dyn + kSampleCount;            // [2 targets][256 bins][32 lane-private copies]
  __shared__ uint32_t s_c[2][256];
  __shared__ uint32_t s_prefix[2], s_krem[2];
  __shared__ int s_last;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the filter pass may become resident now (it waits)
  // every CTA draws kSampleThreads samples (one DRAM sector each: spread over many SMs, a single SM cannot keep
  // 32768 sector misses in flight); the last CTA to finish selects the pivots
  const int64_t stride = n / kSampleCount;        // >= 64 (callers use this path for n >= 2^21 only)
  {
    const int s = blockIdx.x * kSampleThreads + t;
    skeys[s] = float_to_key(__ldg(v + (int64_t)s * stride + (int64_t)(mix32((uint32_t)s) % (uint32_t)stride)));
  }
  __threadfence();
  __syncthreads();
  if (t == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = t; i < SG_SELECT_WS_WORDS; i += kSampleThreads) ws[i] = 0u;   // also resets the ticket
#pragma unroll 8
  for (int j = 0; j < kSampleCount / kSampleThreads; ++j) s_keys[j * kSampleThreads + t] = __ldcg(skeys + j * kSampleThreads + t);
  if (t < 2) { s_prefix[t] = 0u; s_krem[t] = (uint32_t)(t == 0 ? max(r_lo, 0) : min(r_hi, kSampleCount - 1)); }
  __syncthreads();
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const uint32_t p0 = s_prefix[0], p1 = s_prefix[1];
    const bool same = (p0 == p1);                  // both targets still in the same bucket: one histogram serves both
    for (int i = t; i < (same ? 1 : 2) * 256 * 32; i += kSampleThreads) s_h[i] = 0u;
    __syncthreads();
    const int shift = 24 - 8 * pass;
#pragma unroll 4
    for (int s = t; s < kSampleCount; s += kSampleThreads) {
      const uint32_t key = s_keys[s];
      const uint32_t hi = (pass == 0) ? 0u : (key >> (shift + 8));
      const uint32_t d = ((key >> shift) & 255u) * 32u + lane;
      if (hi == p0) atomicAdd(&s_h[d], 1u);
      if (!same && hi == p1) atomicAdd(&s_h[256 * 32 + d], 1u);
    }
    __syncthreads();
    if (t < (same ? 256 : 512)) {
      const uint32_t* h = s_h + (t >> 8) * 256 * 32 + (t & 255) * 32;
      uint32_t c = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) c += h[(j + t) & 31];
      s_c[t >> 8][t & 255] = c;
    }
    __syncthreads();
    if (w < 2) {   // warp q narrows target q: 8 bins per lane
      const uint32_t* cc = s_c[same ? 0 : w];
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = cc[lane * 8 + j]; sum += c[j]; }
      uint32_t x = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      uint32_t run = x - sum;
      const uint32_t kr = s_krem[w];
      if (kr >= run && kr < run + sum) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (kr >= run && kr < run + c[j]) { s_prefix[w] = (s_prefix[w] << 8) | (uint32_t)(lane * 8 + j); s_krem[w] = kr - run; }
          run += c[j];
        }
      }
    }
    __syncthreads();
  }
  if (t == 0) {
    ws[SG_SELECT_WS_MINABOVE] = 0xFFFFFFFFu;
    ws[W_NEXTIN] = 0xFFFFFFFFu;
    *reinterpret_cast<unsigned long long*>(ws + W_KREM_LO) = k;
    // NaN keys in the sample sort last: a NaN pivot opens that side
    const uint32_t klo = s_prefix[0], khi = s_prefix[1];
    ws[W_LO_F] = (r_lo < 0 || klo == 0xFFFFFFFFu) ? 0xFF800000u : __float_as_uint(key_to_float(klo));
    ws[W_HI_F] = (r_hi >= kSampleCount || khi == 0xFFFFFFFFu) ? 0x7F800000u : __float_as_uint(key_to_float(khi));
  }
}

constexpr int kFilterThreads = 512;
constexpr int kSlots = 20;    // staged candidates per THREAD; flushed by the warp when any lane holds > kSlots - 8

__device__ __forceinline__ void count_if_lt(uint32_t& c, float f, float lo) {
  asm("{\n\t.reg .pred q;\n\tsetp.lt.f32 q, %1, %2;\n\t@q add.u32 %0, %0, 1;\n\t}" : "+r"(c) : "f"(f), "f"(lo));
}

__global__ void __launch_bounds__(kFilterThreads, 4) filter_kernel(const float* __restrict__ v, int64_t n,
                                                                   float* __restrict__ cand, unsigned long long cap,
                                                                   uint32_t* __restrict__ ws, unsigned long long k) {
  __shared__ float s_stage[kSlots * kFilterThreads];   // [slot][thread]: a thread's column is bank-conflict free
  __shared__ unsigned long long s_below;
  __shared__ uint32_t s_nan;
  __shared__ int s_lastf;
  const int lane = threadIdx.x & 31;
  // launched with programmatic stream serialisation: the CTAs are resident while the pivot kernel still runs and
  // wait here for its completion (its writes to ws are then visible)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const float lo = __uint_as_float(ws[W_LO_F]), hi = __uint_as_float(ws[W_HI_F]);
  if (threadIdx.x == 0) { s_below = 0ull; s_nan = 0u; }
  __syncthreads();
  float* stage = s_stage + threadIdx.x;
  int cnt = 0;                       // this thread's staged candidates
  uint32_t below = 0, nan = 0;
  auto flush = [&]() {               // whole warp
    int x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    const int total = __shfl_sync(0xffffffffu, x, 31);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(reinterpret_cast<unsigned long long*>(ws + W_CAND), (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + (unsigned long long)total <= cap) {
      float* dst = cand + base + (x - cnt);
      for (int j = 0; j < cnt; ++j) dst[j] = stage[j * kFilterThreads];
    } else if (lane == 0) {
      ws[W_OVERFLOW] = 1u;
    }
    cnt = 0;
  };
  auto push = [&](float f) {
    count_if_lt(below, f, lo);
    if (f >= lo && f <= hi) { stage[cnt * kFilterThreads] = f; ++cnt; }
  };
  auto visit4 = [&](const float4& q) {
    push(q.x); push(q.y); push(q.z); push(q.w);
    const float tsum = (q.x + q.y) + (q.z + q.w);      // NaN if any element is NaN (or inf - inf: checked exactly)
    if (tsum != tsum) nan += (q.x != q.x) + (q.y != q.y) + (q.z != q.z) + (q.w != q.w);
  };
  // warp-uniform trip counts: every lane of a warp runs the same iterations
  const int64_t gwarp = (blockIdx.x * (int64_t)kFilterThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kFilterThreads) >> 5;
  const bool aligned = (reinterpret_cast<uintptr_t>(v) & 15) == 0;
  const int64_t n4 = aligned ? (n >> 2) : 0;
  const float4* v4 = reinterpret_cast<const float4*>(v);
  constexpr int U = 4;
  int64_t i = gwarp * 32;            // float4 index of lane 0
  for (; i + (U - 1) * nwarps * 32 + 32 <= n4; i += U * nwarps * 32) {
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = ldg_stream4(v4 + i + u * nwarps * 32 + lane);
#pragma unroll
    for (int u = 0; u < U; u += 2) {
      visit4(q[u]); visit4(q[u + 1]);
      if (__any_sync(0xffffffffu, cnt > kSlots - 8)) flush();
    }
  }
  for (; i < n4; i += nwarps * 32) {
    if (i + lane < n4) visit4(ldg_stream4(v4 + i + lane));
    if (__any_sync(0xffffffffu, cnt > kSlots - 8)) flush();
  }
  for (int64_t e = (n4 << 2) + gwarp * 32; e < n; e += nwarps * 32) {
    if (e + lane < n) { const float f = v[e + lane]; push(f); nan += (f != f); }
    if (__any_sync(0xffffffffu, cnt > kSlots - 8)) flush();
  }
  if (__any_sync(0xffffffffu, cnt > 0)) flush();
  below = warp_sum(below);
  nan = warp_sum(nan);
  if (lane == 0) {
    if (below) atomicAdd(&s_below, (unsigned long long)below);
    if (nan) atomicAdd(&s_nan, nan);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_below) atomicAdd(reinterpret_cast<unsigned long long*>(ws + W_BELOW), s_below);
    if (s_nan) atomicAdd(&ws[W_NANF], s_nan);
    __threadfence();
    s_lastf = (atomicAdd(&ws[W_TICKET], 1u) == gridDim.x - 1);
    if (s_lastf) {
      __threadfence();
      const unsigned long long b = __ldcg(reinterpret_cast<unsigned long long*>(ws + W_BELOW));
      const unsigned long long c = __ldcg(reinterpret_cast<unsigned long long*>(ws + W_CAND));
      const uint32_t nn = __ldcg(ws + W_NANF);
      const bool over = __ldcg(ws + W_OVERFLOW) != 0u;
      const unsigned long long nv = (unsigned long long)n - nn;      // non-NaN elements
      bool ok = !over && b <= k && (k + 1 < b + c || (k + 1 >= nv && k < b + c));
      if (nn != 0u) { ok = !over; ws[SG_SELECT_WS_NANCOUNT] = nn; }    // any NaN: the result is NaN either way
      if (ok) {
        ws[W_USECAND] = 1u;
        *reinterpret_cast<unsigned long long*>(ws + W_KREM_LO) = (nn != 0u) ? 0ull : (k - b);
      }
    }
  }
}

// Finds the bin holding rank k_rem among the 256 bins, narrows the prefix, clears the histogram for the
// next pass.  Executed by a whole CTA of >= 256 threads (thread t < 256 owns bin t).
__device__ __forceinline__ void step_body(uint32_t* ws, int pass) {
  __shared__ unsigned long long s_warp[8];
  __shared__ unsigned long long s_before, s_cnt;
  __shared__ int s_bucket;
  __shared__ uint32_t s_next;
  const int t = threadIdx.x;
  const bool on = t < 256;
  const unsigned long long c = on ? __ldcg(ws + SG_SELECT_WS_HIST + t) : 0ull;
  const unsigned long long k = __ldcg(reinterpret_cast<unsigned long long*>(ws + W_KREM_LO));
  unsigned long long x = c;
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (on && lane == 31) s_warp[w] = x;
  if (t == 0) { s_bucket = -1; s_next = 0xFFFFFFFFu; }
  __syncthreads();
  unsigned long long run = x - c;
  if (on) for (int ww = 0; ww < w; ++ww) run += s_warp[ww];
  if (on && k >= run && k < run + c) { s_bucket = t; s_before = run; s_cnt = c; }
  __syncthreads();
  if (on) ws[SG_SELECT_WS_HIST + t] = 0u;
  int b = s_bucket;
  if (on && pass == 3 && b >= 0 && t > b && c != 0ull) atomicMin(&s_next, (uint32_t)t);  // next populated digit in the bucket
  __syncthreads();
  if (t == 0) {
    unsigned long long before = s_before, cnt = s_cnt;
    if (b < 0) { b = 255; before = 0; cnt = 0; }  // k >= n: clamp (callers validate k < n)
    const uint32_t prefix = (pass == 0) ? (uint32_t)b : ((__ldcg(ws + W_PREFIX) << 8) | (uint32_t)b);
    ws[W_PREFIX] = prefix;
    const unsigned long long krem = k - before;
    *reinterpret_cast<unsigned long long*>(ws + W_KREM_LO) = krem;
    if (pass == 3) {
      ws[W_SELKEY] = prefix;
      ws[W_NEEDNEXT] = (krem + 1 >= cnt) ? 1u : 0u;
      ws[W_NEXTIN] = (s_next == 0xFFFFFFFFu) ? 0xFFFFFFFFu : ((prefix & ~0xFFu) | s_next);
    }
  }
}

__global__ void __launch_bounds__(256) step_kernel(uint32_t* ws, int pass) { step_body(ws, pass); }

// ---- multi-GPU: the all-reduce of a pass fused into its bucket step, over NVLink peer memory ---------------------------
// Every rank owns a small buffer that all ranks have mapped (CUDA IPC).  One-shot, low-latency protocol: thread t PUSHES
// its word t -- tagged with the call's sequence number in the upper half of ONE 64-bit store -- into slot
// [seq & 1][source rank][t] of EVERY rank's buffer (remote stores are fire-and-forget), then polls its OWN buffer until
// the R tagged words of this call have landed and reduces them in rank order.  No flag, no fence (a 64-bit store is
// single-copy atomic), no NCCL launch: ~one NVLink store latency per pass instead of a collective launch (5 collectives
// cost 0.27 ms per strain in round 1).  Two parity slots suffice: a rank can only start call s+2 after it has read every
// peer's words of call s+1, which each peer writes after it finished reading call s.
// SUM for the 256 digit counters + the NaN counter, MIN for the smallest key above the bucket (last pass).
constexpr unsigned long long kPeerTimeoutNs = 20000000000ull;    // 20 s (NCCL's own watchdog default is minutes)
constexpr int kPeerWords = 258;                                  // ws[0 .. 258): hist[256], nan count, min-above
__host__ __device__ inline size_t peer_slot_words(int nranks) { return (size_t)nranks * 264; }   // 264: 64-byte multiple

__global__ void __launch_bounds__(288) step_peer_kernel(uint32_t* ws, int pass, unsigned long long* const* peers, int rank,
                                                        int nranks, uint32_t seq) {
  const int t = threadIdx.x;
  if (t < kPeerWords) {
    const bool is_min = (t == SG_SELECT_WS_MINABOVE);
    const bool active = !is_min || pass == SG_SELECT_NUM_PASSES - 1;
    if (active) {
      const uint32_t mine = __ldcg(ws + t);
      const unsigned long long tagged = ((unsigned long long)seq << 32) | mine;
      const size_t slot = (size_t)(seq & 1u) * peer_slot_words(nranks);
      for (int r = 0; r < nranks; ++r) {
        volatile unsigned long long* dst = peers[(rank + r) % nranks] + slot + (size_t)rank * 264 + t;
        *dst = tagged;
      }
      const volatile unsigned long long* own = peers[rank] + slot;
      unsigned long long sum = 0ull;
      uint32_t mn = 0xFFFFFFFFu;
      const uint64_t t0 = gtimer_ns();
      for (int r = 0; r < nranks; ++r) {
        unsigned long long w = own[(size_t)r * 264 + t];
        while ((uint32_t)(w >> 32) != seq) {
          // a peer never arrived: report, do not hang.  The bound has to cover the SKEW between ranks, not a latency: in an
          // end-to-end strain every rank reaches its select when its own host-to-device copies are done (the first call of
          // eight ranks that pin their staging buffers at the same time can spread over seconds)
          if (gtimer_ns() - t0 > kPeerTimeoutNs) { ws[W_ERR] = 2u; break; }
          w = own[(size_t)r * 264 + t];
        }
        const uint32_t val = (uint32_t)w;
        sum += val;
        mn = val < mn ? val : mn;
      }
      ws[t] = is_min ? mn : (uint32_t)sum;
    }
  }
  __threadfence_block();
  __syncthreads();
  step_body(ws, pass);
}

__device__ __forceinline__ void finish_body(const uint32_t* ws, float* out2) {
  const float nanv = __uint_as_float(0x7FC00000u);
  if (__ldcg(ws + SG_SELECT_WS_NANCOUNT) != 0u) { out2[0] = nanv; out2[1] = nanv; return; }
  const float a = key_to_float(__ldcg(ws + W_SELKEY));
  float b = a;
  if (__ldcg(ws + W_NEEDNEXT)) {
    const uint32_t in_bucket = __ldcg(ws + W_NEXTIN), above = __ldcg(ws + SG_SELECT_WS_MINABOVE);
    if (in_bucket != 0xFFFFFFFFu) b = key_to_float(in_bucket);
    else if (above != 0xFFFFFFFFu) b = key_to_float(above);
  }
  out2[0] = a;
  out2[1] = b;
}
__global__ void finish_kernel(const uint32_t* ws, float* out2) {
  if (threadIdx.x == 0) finish_body(ws, out2);
}

__global__ void lerp_kernel(const float* stats2, float w, int kind, float* thr) {
  if (threadIdx.x != 0) return;
  const float r = lerp_rule(stats2[0], stats2[1], w, kind);
  thr[0] = r;
}

// Per-segment bitonic sort in shared memory (in-batch quantile: B <= 2048 scores per segment).
__global__ void __launch_bounds__(1024) segment_stats_kernel(const float* __restrict__ v, int seg_len, int pow2, int k,
                                                             int k1, float* __restrict__ out2) {
  extern __shared__ uint32_t s_keys[];
  __shared__ int s_nan;
  const float* src = v + (size_t)blockIdx.x * seg_len;
  if (threadIdx.x == 0) s_nan = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
    uint32_t key = 0xFFFFFFFFu;
    if (i < seg_len) {
      key = float_to_key(src[i]);
      if (key == 0xFFFFFFFFu) s_nan = 1;
    }
    s_keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (pow2 >> 1); i += blockDim.x) {
        const int lo = ((i / stride) * stride * 2) + (i % stride);
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const uint32_t a = s_keys[lo], b = s_keys[hi];
        if ((a > b) == up) { s_keys[lo] = b; s_keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const float nanv = __uint_as_float(0x7FC00000u);
    out2[2 * blockIdx.x] = s_nan ? nanv : key_to_float(s_keys[k]);
    out2[2 * blockIdx.x + 1] = s_nan ? nanv : key_to_float(s_keys[k1]);
  }
}

static int grid_for(int64_t n, int threads, int per_thread) {
  int64_t b = ceil_div(n, (int64_t)threads * per_thread);
  const int64_t cap = (int64_t)state().sm_count * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace sel
}  // namespace sg

extern "C" {

int sg_select_init_attributes() {
  SG_CUDA(cudaFuncSetAttribute(sg::sel::sample_pivot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               sg::sel::kSampleSmem));
  SG_CUDA(cudaFuncSetAttribute(sg::sel::radix_phases_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               sg::sel::kTailSmem));
  return SG_OK;
}

int sg_select_begin(uint32_t* ws, int64_t k, void* stream) {
  SG_READY();
  SG_REQUIRE(ws != nullptr && k >= 0, "ws/k");
  sg::sel::begin_kernel<<<1, 512, 0, sg::as_stream(stream)>>>(ws, (unsigned long long)k);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_select_hist(const float* v, int64_t n, uint32_t* ws, int pass, void* stream) {
  SG_READY();
  SG_REQUIRE(ws != nullptr && n >= 0 && pass >= 0 && pass < SG_SELECT_NUM_PASSES, "arguments");
  SG_REQUIRE(n == 0 || v != nullptr, "v");
  if (n == 0) return SG_OK;
  const int grid = sg::sel::grid_for(n, 512, 16);
  cudaStream_t st = sg::as_stream(stream);
  switch (pass) {
    case 0: sg::sel::hist_kernel<0><<<grid, 512, 0, st>>>(v, n, ws); break;
    case 1: sg::sel::hist_kernel<1><<<grid, 512, 0, st>>>(v, n, ws); break;
    case 2: sg::sel::hist_kernel<2><<<grid, 512, 0, st>>>(v, n, ws); break;
    default: sg::sel::hist_kernel<3><<<grid, 512, 0, st>>>(v, n, ws); break;
  }
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_select_step(uint32_t* ws, int pass, void* stream) {
  SG_READY();
  SG_REQUIRE(ws != nullptr && pass >= 0 && pass < SG_SELECT_NUM_PASSES, "arguments");
  sg::sel::step_kernel<<<1, 256, 0, sg::as_stream(stream)>>>(ws, pass);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

size_t sg_peer_buffer_bytes(int nranks) { return 2 * sg::sel::peer_slot_words(nranks < 1 ? 1 : nranks) * 8; }

int sg_select_step_peer(uint32_t* ws, int pass, void* const* d_peer_buffers, int rank, int nranks, uint32_t seq,
                        void* stream) {
  SG_READY();
  SG_REQUIRE(ws != nullptr && pass >= 0 && pass < SG_SELECT_NUM_PASSES, "arguments");
  SG_REQUIRE(d_peer_buffers != nullptr && nranks >= 1 && nranks <= 64 && rank >= 0 && rank < nranks && seq != 0, "peer arguments");
  sg::sel::step_peer_kernel<<<1, 288, 0, sg::as_stream(stream)>>>(
      ws, pass, reinterpret_cast<unsigned long long* const*>(d_peer_buffers), rank, nranks, seq);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// ---- peer buffers (CUDA IPC): the ONE allocation the library makes itself -- IPC handles need a cudaMalloc'ed base ----
int sg_peer_alloc(int nranks, void** buffer_out, void* h_handle64_out) {
  SG_READY();
  SG_REQUIRE(buffer_out && h_handle64_out && nranks >= 1 && nranks <= 64, "arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  const size_t bytes = sg_peer_buffer_bytes(nranks);
  SG_CUDA(cudaMalloc(&p, bytes));
  SG_CUDA(cudaMemset(p, 0, bytes));          // sequence number 0 is never used: a zeroed buffer holds no valid word
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    (void)cudaFree(p);
    sg::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return SG_ECUDA;
  }
  memcpy(h_handle64_out, &h, 64);
  *buffer_out = p;
  return SG_OK;
}

int sg_peer_open(const void* h_handle64, void** peer_out) {
  SG_READY();
  SG_REQUIRE(h_handle64 && peer_out, "arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle64, 64);
  SG_CUDA(cudaIpcOpenMemHandle(peer_out, h, cudaIpcMemLazyEnablePeerAccess));
  return SG_OK;
}

int sg_peer_close(void* peer) {
  if (peer) SG_CUDA(cudaIpcCloseMemHandle(peer));
  return SG_OK;
}

int sg_peer_free(void* buffer) {
  if (buffer) SG_CUDA(cudaFree(buffer));
  return SG_OK;
}

int sg_select_finish(const uint32_t* ws, float* out2, void* stream) {
  SG_READY();
  SG_REQUIRE(ws != nullptr && out2 != nullptr, "arguments");
  sg::sel::finish_kernel<<<1, 32, 0, sg::as_stream(stream)>>>(ws, out2);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// all four radix passes in one cooperative launch over v (or over the candidate buffer when the filter pass
// validated it); `n_work` sizes the grid
static int radix_phases(const float* v, int64_t n, const float* cand, int64_t n_work, uint32_t* ws, float* out2,
                        cudaStream_t st) {
  int64_t g = sg::ceil_div(n_work, 8192);
  if (g > sg::state().sm_count) g = sg::state().sm_count;
  if (g < 1) g = 1;
  void* args[] = {(void*)&v, (void*)&n, (void*)&cand, (void*)&ws, (void*)&out2};
  SG_CUDA(cudaLaunchCooperativeKernel((const void*)sg::sel::radix_phases_kernel, dim3((unsigned)g), dim3(sg::sel::kTailThreads),
                                      args, sg::sel::kTailSmem, st));
  sg::count_launch();
  return SG_OK;
}

int sg_radix_select(const float* v, int64_t n, int64_t k, uint32_t* ws, float* out2, void* stream) {
  SG_READY();
  SG_REQUIRE(n > 0 && k >= 0 && k < n, "need 0 <= k < n");
  SG_REQUIRE(v && ws && out2, "null pointer");
  int r = sg_select_begin(ws, k, stream);
  if (r != SG_OK) return r;
  return radix_phases(v, n, nullptr, n, ws, out2, sg::as_stream(stream));
}

size_t sg_select_workspace_bytes(int64_t n) {
  size_t b = SG_SELECT_WS_WORDS * 4;
  if (n >= SG_SELECT_ONEPASS_MIN) b += sg::sel::kSampleCount * 4 + sg::align_up((size_t)(n / 16) * 4, 256);
  return b;
}

int sg_select_kth(const float* v, int64_t n, int64_t k, void* workspace, size_t workspace_bytes, float* out2,
                  void* stream) {
  using namespace sg::sel;
  SG_READY();
  SG_REQUIRE(n > 0 && k >= 0 && k < n, "need 0 <= k < n");
  SG_REQUIRE(v && workspace && out2, "null pointer");
  SG_REQUIRE(workspace_bytes >= SG_SELECT_WS_WORDS * 4 && ((uintptr_t)workspace & 15) == 0, "workspace");
  uint32_t* ws = static_cast<uint32_t*>(workspace);
  cudaStream_t st = sg::as_stream(stream);
  const size_t fixed = SG_SELECT_WS_WORDS * 4 + kSampleCount * 4;
  if (n < SG_SELECT_ONEPASS_MIN || workspace_bytes < fixed + (size_t)(n / 16) * 4)
    return sg_radix_select(v, n, k, ws, out2, stream);
  uint32_t* skeys = ws + SG_SELECT_WS_WORDS;
  float* cand = reinterpret_cast<float*>(skeys + kSampleCount);
  const unsigned long long cap = (workspace_bytes - fixed) / 4;
  // the sampling ticket lives in ws itself: every call leaves ws[W_TICKET_SAMPLE] == 0 (the last CTA clears ws),
  // but a fresh workspace holds garbage -> clear it once per call, stream ordered
  SG_CUDA(cudaMemsetAsync(ws + W_TICKET_SAMPLE, 0, 4, st));
  // pivot ranks inside the sample: the ranks of x_(k), x_(k+1) scaled to the sample -+ 5.5 sigma (+ slack)
  const double S = kSampleCount, p = (double)k / (double)n;
  const double sigma = sqrt(S * p * (1.0 - p));
  const int delta = (int)ceil(4.5 * sigma) + 16;
  const int r_lo = (int)floor(p * S) - delta;
  const int r_hi = (int)ceil((double)(k + 1) / (double)n * S) + delta;
  sample_pivot_kernel<<<kSampleCount / kSampleThreads, kSampleThreads, kSampleSmem, st>>>(
      v, n, ws, skeys, ws + W_TICKET_SAMPLE, (unsigned long long)k, r_lo, r_hi);
  SG_LAUNCH_CHECK();
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(sg::state().sm_count * 4));
    cfg.blockDim = dim3(kFilterThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const unsigned long long ku = (unsigned long long)k;
    SG_CUDA(cudaLaunchKernelEx(&cfg, filter_kernel, v, n, cand, cap, ws, ku));
    sg::count_launch();
  }
  return radix_phases(v, n, cand, n / 32, ws, out2, st);
}

int sg_select_check(const void* workspace, void* stream) {
  SG_READY();
  SG_REQUIRE(workspace != nullptr, "workspace");
  uint32_t flag = 0;
  const uint32_t* ws = static_cast<const uint32_t*>(workspace);
  SG_CUDA(cudaMemcpyAsync(&flag, ws + sg::sel::W_ERR, 4, cudaMemcpyDeviceToHost, sg::as_stream(stream)));
  SG_CUDA(cudaStreamSynchronize(sg::as_stream(stream)));
  if (flag != 0) {
    sg::set_error(flag == 2 ? "radix select: a peer rank never arrived in the NVLink all-reduce of a pass (20 s); the order "
                              "statistics of that call are invalid"
                            : "radix select: a grid barrier of the cooperative kernel timed out (device time-sliced?); the "
                              "order statistics of that call were returned as NaN");
    return SG_ECUDA;
  }
  return SG_OK;
}

int sg_lerp_threshold(const float* stats2, float weight, int lerp_kind, float* thr, void* stream) {
  SG_READY();
  SG_REQUIRE(stats2 && thr, "null pointer");
  SG_REQUIRE(lerp_kind == SG_LERP_NUMPY || lerp_kind == SG_LERP_TORCH, "lerp_kind");
  sg::sel::lerp_kernel<<<1, 32, 0, sg::as_stream(stream)>>>(stats2, weight, lerp_kind, thr);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_segment_order_stats(const float* v, int64_t segments, int seg_len, int k, int k1, float* out2, void* stream) {
  SG_READY();
  SG_REQUIRE(v && out2, "null pointer");
  SG_REQUIRE(segments >= 1 && segments <= 0x7FFFFFFF, "segments");
  SG_REQUIRE(seg_len >= 1 && seg_len <= 2048, "seg_len must be in [1, 2048]");
  SG_REQUIRE(k >= 0 && k < seg_len && k1 >= k && k1 < seg_len, "need 0 <= k <= k1 < seg_len");
  int pow2 = 2;
  while (pow2 < seg_len) pow2 <<= 1;
  int threads = pow2 / 2;
  if (threads < 32) threads = 32;
  if (threads > 1024) threads = 1024;
  sg::sel::segment_stats_kernel<<<(unsigned)segments, threads, pow2 * sizeof(uint32_t), sg::as_stream(stream)>>>(
      v, seg_len, pow2, k, k1, out2);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
