// Order statistics and thresholds.
// Replaces np.percentile ("#strainer gan.py:381", "# 종합 loss.py:288-292") and torch.quantile
// ("# 상위 10% 제거해서 fake image에 concate.py:246", "# z_score + DBSCAN.py:323"): the reference
// partitions / fully sorts on the CPU; here a 3-pass MSD radix select (11/11/10 bits) finds x_(k),
// one more pass finds x_(k+1) as the smallest key above it, and a one-thread kernel applies the
// library's interpolation rule with the same rounding steps.  Integer histograms make the result
// independent of summation order, hence bit-identical under any sharding.
#include "common.cuh"

namespace sg {
namespace sel {

// internal workspace words (after the public ones)
constexpr int W_PREFIX = 260;      // selected key bits so far (8, 16, 24, 32 bits)
constexpr int W_KREM_LO = 262;     // remaining rank inside the current bucket (u64)
constexpr int W_SELKEY = 264;      // final key of x_(k)
constexpr int W_NEEDNEXT = 265;    // 1: x_(k+1) has a larger key than x_(k)
constexpr int W_NEXTIN = 266;     // smallest populated last-pass digit above the selected one (or ~0)

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
__global__ void __launch_bounds__(512) hist_kernel(const float* __restrict__ v, int64_t n, uint32_t* __restrict__ ws) {
  __shared__ uint32_t s_hist[256 * 32];
  __shared__ uint32_t s_nan, s_min;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) s_hist[i] = 0u;
  if (threadIdx.x == 0) { s_nan = 0u; s_min = 0xFFFFFFFFu; }
  __syncthreads();
  const uint32_t prefix = ws[W_PREFIX];
  const uint32_t lane = threadIdx.x & 31;
  constexpr int kShift = 24 - 8 * PASS;   // digit position
  uint32_t nan_local = 0, min_local = 0xFFFFFFFFu;
  stream_f32<4>(v, n, [&](float f, int64_t) {
    const uint32_t key = float_to_key(f);
    if constexpr (PASS == 0) {
      nan_local += (key == 0xFFFFFFFFu);
      atomicAdd(&s_hist[((key >> 24) << 5) + lane], 1u);
    } else {
      const uint32_t hi = key >> (kShift + 8);
      if (hi == prefix) atomicAdd(&s_hist[(((key >> kShift) & 255u) << 5) + lane], 1u);
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

// Finds the bin holding rank k_rem among the 256 bins, narrows the prefix, clears the histogram for the
// next pass.  One thread per bin.
__global__ void __launch_bounds__(256) step_kernel(uint32_t* ws, int pass) {
  __shared__ unsigned long long s_warp[8];
  __shared__ unsigned long long s_before, s_cnt;
  __shared__ int s_bucket;
  __shared__ uint32_t s_next;
  const int t = threadIdx.x;
  const unsigned long long c = ws[SG_SELECT_WS_HIST + t];
  const unsigned long long k = *reinterpret_cast<unsigned long long*>(ws + W_KREM_LO);
  unsigned long long x = c;
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[w] = x;
  if (t == 0) { s_bucket = -1; s_next = 0xFFFFFFFFu; }
  __syncthreads();
  unsigned long long run = x - c;
  for (int ww = 0; ww < w; ++ww) run += s_warp[ww];
  if (k >= run && k < run + c) { s_bucket = t; s_before = run; s_cnt = c; }
  __syncthreads();
  ws[SG_SELECT_WS_HIST + t] = 0u;
  int b = s_bucket;
  if (pass == 3 && b >= 0 && t > b && c != 0ull) atomicMin(&s_next, (uint32_t)t);  // next populated digit in the bucket
  __syncthreads();
  if (t == 0) {
    unsigned long long before = s_before, cnt = s_cnt;
    if (b < 0) { b = 255; before = 0; cnt = 0; }  // k >= n: clamp (callers validate k < n)
    const uint32_t prefix = (pass == 0) ? (uint32_t)b : ((ws[W_PREFIX] << 8) | (uint32_t)b);
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

__global__ void finish_kernel(const uint32_t* ws, float* out2) {
  if (threadIdx.x != 0) return;
  const float nanv = __uint_as_float(0x7FC00000u);
  if (ws[SG_SELECT_WS_NANCOUNT] != 0u) { out2[0] = nanv; out2[1] = nanv; return; }
  const float a = key_to_float(ws[W_SELKEY]);
  float b = a;
  if (ws[W_NEEDNEXT]) {
    const uint32_t in_bucket = ws[W_NEXTIN], above = ws[SG_SELECT_WS_MINABOVE];
    if (in_bucket != 0xFFFFFFFFu) b = key_to_float(in_bucket);
    else if (above != 0xFFFFFFFFu) b = key_to_float(above);
  }
  out2[0] = a;
  out2[1] = b;
}

__global__ void lerp_kernel(const float* stats2, float w, int kind, float* thr) {
  if (threadIdx.x != 0) return;
  const float a = stats2[0], b = stats2[1];
  float r;
  if (kind == SG_LERP_NUMPY) {
    // numpy _lerp with fp32 operands: separate roundings, no contraction
    const float d = __fsub_rn(b, a);
    r = __fadd_rn(a, __fmul_rn(d, w));
    if (w >= 0.5f) r = __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, w)));
  } else {
    // torch.lerp: fused forms
    const float d = __fsub_rn(b, a);
    r = (fabsf(w) < 0.5f) ? __fmaf_rn(w, d, a) : __fmaf_rn(-d, __fsub_rn(1.0f, w), b);
  }
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

int sg_select_init_attributes() { return SG_OK; }

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

int sg_select_finish(const uint32_t* ws, float* out2, void* stream) {
  SG_READY();
  SG_REQUIRE(ws != nullptr && out2 != nullptr, "arguments");
  sg::sel::finish_kernel<<<1, 32, 0, sg::as_stream(stream)>>>(ws, out2);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_radix_select(const float* v, int64_t n, int64_t k, uint32_t* ws, float* out2, void* stream) {
  SG_REQUIRE(n > 0 && k >= 0 && k < n, "need 0 <= k < n");
  int r = sg_select_begin(ws, k, stream);
  for (int pass = 0; pass < SG_SELECT_NUM_PASSES && r == SG_OK; ++pass) {
    r = sg_select_hist(v, n, ws, pass, stream);
    if (r == SG_OK) r = sg_select_step(ws, pass, stream);
  }
  if (r == SG_OK) r = sg_select_finish(ws, out2, stream);
  return r;
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
