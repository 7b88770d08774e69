// Device radix sort of fp32 keys and the 1-D DBSCAN clean ratio built on it.
// BASELINE.json north_star asks for "1-D DBSCAN thresholds on a device sort"; the reference's
// estimate_ratio_dbscan ("# z_score + DBSCAN.py:272-301") only consumes the fraction of non-noise
// points, which in one dimension is fully determined by the sorted order:
//   core(i)   <=>  #{j : |x_j - x_i| <= eps} >= min_samples     (float64 distances, as scikit-learn)
//   noise(i)  <=>  not core(i) and no core point within eps
// LSD radix sort: 4 passes x 8 bits over the order-preserving key (NaN last), stable, with the
// original index as payload.  Per pass: per-tile digit histograms -> exclusive scan in (digit, tile)
// order -> stable scatter (warp-level match ranking, per-warp digit counters).
#include "common.cuh"

namespace sg {
namespace srt {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;  // 2048

__global__ void __launch_bounds__(kThreads) keys_init_kernel(const float* __restrict__ v, int64_t n,
                                                             uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = float_to_key(v[i]);
    idx[i] = (int32_t)i;
  }
}

__global__ void __launch_bounds__(kThreads) hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                        uint32_t* __restrict__ hist, int tiles) {
  __shared__ uint32_t s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int64_t i = base + j * kThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * tiles + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan of hist[256 * tiles] in place (single CTA, sequential 1024-wide sweeps)
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* __restrict__ hist, int64_t total) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t b = 0; b < total; b += 1024) {
    const int64_t i = b + threadIdx.x;
    const uint32_t v = i < total ? hist[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    if (w == 0) {
      uint32_t s = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      s_w[lane] = s;
    }
    __syncthreads();
    const uint32_t incl = x + (w ? s_w[w - 1] : 0u) + s_carry;
    if (i < total) hist[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                           const int32_t* __restrict__ idx_in, int64_t n, int shift,
                                                           const uint32_t* __restrict__ offs, int tiles,
                                                           uint32_t* __restrict__ keys_out, int32_t* __restrict__ idx_out) {
  __shared__ uint32_t s_cnt[kThreads / 32][256];  // per-warp digit counters -> per-warp offsets
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kThreads / 32) * 256; i += kThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  // warp w owns elements [w*256, (w+1)*256) of the tile, visited in 8 rounds of 32 (index order)
  const int64_t wbase = (int64_t)blockIdx.x * kTile + w * (kItems * 32);
  uint32_t key[kItems];
  int32_t pay[kItems];
  uint32_t rank[kItems];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    const bool valid = i < n;
    key[j] = valid ? keys_in[i] : 0xFFFFFFFFu;
    pay[j] = valid ? idx_in[i] : 0;
    const uint32_t d = (key[j] >> shift) & 255u;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0x100u + lane) & vmask;
    const uint32_t before = valid ? s_cnt[w][d] : 0u;
    rank[j] = before + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && (peers & ((1u << lane) - 1u)) == 0u) s_cnt[w][d] = before + __popc(peers);  // leader updates
    __syncwarp();
  }
  __syncthreads();
  // digit d: exclusive scan of the 8 warp counts
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < kThreads / 32; ++ww) {
      const uint32_t c = s_cnt[ww][d];
      s_cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[j] >> shift) & 255u;
      const uint32_t pos = offs[(size_t)d * tiles + blockIdx.x] + s_cnt[w][d] + rank[j];
      keys_out[pos] = key[j];
      idx_out[pos] = pay[j];
    }
  }
}

__global__ void __launch_bounds__(kThreads) keys_to_float_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                                 float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = key_to_float(keys[i]);
}

// neighbourhood [lo, hi] of sorted point i (|s_j - s_i| <= eps in float64) and the core flag
__global__ void __launch_bounds__(kThreads) dbscan_core_kernel(const float* __restrict__ s, int64_t n, double eps,
                                                               int min_samples, int32_t* __restrict__ lo_out,
                                                               int32_t* __restrict__ hi_out, float* __restrict__ core_f) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = (double)s[i];
    int64_t a = 0, b = i;  // first j in [0, i] with x - s[j] <= eps
    while (a < b) {
      const int64_t m = (a + b) >> 1;
      if (x - (double)s[m] <= eps) b = m; else a = m + 1;
    }
    const int64_t lo = a;
    a = i; b = n - 1;      // last j in [i, n-1] with s[j] - x <= eps
    while (a < b) {
      const int64_t m = (a + b + 1) >> 1;
      if ((double)s[m] - x <= eps) a = m; else b = m - 1;
    }
    const int64_t hi = a;
    lo_out[i] = (int32_t)lo;
    hi_out[i] = (int32_t)hi;
    core_f[i] = (hi - lo + 1 >= min_samples) ? 1.0f : 0.0f;
  }
}

// noise(i) = not core and no core position within [lo, hi]; core_pos ascending (ncore entries)
__global__ void __launch_bounds__(kThreads) dbscan_noise_kernel(const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                                                                const float* __restrict__ core_f,
                                                                const int64_t* __restrict__ core_pos,
                                                                const int64_t* __restrict__ ncore_p,
                                                                const int32_t* __restrict__ order, int64_t n,
                                                                uint8_t* __restrict__ noise_out,
                                                                unsigned long long* __restrict__ clean_count) {
  const int64_t ncore = *ncore_p;
  unsigned local = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    bool clean = core_f[i] != 0.0f;
    if (!clean && ncore > 0) {
      int64_t a = 0, b = ncore;  // first core position >= lo[i]
      while (a < b) {
        const int64_t m = (a + b) >> 1;
        if (core_pos[m] >= lo[i]) b = m; else a = m + 1;
      }
      clean = (a < ncore) && (core_pos[a] <= hi[i]);
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
  uint32_t* hist;
  size_t total;
};
static SortWs carve(void* ws, int64_t n) {
  SortWs s;
  uint8_t* p = static_cast<uint8_t*>(ws);
  const size_t nb = align_up((size_t)(n > 0 ? n : 1) * 4, 256);
  const int64_t tiles = ceil_div(n > 0 ? n : 1, kTile);
  s.keys[0] = reinterpret_cast<uint32_t*>(p); p += nb;
  s.keys[1] = reinterpret_cast<uint32_t*>(p); p += nb;
  s.idx[0] = reinterpret_cast<int32_t*>(p); p += nb;
  s.idx[1] = reinterpret_cast<int32_t*>(p); p += nb;
  s.hist = reinterpret_cast<uint32_t*>(p); p += align_up((size_t)tiles * 256 * 4, 256);
  s.total = (size_t)(p - static_cast<uint8_t*>(ws));
  return s;
}

static int grid1d(int64_t n) {
  int64_t b = ceil_div(n, kThreads * 4);
  const int64_t cap = (int64_t)state().sm_count * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// sorts into keys[0]/idx[0] (4 passes ping-pong back to buffer 0)
static int sort_keys(const float* v, int64_t n, SortWs& s, cudaStream_t st) {
  const int tiles = (int)ceil_div(n, kTile);
  keys_init_kernel<<<grid1d(n), kThreads, 0, st>>>(v, n, s.keys[0], s.idx[0]);
  SG_LAUNCH_CHECK();
  for (int pass = 0; pass < 4; ++pass) {
    const int src = pass & 1, dst = src ^ 1;
    hist_kernel<<<tiles, kThreads, 0, st>>>(s.keys[src], n, pass * 8, s.hist, tiles);
    scan_kernel<<<1, 1024, 0, st>>>(s.hist, (int64_t)tiles * 256);
    scatter_kernel<<<tiles, kThreads, 0, st>>>(s.keys[src], s.idx[src], n, pass * 8, s.hist, tiles, s.keys[dst], s.idx[dst]);
    SG_LAUNCH_CHECK();
  }
  return SG_OK;
}

}  // namespace srt
}  // namespace sg

extern "C" {

size_t sg_sort_workspace_bytes(int64_t n) {
  sg::srt::SortWs s = sg::srt::carve(nullptr, n);
  return s.total + 256;
}

int sg_sort_f32(const float* v, int64_t n, float* sorted_out, int32_t* order_out, void* workspace, void* stream) {
  using namespace sg::srt;
  SG_READY();
  SG_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && workspace, "arguments");
  if (n == 0) return SG_OK;
  SG_REQUIRE(v != nullptr, "v");
  SG_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t st = sg::as_stream(stream);
  SortWs s = carve(workspace, n);
  int r = sort_keys(v, n, s, st);
  if (r != SG_OK) return r;
  if (sorted_out) keys_to_float_kernel<<<grid1d(n), kThreads, 0, st>>>(s.keys[0], n, sorted_out);
  if (order_out) SG_CUDA(cudaMemcpyAsync(order_out, s.idx[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  SG_LAUNCH_CHECK();
  return SG_OK;
}

size_t sg_dbscan1d_workspace_bytes(int64_t n) {
  const size_t nn = (size_t)(n > 0 ? n : 1);
  return sg_sort_workspace_bytes(n) + sg::align_up(nn * 4, 256) * 4 + sg::align_up(nn * 8, 256) + 256 +
         sg_compact_workspace_bytes(n);
}

int sg_dbscan1d(const float* v, int64_t n, double eps, int min_samples, int64_t* counts_out, uint8_t* noise_out,
                void* workspace, void* stream) {
  using namespace sg::srt;
  SG_READY();
  SG_REQUIRE(v && counts_out && workspace, "null pointer");
  SG_REQUIRE(n >= 1 && n < ((int64_t)1 << 31), "n");
  SG_REQUIRE(eps >= 0.0 && min_samples >= 1, "eps/min_samples");
  SG_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t st = sg::as_stream(stream);
  SortWs s = carve(workspace, n);
  uint8_t* p = static_cast<uint8_t*>(workspace) + sg::align_up(s.total, 256);
  const size_t nb4 = sg::align_up((size_t)n * 4, 256);
  float* sorted = reinterpret_cast<float*>(p); p += nb4;
  int32_t* lo = reinterpret_cast<int32_t*>(p); p += nb4;
  int32_t* hi = reinterpret_cast<int32_t*>(p); p += nb4;
  float* core_f = reinterpret_cast<float*>(p); p += nb4;
  int64_t* core_pos = reinterpret_cast<int64_t*>(p); p += sg::align_up((size_t)n * 8, 256);
  int64_t* ncore = reinterpret_cast<int64_t*>(p);
  float* half = reinterpret_cast<float*>(p + 8);
  unsigned long long* clean = reinterpret_cast<unsigned long long*>(p + 16);
  p += 256;
  void* cws = p;
  int r = sort_keys(v, n, s, st);
  if (r != SG_OK) return r;
  keys_to_float_kernel<<<grid1d(n), kThreads, 0, st>>>(s.keys[0], n, sorted);
  dbscan_core_kernel<<<grid1d(n), kThreads, 0, st>>>(sorted, n, eps, min_samples, lo, hi, core_f);
  SG_LAUNCH_CHECK();
  const float h = 0.5f;
  SG_CUDA(cudaMemcpyAsync(half, &h, 4, cudaMemcpyHostToDevice, st));
  SG_CUDA(cudaMemsetAsync(clean, 0, 8, st));
  r = sg_compact_indices(core_f, n, half, SG_GT, 0, core_pos, ncore, nullptr, cws, stream);
  if (r != SG_OK) return r;
  dbscan_noise_kernel<<<grid1d(n), kThreads, 0, st>>>(lo, hi, core_f, core_pos, ncore, s.idx[0], n, noise_out, clean);
  SG_LAUNCH_CHECK();
  SG_CUDA(cudaMemcpyAsync(counts_out, clean, 8, cudaMemcpyDeviceToDevice, st));
  return SG_OK;
}

}  // extern "C"
