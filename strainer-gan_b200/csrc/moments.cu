// Moments, feature z-score, min/max and numpy-exact uniform histogram.
// Replaces err.mean() + k*err.std() ("#autoencoder.py:320"), the feature z-score
// ("#z_score.py:286-291", "# 1,2,8.py:164-168") and np.histogram ("#strainer gan.py:293").
// Sums are fp64 with a FIXED association order (per-thread serial, fixed smem tree, partials in
// index order), so results do not depend on grid size, scheduling or sharding.
#include "common.cuh"

namespace sg {
namespace mom {

// One CTA per chunk of SG_MOMENT_CHUNK values.  Thread t accumulates elements {1024*q + 4*t + j}
// (q = 0..3, j = 0..3) in that order -- 128-bit coalesced loads -- then a fixed shared-memory tree.
__global__ void __launch_bounds__(256) chunk_moments_kernel(const float* __restrict__ v, int64_t n,
                                                            double* __restrict__ partial) {
  __shared__ double s_s[256], s_q[256];
  const int64_t base = (int64_t)blockIdx.x * SG_MOMENT_CHUNK;
  double s = 0.0, q = 0.0;
  const bool fast = (base + SG_MOMENT_CHUNK <= n) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
  if (fast) {
    float4 x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = ldg_stream4(reinterpret_cast<const float4*>(v + base) + k * 256 + threadIdx.x);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float e[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const double d = (double)e[j]; s += d; q = fma(d, d, q); }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t i = base + k * 1024 + threadIdx.x * 4 + j;
        if (i < n) { const double d = (double)v[i]; s += d; q = fma(d, d, q); }
      }
  }
  s_s[threadIdx.x] = s;
  s_q[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_s[threadIdx.x] += s_s[threadIdx.x + o];
      s_q[threadIdx.x] += s_q[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = s_s[0];
    partial[2 * blockIdx.x + 1] = s_q[0];
  }
}

__global__ void moments_finish_kernel(const double* __restrict__ partial, int64_t chunks, int64_t n, float k,
                                      double* __restrict__ stats, float* __restrict__ thr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0, q = 0.0;
  for (int64_t c = 0; c < chunks; ++c) {
    s += partial[2 * c];
    q += partial[2 * c + 1];
  }
  const double mean = s / (double)n;
  double var = (q - s * mean) / (double)(n - 1);  // unbiased; n == 1 -> NaN as torch.std
  if (var < 0.0) var = 0.0;
  const double sd = sqrt(var);
  if (stats) { stats[0] = mean; stats[1] = sd; }
  if (thr) thr[0] = __fadd_rn((float)mean, __fmul_rn(k, (float)sd));
}

// Column sums of x[n, d]: block b covers rows [b*R, (b+1)*R); thread t covers columns t, t+256, ...
constexpr int kRowsPerBlock = 256;   // minimum rows per CTA
constexpr int kMaxColBlocks = 1024;  // fixed cap: the partition (hence the fp64 sum order) depends on n only
__host__ __device__ inline int64_t col_rows_per_block(int64_t n) {
  int64_t rows = (n + kMaxColBlocks - 1) / kMaxColBlocks;
  if (rows < kRowsPerBlock) rows = kRowsPerBlock;
  return (rows + 1) & ~(int64_t)1;
}
__global__ void __launch_bounds__(256) col_partial_kernel(const float* __restrict__ x, int64_t n, int d,
                                                          double* __restrict__ part) {
  const int64_t rpb = col_rows_per_block(n);
  const int64_t r0 = (int64_t)blockIdx.x * rpb;
  const int64_t r1 = min(r0 + rpb, n);
  if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // thread = (4 adjacent columns, row phase): with d = 512 the 256 threads cover 2 rows per step and
    // keep 8 x 128-bit loads in flight each; the two row phases are combined in a fixed order.
    __shared__ double s_part[2][128][8];
    const int groups = d >> 2;                      // float4 column groups per row
    const int phases = (groups == 128) ? 2 : 1;   // two row phases only in the tuned d = 512 case (their partials are combined below)
    const int cg = threadIdx.x % groups, ph = threadIdx.x / groups;
    for (int c0 = 0; c0 < groups; c0 += 256) {
      const int g = c0 + cg;
      double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
      if (g < groups && ph < phases) {
        int64_t r = r0 + ph;
        for (; r + 7 * phases < r1; r += 8 * phases) {
          float4 a[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) a[u] = ldg_stream4(reinterpret_cast<const float4*>(x + (r + u * phases) * d) + g);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float e[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { const double v = (double)e[j]; s[j] += v; q[j] = fma(v, v, q[j]); }
          }
        }
        for (; r < r1; r += phases) {
          const float4 a = ldg_stream4(reinterpret_cast<const float4*>(x + r * d) + g);
          const float e[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { const double v = (double)e[j]; s[j] += v; q[j] = fma(v, v, q[j]); }
        }
      }
      if (phases == 2) {  // d == 512: combine the two row phases deterministically
        if (ph < 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { s_part[ph][cg][j] = s[j]; s_part[ph][cg][4 + j] = q[j]; }
        }
        __syncthreads();
        if (ph == 0 && g < groups) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            part[((int64_t)blockIdx.x * d + g * 4 + j) * 2] = s_part[0][cg][j] + s_part[1][cg][j];
            part[((int64_t)blockIdx.x * d + g * 4 + j) * 2 + 1] = s_part[0][cg][4 + j] + s_part[1][cg][4 + j];
          }
        }
        __syncthreads();
      } else if (g < groups && ph == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          part[((int64_t)blockIdx.x * d + g * 4 + j) * 2] = s[j];
          part[((int64_t)blockIdx.x * d + g * 4 + j) * 2 + 1] = q[j];
        }
      }
    }
    return;
  }
  for (int c = threadIdx.x; c < d; c += 256) {
    double s = 0.0, q = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
      const double val = (double)x[r * d + c];
      s += val;
      q = fma(val, val, q);
    }
    part[((int64_t)blockIdx.x * d + c) * 2] = s;
    part[((int64_t)blockIdx.x * d + c) * 2 + 1] = q;
  }
}

__global__ void col_finish_kernel(const double* __restrict__ part, int64_t blocks, int64_t n, int d, int ddof,
                                  float eps_add, float* __restrict__ mean, float* __restrict__ denom) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  double s = 0.0, q = 0.0;
  int64_t b = 0;
  for (; b + 8 <= blocks; b += 8) {  // 8 independent loads in flight, summed in block order
    double2 p[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) p[u] = *reinterpret_cast<const double2*>(part + ((b + u) * d + c) * 2);
#pragma unroll
    for (int u = 0; u < 8; ++u) { s += p[u].x; q += p[u].y; }
  }
  for (; b < blocks; ++b) {
    s += part[(b * d + c) * 2];
    q += part[(b * d + c) * 2 + 1];
  }
  const double m = s / (double)n;
  double var = (q - s * m) / (double)(n - ddof);
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  denom[c] = __fadd_rn((float)sqrt(var), eps_add);
}

// out[i] = max_j |(x[i,j] - mean[j]) / denom[j]|; one warp per row, IEEE division, NaN propagates.
__global__ void __launch_bounds__(256) row_max_absz_kernel(const float* __restrict__ x, int64_t n, int d,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ denom, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + row * d;
  float best = 0.f;
  bool has_nan = false;
  const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec && d == 512) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = ldg_stream4(reinterpret_cast<const float4*>(xr + lane * 4 + u * 128));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = lane * 4 + u * 128;
      const float4 m = *reinterpret_cast<const float4*>(mean + c);
      const float4 s = *reinterpret_cast<const float4*>(denom + c);
      const float z0 = fabsf(__fdiv_rn(__fsub_rn(a[u].x, m.x), s.x));
      const float z1 = fabsf(__fdiv_rn(__fsub_rn(a[u].y, m.y), s.y));
      const float z2 = fabsf(__fdiv_rn(__fsub_rn(a[u].z, m.z), s.z));
      const float z3 = fabsf(__fdiv_rn(__fsub_rn(a[u].w, m.w), s.w));
      has_nan |= (z0 != z0) | (z1 != z1) | (z2 != z2) | (z3 != z3);
      best = fmaxf(best, fmaxf(fmaxf(z0, z1), fmaxf(z2, z3)));
    }
  } else if (vec) {
    for (int c = lane * 4; c < d; c += 128) {
      const float4 a = ldg_stream4(reinterpret_cast<const float4*>(xr + c));
      const float4 m = *reinterpret_cast<const float4*>(mean + c);
      const float4 s = *reinterpret_cast<const float4*>(denom + c);
      const float z0 = fabsf(__fdiv_rn(__fsub_rn(a.x, m.x), s.x));
      const float z1 = fabsf(__fdiv_rn(__fsub_rn(a.y, m.y), s.y));
      const float z2 = fabsf(__fdiv_rn(__fsub_rn(a.z, m.z), s.z));
      const float z3 = fabsf(__fdiv_rn(__fsub_rn(a.w, m.w), s.w));
      has_nan |= (z0 != z0) | (z1 != z1) | (z2 != z2) | (z3 != z3);
      best = fmaxf(best, fmaxf(fmaxf(z0, z1), fmaxf(z2, z3)));
    }
  } else {
    for (int c = lane; c < d; c += 32) {
      const float z = fabsf(__fdiv_rn(__fsub_rn(xr[c], mean[c]), denom[c]));
      has_nan |= (z != z);
      best = fmaxf(best, z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    has_nan |= (bool)__shfl_xor_sync(0xffffffffu, (int)has_nan, o);
  }
  if (lane == 0) out[row] = has_nan ? __uint_as_float(0x7FC00000u) : best;
}

// min / max through the order-preserving radix key (NaN -> largest key -> flagged)
__global__ void minmax_init_kernel(uint32_t* kk) { kk[0] = 0xFFFFFFFFu; kk[1] = 0u; kk[2] = 0u; }
__global__ void __launch_bounds__(512) minmax_kernel(const float* __restrict__ v, int64_t n, uint32_t* __restrict__ kk) {
  uint32_t lo = 0xFFFFFFFFu, hi = 0u, nan = 0u;
  stream_f32<4>(v, n, [&](float f, int64_t) {
    if (f != f) { nan = 1u; return; }
    const uint32_t key = __float_as_uint(f) & 0x80000000u ? ~__float_as_uint(f) : (__float_as_uint(f) | 0x80000000u);
    lo = min(lo, key);
    hi = max(hi, key);
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    nan |= __shfl_xor_sync(0xffffffffu, nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&kk[0], lo);
    atomicMax(&kk[1], hi);
    if (nan) atomicOr(&kk[2], 1u);
  }
}
__global__ void minmax_finish_kernel(const uint32_t* kk, float* out) {
  if (kk[2]) { out[0] = out[1] = __uint_as_float(0x7FC00000u); return; }
  const uint32_t a = kk[0], b = kk[1];
  out[0] = __uint_as_float((a & 0x80000000u) ? (a & 0x7FFFFFFFu) : ~a);
  out[1] = __uint_as_float((b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b);
}

// numpy's uniform-bin fast path (lib/_histograms_impl.py): fp32 index arithmetic + edge correction.
// `copies` interleaved private histograms (bin*copies + lane%copies) remove the same-address atomic
// serialisation a skewed distribution causes.
__global__ void __launch_bounds__(512) hist_uniform_kernel(const float* __restrict__ v, int64_t n,
                                                           const float* __restrict__ edges, int bins, int copies,
                                                           unsigned long long* __restrict__ counts) {
  extern __shared__ uint32_t s_cnt[];  // [bins][copies] then edges[bins + 1]
  float* s_edges = reinterpret_cast<float*>(s_cnt + bins * copies);
  for (int i = threadIdx.x; i < bins * copies; i += blockDim.x) s_cnt[i] = 0u;
  for (int i = threadIdx.x; i <= bins; i += blockDim.x) s_edges[i] = edges[i];
  __syncthreads();
  const float first = s_edges[0], last = s_edges[bins];
  // numpy: idx = int(((x - first) / (last - first)) * bins), then +-1 corrected against the edge array.
  // The corrected index only depends on the raw index being within one bin of the edge-true bin, so
  // the division is replaced by a multiplication with bins / (last - first): same final bins (pinned
  // by tests against np.histogram), ~8 instructions less per element.
  const float scale = __fdiv_rn((float)bins, __fsub_rn(last, first));
  const int copy = threadIdx.x & (copies - 1);
  const uint32_t cnt_base = smem_addr_reg(s_cnt) + (uint32_t)copy * 4u;
  // The edge look-ups (two shared-memory loads per element) are only needed when x sits within rounding
  // distance of an edge: the raw position t = (x - first) * scale carries at most ~3 ulp of relative error and
  // an fp32 linspace edge is within 1 ulp(max |edge|) of its ideal value, i.e. `delta` bins in total.  If the
  // fractional part of t is further than delta from 0 and 1 the bin is floor(t) for certain.
  const float delta = 16.f * fmaxf(fabsf(first), fabsf(last)) * 1.1920929e-7f * scale + (float)bins * 1e-6f + 1e-5f;
  const float hi_ok = 1.f - delta;
  stream_f32<4>(v, n, [&](float x, int64_t) {
    if (!(x >= first && x <= last)) return;
    const float t = __fmul_rn(__fsub_rn(x, first), scale);
    int idx = (int)t;
    const float frac = t - (float)idx;
    if (!(frac > delta && frac < hi_ok) || idx >= bins) {
      if (idx >= bins) idx = bins - 1;
      if (x < s_edges[idx]) --idx;
      else if (x >= s_edges[idx + 1] && idx != bins - 1) ++idx;
    }
    red_shared_add(cnt_base + (uint32_t)(idx * copies) * 4u, 1u);   // see smem_addr_reg: the loop is issue bound
  });
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    unsigned long long c = 0;
    for (int k = 0; k < copies; ++k) c += s_cnt[i * copies + k];
    if (c) atomicAdd(&counts[i], c);
  }
}

// out[i] = (float)((v[i] - mean) / std) in fp64, std from the unbiased one of sg_moments_finish (ddof 0: scaled by
// sqrt((n-1)/n), StandardScaler's population std); a zero std leaves the column unscaled like StandardScaler.
__global__ void __launch_bounds__(256) standardize_kernel(const float* __restrict__ v, int64_t n,
                                                          const double* __restrict__ stats, int ddof,
                                                          float* __restrict__ out) {
  const double mean = stats[0];
  double sd = stats[1];
  if (ddof == 0 && n > 1) sd *= sqrt((double)(n - 1) / (double)n);
  if (!(sd > 0.0)) sd = 1.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (float)(((double)v[i] - mean) / sd);
}

}  // namespace mom
}  // namespace sg

extern "C" {

int sg_chunk_moments(const float* v, int64_t n, double* partial, void* stream) {
  SG_READY();
  SG_REQUIRE(n >= 0 && partial, "arguments");
  if (n == 0) return SG_OK;
  SG_REQUIRE(v != nullptr, "v");
  const int64_t chunks = sg::ceil_div(n, SG_MOMENT_CHUNK);
  SG_REQUIRE(chunks <= 0x7FFFFFFF, "n too large");
  sg::mom::chunk_moments_kernel<<<(unsigned)chunks, 256, 0, sg::as_stream(stream)>>>(v, n, partial);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_moments_finish(const double* partial, int64_t chunks, int64_t n, float k, double* stats, float* thr,
                      void* stream) {
  SG_READY();
  SG_REQUIRE(partial && chunks >= 1 && n >= 1, "arguments");
  sg::mom::moments_finish_kernel<<<1, 32, 0, sg::as_stream(stream)>>>(partial, chunks, n, k, stats, thr);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_standardize(const float* v, int64_t n, const double* stats, int ddof, float* out, void* stream) {
  SG_READY();
  SG_REQUIRE(v && stats && out && n >= 1 && (ddof == 0 || ddof == 1), "arguments");
  int64_t blocks = sg::ceil_div(n, 1024);
  const int64_t cap = (int64_t)sg::state().sm_count * 16;
  if (blocks > cap) blocks = cap;
  sg::mom::standardize_kernel<<<(unsigned)blocks, 256, 0, sg::as_stream(stream)>>>(v, n, stats, ddof, out);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

size_t sg_col_moments_workspace_bytes(int64_t n, int d) {
  const int64_t blocks = sg::ceil_div(n > 0 ? n : 1, sg::mom::col_rows_per_block(n > 0 ? n : 1));
  return sg::align_up((size_t)blocks * d * 2 * sizeof(double), 256);
}

int sg_col_moments(const float* x, int64_t n, int d, int ddof, float eps_add, float* mean, float* denom,
                   void* workspace, void* stream) {
  SG_READY();
  SG_REQUIRE(x && mean && denom && workspace, "null pointer");
  SG_REQUIRE(n >= 1 && d >= 1 && (ddof == 0 || ddof == 1), "n/d/ddof");
  const int64_t blocks = sg::ceil_div(n, sg::mom::col_rows_per_block(n));
  SG_REQUIRE(blocks <= 0x7FFFFFFF, "n too large");
  cudaStream_t st = sg::as_stream(stream);
  sg::mom::col_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, d, static_cast<double*>(workspace));
  SG_LAUNCH_CHECK();
  sg::mom::col_finish_kernel<<<(unsigned)sg::ceil_div(d, 128), 128, 0, st>>>(static_cast<const double*>(workspace),
                                                                            blocks, n, d, ddof, eps_add, mean, denom);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_row_max_absz(const float* x, int64_t n, int d, const float* mean, const float* denom, float* out, void* stream) {
  SG_READY();
  SG_REQUIRE(x && mean && denom && out, "null pointer");
  SG_REQUIRE(n >= 1 && d >= 1, "n/d");
  SG_REQUIRE(((uintptr_t)mean & 15) == 0 && ((uintptr_t)denom & 15) == 0, "mean/denom must be 16-byte aligned");
  sg::mom::row_max_absz_kernel<<<(unsigned)sg::ceil_div(n, 8), 256, 0, sg::as_stream(stream)>>>(x, n, d, mean, denom, out);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_minmax(const float* v, int64_t n, float* minmax, void* stream) {
  SG_READY();
  SG_REQUIRE(v && minmax && n >= 1, "arguments");
  // scratch: reuse the output's neighbourhood is not possible (2 floats); use a small static device buffer per call
  // via the caller-visible contract: minmax must have room for 2 floats + 3 uint32 scratch words (5 words).
  uint32_t* kk = reinterpret_cast<uint32_t*>(minmax + 2);
  cudaStream_t st = sg::as_stream(stream);
  sg::mom::minmax_init_kernel<<<1, 1, 0, st>>>(kk);
  int64_t b = sg::ceil_div(n, 512 * 8);
  const int64_t cap = (int64_t)sg::state().sm_count * 4;
  if (b > cap) b = cap;
  sg::mom::minmax_kernel<<<(unsigned)b, 512, 0, st>>>(v, n, kk);
  sg::mom::minmax_finish_kernel<<<1, 1, 0, st>>>(kk, minmax);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_hist_uniform(const float* v, int64_t n, const float* edges, int bins, long long* counts, void* stream) {
  SG_READY();
  SG_REQUIRE(edges && counts && n >= 0, "arguments");
  SG_REQUIRE(bins >= 1 && bins <= 4096, "bins must be in [1, 4096]");
  if (n == 0) return SG_OK;
  SG_REQUIRE(v != nullptr, "v");
  int64_t b = sg::ceil_div(n, 512 * 8);
  const int64_t cap = (int64_t)sg::state().sm_count * 4;
  if (b > cap) b = cap;
  int copies = 32;
  while (copies > 1 && (size_t)bins * copies * 4 > 40 * 1024) copies >>= 1;
  const size_t smem = (size_t)bins * copies * 4 + (size_t)(bins + 1) * 4;
  sg::mom::hist_uniform_kernel<<<(unsigned)b, 512, smem, sg::as_stream(stream)>>>(
      v, n, edges, bins, copies, reinterpret_cast<unsigned long long*>(counts));
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
