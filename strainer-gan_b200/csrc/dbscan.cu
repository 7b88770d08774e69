// DBSCAN clean ratio of high-dimensional features on the tensor cores.
// Replaces estimate_ratio_dbscan ("# z_score + DBSCAN.py:272-301"): StandardScaler -> DBSCAN(eps, min_samples)
// -> mean(labels != -1) on the [N, 512] feature matrix; scikit-learn does an O(N^2) neighbour search on the host.
//
// labels != -1  <=>  the point is a core point (>= min_samples points within eps, itself included) or lies within
// eps of a core point.  Both questions are thresholded pairwise squared distances
//     d2(i, j) = |z_i|^2 + |z_j|^2 - 2 z_i . z_j,
// i.e. one N x N x D GEMM each, run as a tcgen05 implicit GEMM with the threshold fused into the epilogue (nothing
// of size N^2 is ever written).  fp32-grade dot products from bf16 tensor cores: z = hi + lo (bf16 each), three
// K segments hi.hi + lo.hi + hi.lo as in the fp32-parity conv mode (error ~2^-16 |z_i||z_j|).
//   pass 0: counts[i] = #{j : d2 <= eps^2}                       pass 1: reach[i] = any core j with d2 <= eps^2
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sg {
namespace dbs {

using namespace ptx;

constexpr int kErrBase = 60;

// z = (x - mean) / denom (StandardScaler), split into bf16 hi | lo: zs [n_pad][2 d]; norms[i] = sum z^2 (fp32 of
// the fp64 sum).  Rows >= n are zero with norm +inf (never within eps of anything).
__global__ void __launch_bounds__(256) standardize_split_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int d,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ denom,
                                                                __nv_bfloat16* __restrict__ zs, float* __restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  double acc = 0.0;
  for (int j = lane; j < d; j += 32) {
    float z = 0.f;
    if (row < n) z = (x[row * d + j] - mean[j]) / denom[j];
    const __nv_bfloat16 hi = __float2bfloat16_rn(z);
    const __nv_bfloat16 lo = __float2bfloat16_rn(z - __bfloat162float(hi));
    zs[row * 2 * d + j] = hi;
    zs[row * 2 * d + d + j] = lo;
    acc += (double)z * (double)z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) norms[row] = row < n ? (float)acc : __int_as_float(0x7f800000);
}

// |z_i - z_j|^2 from the stored hi/lo halves (z = hi + lo to 2^-17), accumulated in double: the tie breaker for pairs
// whose GEMM distance lies within the error bound of the threshold (exact duplicates at tiny eps, borderline pairs)
__device__ __noinline__ float exact_d2(const __nv_bfloat16* __restrict__ zs, int64_t i, int64_t j, int d) {
  const __nv_bfloat16* a = zs + i * 2 * d;
  const __nv_bfloat16* b = zs + j * 2 * d;
  double acc = 0.0;
  for (int k = 0; k < d; ++k) {
    const float x = (__bfloat162float(a[k]) + __bfloat162float(a[d + k])) - (__bfloat162float(b[k]) + __bfloat162float(b[d + k]));
    acc += (double)x * (double)x;
  }
  return (float)acc;
}

struct Cfg {
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = 256 * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 4;
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 2 * 256 * 4 + 1024;   // + barriers, column norms x 2
  static constexpr int kThreads = 192;
};

// PASS 0: counts[i] += neighbours in this column tile; PASS 1: reach[i] |= a core neighbour in this column tile.
template <int PASS>
__global__ void __launch_bounds__(192, 1)
pairdist_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* __restrict__ zs,
                const float* __restrict__ norms, int d, float eps2,
                const uint8_t* __restrict__ core, unsigned int* __restrict__ counts, uint8_t* __restrict__ reach,
                int m_tiles, int n_tiles, int* err) {
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  const uint32_t bar0 = base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * S + 4);
  volatile int* s_abort = reinterpret_cast<volatile int*>(bars + 2 * S + 5);
  float* s_nj = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 256);            // [2][256]: per accumulator
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = m_tiles * n_tiles;
  const int k_steps = (d / 64) * 3;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    *s_abort = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
        const int mt = tile / n_tiles, nt = tile % n_tiles;
        for (int ks = 0; ks < k_steps; ++ks) {
          if (!mbar_wait(empty_bar(stage), phase ^ 1u, s_abort, err, kErrBase + 1)) { ok = false; break; }
          const int seg = ks % 3, chunk = ks / 3;
          const int ka = chunk * 64 + (seg == 1 ? d : 0);      // lo . hi
          const int kb = chunk * 64 + (seg == 2 ? d : 0);      // hi . lo
          const uint32_t sa = base + stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_2d(sa, &tmap, full_bar(stage), ka, mt * 128);
          tma_load_2d(sa + Cfg::kABytes, &tmap, full_bar(stage), kb, nt * 256);
          tma_load_2d(sa + Cfg::kABytes + 128 * 128, &tmap, full_bar(stage), kb, nt * 256 + 128);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(256);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
        if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1u, s_abort, err, kErrBase + 3)) break;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
        for (int ks = 0; ks < k_steps; ++ks) {
          if (!mbar_wait(full_bar(stage), phase, s_abort, err, kErrBase + 2)) { ok = false; break; }
          tc_fence_after();
          const uint32_t sa = base + stage * Cfg::kStageBytes;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((ks | k) != 0));
          umma_commit(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (!ok) break;
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int mt = tile / n_tiles, nt = tile % n_tiles;
      const int64_t i = (int64_t)mt * 128 + row;
      const float ni = __ldg(norms + i);
      // column data of this tile: squared norms (and, pass 1, the core flags folded in as +inf for non-core)
      float* nj = s_nj + acc * 256;
      for (int c = row; c < 256; c += 128) {
        float v = __ldg(norms + (int64_t)nt * 256 + c);
        if (PASS == 1 && !core[(int64_t)nt * 256 + c]) v = __int_as_float(0x7f800000);
        nj[c] = v;
      }
      named_bar_sync(1, 128);
      if (!mbar_wait(tfull_bar(acc), acc_phase, s_abort, err, kErrBase + 4)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 256);
      unsigned int cnt = 0;
#pragma unroll 2
      for (int cb = 0; cb < 256; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sn = ni + nj[cb + j];
          float d2 = sn - 2.f * __uint_as_float(v[j]);
          // split-product error bound ~2^-15 |z_i||z_j| <= 2^-16 (|z_i|^2 + |z_j|^2) (+ fp32 rounding of the norms):
          // inside it, decide exactly (+inf norms mark padding / non-core columns: never rechecked)
          if (fabsf(d2 - eps2) <= 3.2e-5f * sn && sn < __int_as_float(0x7f800000))
            d2 = exact_d2(zs, i, (int64_t)nt * 256 + cb + j, d);
          cnt += (d2 <= eps2);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (cnt) {
        if (PASS == 0) atomicAdd(counts + i, cnt);
        else reach[i] = 1;
      }
      named_bar_sync(1, 128);   // nj[acc] is rewritten two tiles later: everyone is done reading it
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// core[i] = counts[i] >= min_samples
__global__ void core_kernel(const unsigned int* __restrict__ counts, int64_t n_pad, int64_t n, int min_samples,
                            uint8_t* __restrict__ core) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_pad) core[i] = (i < n && counts[i] >= (unsigned)min_samples) ? 1 : 0;
}

// out[0] = #core, out[1] = #non-noise (core or within eps of a core point)
__global__ void __launch_bounds__(256) tally_kernel(const uint8_t* __restrict__ core, const uint8_t* __restrict__ reach,
                                                    int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long c = 0, r = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    c += core[i];
    r += (core[i] | reach[i]) ? 1 : 0;
  }
  c = warp_sum(c);
  r = warp_sum(r);
  if ((threadIdx.x & 31) == 0) {
    if (c) atomicAdd(out, c);
    if (r) atomicAdd(out + 1, r);
  }
}

struct Layout {
  size_t flag, zs, norms, counts, core, reach, tally, total;
  int64_t n_pad;
};
static Layout layout(int64_t n, int d) {
  Layout L;
  L.n_pad = (int64_t)align_up((size_t)(n > 0 ? n : 1), 256);
  size_t o = 0;
  L.flag = o; o += 1024;
  L.zs = o; o += align_up((size_t)L.n_pad * 2 * d * 2, 1024);
  L.norms = o; o += align_up((size_t)L.n_pad * 4, 1024);
  L.counts = o; o += align_up((size_t)L.n_pad * 4, 1024);
  L.core = o; o += align_up((size_t)L.n_pad, 1024);
  L.reach = o; o += align_up((size_t)L.n_pad, 1024);
  L.tally = o; o += 1024;
  L.total = o;
  return L;
}

}  // namespace dbs
}  // namespace sg

extern "C" {

int sg_dbscan_init_attributes() {
  using namespace sg::dbs;
  SG_CUDA(cudaFuncSetAttribute(pairdist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(pairdist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  return SG_OK;
}

size_t sg_dbscan_nd_workspace_bytes(int64_t n, int d) { return sg::dbs::layout(n, d).total; }

int sg_dbscan_nd(const float* x, int64_t n, int d, const float* mean, const float* denom, double eps, int min_samples,
                 int64_t* counts_out, void* workspace, void* stream) {
  using namespace sg::dbs;
  SG_READY();
  SG_REQUIRE(x && mean && denom && counts_out && workspace, "null pointer");
  SG_REQUIRE(n >= 1 && n <= (1 << 22), "n out of range");
  SG_REQUIRE(d >= 64 && d % 64 == 0 && d <= 4096, "d must be a multiple of 64 in [64, 4096]");
  SG_REQUIRE(eps > 0 && min_samples >= 1, "eps / min_samples");
  SG_REQUIRE(((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");
  cudaStream_t st = sg::as_stream(stream);
  const Layout L = layout(n, d);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* err = reinterpret_cast<int*>(ws + L.flag);
  __nv_bfloat16* zs = reinterpret_cast<__nv_bfloat16*>(ws + L.zs);
  float* norms = reinterpret_cast<float*>(ws + L.norms);
  unsigned int* counts = reinterpret_cast<unsigned int*>(ws + L.counts);
  uint8_t* core = ws + L.core;
  uint8_t* reach = ws + L.reach;
  unsigned long long* tally = reinterpret_cast<unsigned long long*>(ws + L.tally);
  SG_CUDA(cudaMemsetAsync(ws + L.flag, 0, 1024, st));
  SG_CUDA(cudaMemsetAsync(counts, 0, (size_t)L.n_pad * 4, st));
  SG_CUDA(cudaMemsetAsync(reach, 0, (size_t)L.n_pad, st));
  SG_CUDA(cudaMemsetAsync(tally, 0, 16, st));
  standardize_split_kernel<<<(unsigned)sg::ceil_div(L.n_pad, 8), 256, 0, st>>>(x, n, L.n_pad, d, mean, denom, zs, norms);
  SG_LAUNCH_CHECK();
  CUtensorMap tm;
  {
    cuuint64_t dims[2] = {(cuuint64_t)2 * d, (cuuint64_t)L.n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)2 * d * 2};
    cuuint32_t box[2] = {64, 128};
    int r = sg::encode_tmap(&tm, 2, zs, dims, strides, box);
    if (r != SG_OK) return r;
  }
  const int m_tiles = (int)(L.n_pad / 128), n_tiles = (int)(L.n_pad / 256);
  const int64_t total = (int64_t)m_tiles * n_tiles;
  const int grid = (int)(total < sg::state().sm_count ? total : sg::state().sm_count);
  const float eps2 = (float)(eps * eps);
  pairdist_kernel<0><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tm, zs, norms, d, eps2, nullptr, counts, nullptr, m_tiles,
                                                                 n_tiles, err);
  SG_LAUNCH_CHECK();
  core_kernel<<<(unsigned)sg::ceil_div(L.n_pad, 256), 256, 0, st>>>(counts, L.n_pad, n, min_samples, core);
  SG_LAUNCH_CHECK();
  pairdist_kernel<1><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tm, zs, norms, d, eps2, core, nullptr, reach, m_tiles,
                                                                 n_tiles, err);
  SG_LAUNCH_CHECK();
  tally_kernel<<<sg::state().sm_count, 256, 0, st>>>(core, reach, n, tally);
  SG_LAUNCH_CHECK();
  SG_CUDA(cudaMemcpyAsync(counts_out, tally, 16, cudaMemcpyDeviceToDevice, st));
  return SG_OK;
}

int sg_dbscan_nd_check(const void* workspace, void* stream) {
  SG_READY();
  SG_REQUIRE(workspace != nullptr, "workspace");
  int flag = 0;
  SG_CUDA(cudaMemcpyAsync(&flag, workspace, 4, cudaMemcpyDeviceToHost, sg::as_stream(stream)));
  SG_CUDA(cudaStreamSynchronize(sg::as_stream(stream)));
  if (flag != 0) {
    sg::set_error("pairwise-distance pipeline timed out waiting on an mbarrier (code %d)", flag);
    return SG_ECUDA;
  }
  return SG_OK;
}

}  // extern "C"
