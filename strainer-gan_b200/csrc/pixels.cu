// uint8 pixels -> the fp32 NCHW tensor the reference's DataLoader yields:
//   transforms.ToTensor()            img.to(float32).div(255)                ("#strainer gan.py:89")
//   transforms.Normalize(mean, std)  tensor.sub_(mean[c]).div_(std[c])       ("#strainer gan.py:90")
// Both steps are correctly rounded fp32 operations, so the result is bit-identical to the host transform and the
// dataset can stay uint8 in host memory / HBM (12 288 B per 64x64 RGB sample instead of 49 152 B: 4x less PCIe
// and 4x more resident samples).  HBM bound: 1 B read + 4 B written per element.
//
// A pixel has 256 possible values per channel: every CTA builds the [channels][256] fp32 table once (two IEEE
// divisions per entry) and the streaming loop is one shared-memory look-up per element.
#include <cuda_fp16.h>

#include "common.cuh"

namespace sg {
namespace pix {

constexpr int kMaxChannels = 4;
struct NormArgs {
  float mean[kMaxChannels];
  float stddev[kMaxChannels];
};

__device__ __forceinline__ float to_tensor_normalize(uint32_t u, float mean, float stddev) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), stddev);
}

__device__ __forceinline__ void build_lut(float* lut, const NormArgs& a, int channels) {
  for (int i = threadIdx.x; i < channels * 256; i += blockDim.x) {
    const int c = i >> 8;
    lut[i] = to_tensor_normalize((uint32_t)(i & 255), a.mean[c], a.stddev[c]);
  }
  __syncthreads();
}

__device__ __forceinline__ void stg_stream_f4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// NCHW uint8 -> NCHW fp32; plane % 4 == 0, x 4-byte and out 16-byte aligned.  One thread = 4 consecutive elements of
// one (image, channel) plane: a warp reads 128 contiguous bytes and writes 512 contiguous bytes per instruction (whole
// cache lines; a 16-element thread tile wrote four half-covered sectors per line and reached 58 % of the HBM peak);
// U loads in flight per thread.
template <int U>
__global__ void __launch_bounds__(256) nchw_kernel(const uint8_t* __restrict__ x, float* __restrict__ out, uint32_t groups,
                                                   uint32_t plane4, uint32_t channels, const NormArgs a) {
  __shared__ float lut[kMaxChannels * 256];
  build_lut(lut, a, channels);
  const uint32_t* x4 = reinterpret_cast<const uint32_t*>(x);
  const uint32_t gsz = gridDim.x * blockDim.x;   // groups + U * gsz < 2^32 (host slices the batch)
  for (uint32_t g0 = blockIdx.x * blockDim.x + threadIdx.x; g0 < groups; g0 += U * gsz) {
    uint32_t q[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (g0 + u * gsz < groups) q[u] = ldg_stream_u32(x4 + g0 + u * gsz);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t g = g0 + u * gsz;
      if (g >= groups) break;
      const float* t = lut + (((g / plane4) % channels) << 8);
      const uint32_t w = q[u];
      stg_stream_f4(out + ((size_t)g << 2), t[w & 255u], t[(w >> 8) & 255u], t[(w >> 16) & 255u], t[w >> 24]);
    }
  }
}

// NHWC uint8 (the PIL / np.asarray layout) -> NCHW fp32, 3 channels, plane % 4 == 0: one thread = 4 pixels (12
// interleaved bytes; the three 32-bit loads of a warp cover 384 contiguous bytes and hit L1 after the first) -> one
// 128-bit store into each of the three planes (512 contiguous bytes per warp and plane).
__global__ void __launch_bounds__(256) nhwc3_kernel(const uint8_t* __restrict__ x, float* __restrict__ out, uint32_t groups,
                                                    uint32_t plane4, const NormArgs a) {
  __shared__ float lut[kMaxChannels * 256];
  build_lut(lut, a, 3);
  const uint32_t* x4 = reinterpret_cast<const uint32_t*>(x);
  const uint32_t gsz = gridDim.x * blockDim.x;
  const size_t plane = (size_t)plane4 << 2;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gsz) {
    const uint32_t* src = x4 + (size_t)3 * g;
    const uint32_t w[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
    const uint32_t n = g / plane4;
    float* o = out + (size_t)n * 3 * plane + ((size_t)(g - n * plane4) << 2);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int b = 3 * p + c;
        v[p] = lut[(c << 8) + ((w[b >> 2] >> ((b & 3) * 8)) & 255u)];
      }
      stg_stream_f4(o + c * plane, v[0], v[1], v[2], v[3]);
    }
  }
}

// any shape / alignment: one element per thread
__global__ void __launch_bounds__(256) generic_kernel(const uint8_t* __restrict__ x, float* __restrict__ out, int64_t count,
                                                      int64_t plane, int channels, int nhwc, const NormArgs a) {
  __shared__ float lut[kMaxChannels * 256];
  build_lut(lut, a, channels);
  const int64_t total = count * channels * plane;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e % plane, nc = e / plane;
    const int c = (int)(nc % channels);
    const int64_t n = nc / channels;
    const int64_t src = nhwc ? (n * plane + p) * channels + c : e;
    out[e] = lut[(c << 8) + x[src]];
  }
}

// fp16 bits -> fp32 (exact): the device half of the host-packed PCIe copy (csrc/host_pack.cpp).  One thread = 8
// elements: one 128-bit load, two 128-bit streaming stores; HBM bound (2 B read + 4 B written per element).
__global__ void __launch_bounds__(256) f16_expand_kernel(const uint4* __restrict__ x, float* __restrict__ out, int64_t groups) {
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    uint4 q;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(x + g));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    float v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 f = __half22float2(h);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
    float* o = out + (g << 3);
    stg_stream_f4(o, v[0], v[1], v[2], v[3]);
    stg_stream_f4(o + 4, v[4], v[5], v[6], v[7]);
  }
}

__global__ void __launch_bounds__(256) f16_expand_tail_kernel(const uint16_t* __restrict__ x, float* __restrict__ out, int64_t n) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = __half2float(__ushort_as_half(x[e]));
}

}  // namespace pix
}  // namespace sg

extern "C" int sg_u8_normalize(const uint8_t* x, int64_t count, int channels, int64_t plane, int layout,
                               const float* h_mean, const float* h_std, float* out, void* stream) {
  using namespace sg::pix;
  SG_READY();
  SG_REQUIRE(count >= 0 && plane > 0, "count / plane");
  SG_REQUIRE(channels >= 1 && channels <= kMaxChannels, "1..4 channels");
  SG_REQUIRE(layout == SG_LAYOUT_NCHW || layout == SG_LAYOUT_NHWC, "layout");
  SG_REQUIRE(h_mean != nullptr && h_std != nullptr, "mean / std (host arrays of `channels` floats)");
  if (count == 0) return SG_OK;
  SG_REQUIRE(x != nullptr && out != nullptr, "null pointer");
  NormArgs a;
  for (int c = 0; c < kMaxChannels; ++c) {
    a.mean[c] = c < channels ? h_mean[c] : 0.f;
    a.stddev[c] = c < channels ? h_std[c] : 1.f;
  }
  cudaStream_t st = sg::as_stream(stream);
  const int maxb = sg::state().sm_count * 8;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                       plane % 4 == 0 && plane / 4 < (1ll << 31);
  const bool planar = layout == SG_LAYOUT_NCHW || channels == 1;
  if (aligned && (planar || channels == 3)) {
    // 32-bit index arithmetic in the kernels: slices of at most 2^30 four-element groups
    const int64_t per_image = (planar ? channels : 1) * (plane / 4);
    const int64_t slice = (1ll << 30) / per_image;
    SG_REQUIRE(slice >= 1, "image too large");
    for (int64_t i = 0; i < count; i += slice) {
      const int64_t c = count - i < slice ? count - i : slice;
      const int64_t groups = c * per_image;
      const uint8_t* xs = x + i * channels * plane;
      float* os = out + i * channels * plane;
      if (planar) {
        constexpr int U = 8;
        int64_t blocks = sg::ceil_div(groups, 256 * U);
        if (blocks > maxb) blocks = maxb;
        nchw_kernel<U><<<(int)blocks, 256, 0, st>>>(xs, os, (uint32_t)groups, (uint32_t)(plane / 4), (uint32_t)channels, a);
      } else {
        int64_t blocks = sg::ceil_div(groups, 256);
        if (blocks > maxb) blocks = maxb;
        nhwc3_kernel<<<(int)blocks, 256, 0, st>>>(xs, os, (uint32_t)groups, (uint32_t)(plane / 4), a);
      }
    }
  } else {
    int64_t blocks = sg::ceil_div(count * channels * plane, 256);
    if (blocks > maxb) blocks = maxb;
    generic_kernel<<<(int)blocks, 256, 0, st>>>(x, out, count, plane, channels, layout == SG_LAYOUT_NHWC, a);
  }
  SG_LAUNCH_CHECK();
  return SG_OK;
}

extern "C" int sg_f16_expand(const uint16_t* x, int64_t count, float* out, void* stream) {
  using namespace sg::pix;
  SG_READY();
  SG_REQUIRE(count >= 0, "count");
  if (count == 0) return SG_OK;
  SG_REQUIRE(x != nullptr && out != nullptr, "null pointer");
  cudaStream_t st = sg::as_stream(stream);
  const int maxb = sg::state().sm_count * 8;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int64_t groups = aligned ? count / 8 : 0;
  if (groups > 0) {
    int64_t blocks = sg::ceil_div(groups, 256);
    if (blocks > maxb) blocks = maxb;
    f16_expand_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), out, groups);
    SG_LAUNCH_CHECK();
  }
  const int64_t rest = count - groups * 8;
  if (rest > 0) {
    int64_t blocks = sg::ceil_div(rest, 256);
    if (blocks > maxb) blocks = maxb;
    f16_expand_tail_kernel<<<(int)blocks, 256, 0, st>>>(x + groups * 8, out + groups * 8, rest);
    SG_LAUNCH_CHECK();
  }
  return SG_OK;
}
