// Dense layers on tcgen05: the 28x28 configurations at dataset scale.
//
//   MLP discriminator 784-1024-512-256-1 ("Untitled-2.py:79-94", "# 1,2,8.py:110-128"):
//     fp32 rows -> fp16 -> three GEMMs with fused bias + LeakyReLU(0.2) epilogues -> dot + sigmoid + BCE head
//   DCGAN-28 conv discriminator (BASELINE config 1, SURVEY 8d C1 option ii; repo defined, there is no 28x28 DCGAN upstream):
//     Conv 1->64 k4 s2 p1 + LeakyReLU (CUDA cores, K = 16) written straight as the im2col rows of
//     Conv 64->128 k4 s2 p1 + BatchNorm + LeakyReLU = ONE GEMM [B*49, 1024] x [128, 1024]^T with the folded BN in the epilogue,
//     Conv 128->1 k7 + Sigmoid + BCE = a 6272-long dot per image.
//
// One kernel serves every GEMM: gemm16_kernel<BLOCK_N>, D[M, N] = act((A[M, K] . B[N, K]^T) * scale[n] + shift[n]) with fp16
// operands (11-bit significands: the 1e-3 fp32 bar in one tensor pass), fp32 accumulation in TMEM, fp16 output that is the
// next layer's K-major A operand as it stands.  Persistent, warp specialised: warp 0 TMA producer (A 128 x 64 and
// B BLOCK_N x 64 boxes, SWIZZLE_128B, K and M tails zero-filled by TMA), warp 1 tcgen05.mma issuer, warps 2-5 epilogue;
// two accumulators ping-pong in TMEM.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sg {
namespace gemm {

using namespace ptx;

constexpr int kErrBase = 60;
constexpr int kOverflowMagic = 0x46503136;   // status word 1: a non-finite logit left the head (same value as d64.cu)

struct GemmParams {
  int m, n, k_steps;          // k_steps = ceil(K / 64)
  int m_tiles, n_tiles;
  const float* scale;         // [n] or nullptr (= 1)
  const float* shift;         // [n] (bias, or folded BN shift)
  float slope;                // LeakyReLU negative slope; 1 = identity
  __half* out;                // [m][ldo] fp16
  int ldo;
  int* err;
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 2 * BLOCK_N * 4 + 1024;
  static constexpr int kThreads = 192;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1)
gemm16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar0 = base + S * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * S + 4);
  volatile int* s_abort = reinterpret_cast<volatile int*>(bars + 2 * S + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
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
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;   // n fastest: the A tile is shared through L2
        for (int ks = 0; ks < p.k_steps; ++ks) {
          if (!mbar_wait(empty_bar(stage), phase ^ 1u, s_abort, p.err, kErrBase + 1)) { ok = false; break; }
          const uint32_t sa = base + stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_2d(sa, &tmap_a, full_bar(stage), ks * 64, mt * 128);
          tma_load_2d(sa + Cfg::kABytes, &tmap_b, full_bar(stage), ks * 64, nt * BLOCK_N);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_16(128, BLOCK_N, true);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1u, s_abort, p.err, kErrBase + 3)) break;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int ks = 0; ks < p.k_steps; ++ks) {
          if (!mbar_wait(full_bar(stage), phase, s_abort, p.err, kErrBase + 2)) { ok = false; break; }
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
    const int rowl = lg * 32 + lane;
    const float slope = p.slope;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const int row = mt * 128 + rowl;
      const bool valid = row < p.m;
      __half* dst = p.out + (size_t)row * p.ldo + (size_t)nt * BLOCK_N;
      const float* sc = p.scale ? p.scale + nt * BLOCK_N : nullptr;
      const float* sh = p.shift + nt * BLOCK_N;
      if (!mbar_wait(tfull_bar(acc), acc_phase, s_abort, p.err, kErrBase + 4)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 2
      for (int cb = 0; cb < BLOCK_N; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + cb, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 h4 = __ldg(reinterpret_cast<const float4*>(sh + cb) + q);
          float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (sc) s4 = __ldg(reinterpret_cast<const float4*>(sc + cb) + q);
          float a0 = fmaf(__uint_as_float(v[4 * q]), s4.x, h4.x), a1 = fmaf(__uint_as_float(v[4 * q + 1]), s4.y, h4.y);
          float a2 = fmaf(__uint_as_float(v[4 * q + 2]), s4.z, h4.z), a3 = fmaf(__uint_as_float(v[4 * q + 3]), s4.w, h4.w);
          a0 = fmaxf(a0, slope * a0); a1 = fmaxf(a1, slope * a1);
          a2 = fmaxf(a2, slope * a2); a3 = fmaxf(a3, slope * a3);
          const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
          pk[2 * q] = *reinterpret_cast<const uint32_t*>(&h01);
          pk[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h23);
        }
        if (valid) {
          uint4* d = reinterpret_cast<uint4*>(dst + cb);
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
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

// A [m, k] fp16 (row pitch lda elements), B [n, k] fp16 (row pitch ldb); n a multiple of 128; pitches multiples of 8.
static int launch_gemm16(const __half* a, int64_t lda, const __half* b, int64_t ldb, int64_t m, int n, int k,
                         const float* scale, const float* shift, float slope, __half* out, int ldo, int* err,
                         cudaStream_t st) {
  SG_REQUIRE(n % 128 == 0 && (lda & 7) == 0 && (ldb & 7) == 0 && (ldo & 7) == 0, "gemm16: n % 128, pitches % 8");
  SG_REQUIRE(m >= 1 && m <= 0x7FFFFFFF && k >= 1, "gemm16: m / k");
  const int block_n = (n % 256 == 0) ? 256 : 128;
  CUtensorMap ta, tb;
  {
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)m};
    cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
    cuuint32_t box[2] = {64, 128};
    int r = encode_tmap(&ta, 2, a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    if (r != SG_OK) return r;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)ldb * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    int r = encode_tmap(&tb, 2, b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    if (r != SG_OK) return r;
  }
  GemmParams p;
  p.m = (int)m;
  p.n = n;
  p.k_steps = (k + 63) / 64;
  p.m_tiles = (int)ceil_div(m, 128);
  p.n_tiles = n / block_n;
  p.scale = scale;
  p.shift = shift;
  p.slope = slope;
  p.out = out;
  p.ldo = ldo;
  p.err = err;
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < state().sm_count ? tiles : state().sm_count);
  if (block_n == 256) gemm16_kernel<256><<<grid, 192, GemmCfg<256>::kSmemBytes, st>>>(ta, tb, p);
  else gemm16_kernel<128><<<grid, 192, GemmCfg<128>::kSmemBytes, st>>>(ta, tb, p);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// ---- element conversions / small layers around the GEMMs ----------------------------------------------------------
__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float* __restrict__ x, int64_t n, __half* __restrict__ y) {
  const int64_t n8 = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += stride) {
      const float4 a = ldg_stream4(reinterpret_cast<const float4*>(x) + 2 * i);
      const float4 b = ldg_stream4(reinterpret_cast<const float4*>(x) + 2 * i + 1);
      const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      const __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
      reinterpret_cast<uint4*>(y)[i] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                  *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    }
    for (int64_t i = (n8 << 3) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __float2half_rn(x[i]);
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __float2half_rn(x[i]);
  }
}

// logit[b] = <h[b, :k], w> + bias (fp32, fixed order), prob = sigmoid, loss = BCE vs 1; one warp per row, k % 8 == 0
__global__ void __launch_bounds__(256) head16_kernel(const __half* __restrict__ h, int64_t rows, int k, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ logit,
                                                     float* __restrict__ prob, float* __restrict__ loss, int* __restrict__ status) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const __half* hr = h + (size_t)r * k;
  float acc = 0.f;
  for (int c = lane * 8; c < k; c += 256) {
    const uint4 raw = *reinterpret_cast<const uint4*>(hr + c);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
    const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
    float xv[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[q]));
      xv[2 * q] = f.x;
      xv[2 * q + 1] = f.y;
    }
    acc = fmaf(xv[0], w0.x, acc); acc = fmaf(xv[1], w0.y, acc); acc = fmaf(xv[2], w0.z, acc); acc = fmaf(xv[3], w0.w, acc);
    acc = fmaf(xv[4], w1.x, acc); acc = fmaf(xv[5], w1.y, acc); acc = fmaf(xv[6], w1.z, acc); acc = fmaf(xv[7], w1.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += bias ? bias[0] : 0.f;
    if (!(fabsf(acc) <= 3.4e38f) && status) atomicExch(status + 1, kOverflowMagic);   // an activation left fp16's range
    const float pr = 1.0f / (1.0f + expf(-acc));
    if (logit) logit[r] = acc;
    if (prob) prob[r] = pr;
    if (loss) loss[r] = -fmaxf(logf(pr), -100.0f);
  }
}

// DCGAN-28 layer 1 fused with the im2col of layer 2: one CTA per image.  act1 = LeakyReLU(conv 1->64 k4 s2 p1 (x)) is built
// in shared memory ([14][14][64] fp16, 25 KB) and written out as the 49 im2col rows of the stride-2 4x4 conv that follows:
// row (oy, ox), column (kh*4 + kw)*64 + c = act1[2oy-1+kh][2ox-1+kw][c] (zero outside), 2 KB per row, 16-byte stores.
__global__ void __launch_bounds__(256) d28_conv1_im2col_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                               int64_t batch, __half* __restrict__ a2) {
  __shared__ float s_x[30 * 30];                    // zero-padded 28 x 28 input
  __shared__ float s_w[64 * 16];
  __shared__ __align__(16) __half s_a[14 * 14 * 64];
  const int64_t img = blockIdx.x;
  if (img >= batch) return;
  for (int i = threadIdx.x; i < 900; i += 256) {
    const int y = i / 30 - 1, xx = i % 30 - 1;
    s_x[i] = (y >= 0 && y < 28 && xx >= 0 && xx < 28) ? x[img * 784 + y * 28 + xx] : 0.f;
  }
  for (int i = threadIdx.x; i < 1024; i += 256) s_w[i] = w1[i];       // [co][kh*4 + kw]
  __syncthreads();
  for (int i = threadIdx.x; i < 196 * 64; i += 256) {
    const int c = i & 63, px = i >> 6, oy = px / 14, ox = px % 14;
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh)
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) acc = fmaf(s_x[(2 * oy + kh) * 30 + 2 * ox + kw], s_w[c * 16 + kh * 4 + kw], acc);
    acc = acc > 0.f ? acc : 0.2f * acc;
    s_a[px * 64 + c] = __float2half_rn(acc);
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(a2 + (size_t)img * 49 * 1024);
  for (int i = threadIdx.x; i < 49 * 128; i += 256) {        // 128 x 16-byte chunks per im2col row
    const int row = i >> 7, ch = i & 127, tap = ch >> 3, c8 = (ch & 7) * 8;
    const int oy = row / 7, ox = row % 7, iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < 14 && ix >= 0 && ix < 14) v = *reinterpret_cast<const uint4*>(s_a + (iy * 14 + ix) * 64 + c8);
    dst[i] = v;
  }
}

// weights -> fp16; the 7x7 head filter [1][128][7][7] -> fp32 [p = ky*7 + kx][c] (the order of the GEMM output rows)
__global__ void __launch_bounds__(256) pack16_kernel(const float* __restrict__ w, int64_t n, __half* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(w[i]);
}
__global__ void __launch_bounds__(256) d28_pack_kernel(const float* __restrict__ w2, const float* __restrict__ w3,
                                                       const float* __restrict__ g, const float* __restrict__ b,
                                                       const float* __restrict__ m, const float* __restrict__ v, float eps,
                                                       __half* __restrict__ w2p, float* __restrict__ w3p, float* __restrict__ ss) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < 128 * 1024) {          // w2 [co][ci][kh][kw] -> [co][(kh*4 + kw)*64 + ci]
    const int co = i >> 10, k = i & 1023, tap = k >> 6, ci = k & 63;
    w2p[i] = __float2half_rn(w2[(co * 64 + ci) * 16 + tap]);
  }
  if (i < 49 * 128) {            // w3 [0][c][ky][kx] -> [p][c]
    const int p = i >> 7, c = i & 127;
    w3p[i] = w3[c * 49 + p];
  }
  if (i < 128) {                 // eval-mode BatchNorm folded to y = x * scale + shift
    const float sc = g[i] / sqrtf(v[i] + eps);
    ss[i] = sc;
    ss[128 + i] = b[i] - m[i] * sc;
  }
}

}  // namespace gemm
}  // namespace sg

extern "C" {

int sg_gemm_init_attributes() {
  using namespace sg::gemm;
  SG_CUDA(cudaFuncSetAttribute(gemm16_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(gemm16_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::kSmemBytes));
  return SG_OK;
}

// ---- MLP discriminator on the tensor cores ---------------------------------------------------------------------------
static const int kMlpDims[4] = {784, 1024, 512, 256};
static size_t mlp_w_off(int l) {   // fp16 weights of layer l inside the packed block
  size_t o = 0;
  for (int i = 0; i < l; ++i) o += sg::align_up((size_t)kMlpDims[i + 1] * kMlpDims[i] * 2, 1024);
  return o;
}

size_t sg_mlp_tc_packed_bytes(void) { return mlp_w_off(3); }

size_t sg_mlp_tc_workspace_bytes(int64_t max_batch) {
  const size_t b = (size_t)(max_batch < 1 ? 1 : max_batch);
  return 1024 + sg::align_up(b * 784 * 2, 1024) + sg::align_up(b * 1024 * 2, 1024) + sg::align_up(b * 512 * 2, 1024) +
         sg::align_up(b * 256 * 2, 1024);
}

int sg_mlp_tc_pack(const float* const* h_params, void* packed, void* stream) {
  using namespace sg::gemm;
  SG_READY();
  SG_REQUIRE(h_params && packed && ((uintptr_t)packed & 1023) == 0, "h_params / 1024-byte aligned packed block");
  for (int i = 0; i < 8; ++i) SG_REQUIRE(h_params[i] != nullptr, "h_params must hold 8 device pointers (w, b) x 4");
  cudaStream_t st = sg::as_stream(stream);
  for (int l = 0; l < 3; ++l) {
    const int64_t n = (int64_t)kMlpDims[l + 1] * kMlpDims[l];
    pack16_kernel<<<(unsigned)sg::ceil_div(n, 1024), 256, 0, st>>>(h_params[2 * l], n,
                                                                  reinterpret_cast<__half*>(static_cast<uint8_t*>(packed) + mlp_w_off(l)));
    SG_LAUNCH_CHECK();
  }
  return SG_OK;
}

int sg_mlp_score_tc(const float* x, int64_t batch, const float* const* h_params, const void* packed, void* workspace,
                    float* logit, float* prob, float* loss, int32_t* status2, void* stream) {
  using namespace sg::gemm;
  SG_READY();
  SG_REQUIRE(x && h_params && packed && workspace, "null pointer");
  SG_REQUIRE(((uintptr_t)packed & 1023) == 0 && ((uintptr_t)workspace & 1023) == 0, "packed / workspace must be 1024-byte aligned");
  SG_REQUIRE(batch >= 0 && batch <= (1 << 24), "batch out of range");
  for (int i = 0; i < 8; ++i) SG_REQUIRE(h_params[i] != nullptr, "h_params must hold 8 device pointers (w, b) x 4");
  if (batch == 0) return SG_OK;
  cudaStream_t st = sg::as_stream(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* err = status2 ? reinterpret_cast<int*>(status2) : reinterpret_cast<int*>(ws);
  const size_t b = (size_t)batch;
  __half* x16 = reinterpret_cast<__half*>(ws + 1024);
  __half* h1 = reinterpret_cast<__half*>(ws + 1024 + sg::align_up(b * 784 * 2, 1024));
  __half* h2 = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(h1) + sg::align_up(b * 1024 * 2, 1024));
  __half* h3 = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(h2) + sg::align_up(b * 512 * 2, 1024));
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  int64_t blocks = sg::ceil_div(batch * 784 / 8 + 1, 256);
  const int64_t cap = (int64_t)sg::state().sm_count * 16;
  f32_to_f16_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(x, batch * 784, x16);
  SG_LAUNCH_CHECK();
  const __half* in = x16;
  __half* outs[3] = {h1, h2, h3};
  for (int l = 0; l < 3; ++l) {
    int r = launch_gemm16(in, kMlpDims[l], reinterpret_cast<const __half*>(pk + mlp_w_off(l)), kMlpDims[l], batch, kMlpDims[l + 1],
                          kMlpDims[l], nullptr, h_params[2 * l + 1], 0.2f, outs[l], kMlpDims[l + 1], err, st);
    if (r != SG_OK) return r;
    in = outs[l];
  }
  head16_kernel<<<(unsigned)sg::ceil_div(batch, 8), 256, 0, st>>>(h3, batch, 256, h_params[6], h_params[7], logit, prob, loss, err);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// ---- DCGAN-28 conv discriminator -------------------------------------------------------------------------------------
size_t sg_d28_packed_bytes(void) { return 128 * 1024 * 2 + 49 * 128 * 4 + 1024; }

size_t sg_d28_workspace_bytes(int64_t max_batch) {
  const size_t b = (size_t)(max_batch < 1 ? 1 : max_batch);
  return 1024 + sg::align_up(b * 49 * 1024 * 2, 1024) + sg::align_up(b * 49 * 128 * 2, 1024);
}

int sg_d28_pack(const float* w2, const float* w3, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                const float* bn_var, float bn_eps, void* packed, void* stream) {
  using namespace sg::gemm;
  SG_READY();
  SG_REQUIRE(w2 && w3 && bn_gamma && bn_beta && bn_mean && bn_var && packed, "null pointer");
  SG_REQUIRE(((uintptr_t)packed & 1023) == 0, "packed block must be 1024-byte aligned");
  uint8_t* pk = static_cast<uint8_t*>(packed);
  d28_pack_kernel<<<512, 256, 0, sg::as_stream(stream)>>>(w2, w3, bn_gamma, bn_beta, bn_mean, bn_var, bn_eps,
                                                          reinterpret_cast<__half*>(pk), reinterpret_cast<float*>(pk + 128 * 1024 * 2),
                                                          reinterpret_cast<float*>(pk + 128 * 1024 * 2 + 49 * 128 * 4));
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_d28_score(const float* x, int64_t batch, const float* w1, const void* packed, void* workspace, float* logit,
                 float* prob, float* loss, int32_t* status2, void* stream) {
  using namespace sg::gemm;
  SG_READY();
  SG_REQUIRE(x && w1 && packed && workspace, "null pointer");
  SG_REQUIRE(((uintptr_t)packed & 1023) == 0 && ((uintptr_t)workspace & 1023) == 0, "packed / workspace must be 1024-byte aligned");
  SG_REQUIRE(batch >= 0 && batch <= (1 << 21), "batch out of range");
  if (batch == 0) return SG_OK;
  cudaStream_t st = sg::as_stream(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  int* err = status2 ? reinterpret_cast<int*>(status2) : reinterpret_cast<int*>(ws);
  const size_t b = (size_t)batch;
  __half* a2 = reinterpret_cast<__half*>(ws + 1024);
  __half* act2 = reinterpret_cast<__half*>(ws + 1024 + sg::align_up(b * 49 * 1024 * 2, 1024));
  d28_conv1_im2col_kernel<<<(unsigned)batch, 256, 0, st>>>(x, w1, batch, a2);
  SG_LAUNCH_CHECK();
  const float* ss = reinterpret_cast<const float*>(pk + 128 * 1024 * 2 + 49 * 128 * 4);
  int r = launch_gemm16(a2, 1024, reinterpret_cast<const __half*>(pk), 1024, batch * 49, 128, 1024, ss, ss + 128, 0.2f, act2, 128,
                        err, st);
  if (r != SG_OK) return r;
  // conv 128 -> 1 k7 over the 7x7 map = a dot over the image's 49 x 128 GEMM output rows
  head16_kernel<<<(unsigned)sg::ceil_div(batch, 8), 256, 0, st>>>(act2, batch, 49 * 128, reinterpret_cast<const float*>(pk + 128 * 1024 * 2),
                                                                  nullptr, logit, prob, loss, err);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
