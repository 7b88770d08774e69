// MLP discriminator scoring (28x28 path): Linear 784-1024-512-256-1, LeakyReLU(0.2), Sigmoid
// ("Untitled-2.py:79-94"; "# 1,2,8.py:110-128" in eval mode, where Dropout is the identity).
// fp32 on the CUDA cores: a 64x64-tiled SGEMM with fused bias + LeakyReLU per hidden layer and a
// warp-per-row dot + sigmoid + BCE head.  At the reference's batch sizes (B = 64) the whole chain is
// launch/latency bound; fp32 keeps it bit-close to the reference.
#include "common.cuh"

namespace sg {
namespace mlp {

// y[b][o] = act(sum_i x[b][i] * w[o][i] + bias[o]);  x [B,I], w [O,I] (nn.Linear layout), y [B,O]
__global__ void __launch_bounds__(256) linear_lrelu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           int batch, int in_f, int out_f, float slope) {
  __shared__ float sx[16][64 + 1];
  __shared__ float sw[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int b0 = blockIdx.y * 64, o0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < in_f; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      const int b = b0 + r, o = o0 + r, kk = k0 + k;
      sx[k][r] = (b < batch && kk < in_f) ? x[(size_t)b * in_f + kk] : 0.f;
      sw[k][r] = (o < out_f && kk < in_f) ? w[(size_t)o * in_f + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sx[k][ty * 4 + i]; b[i] = sw[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty * 4 + i;
    if (b >= batch) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + tx * 4 + j;
      if (o >= out_f) continue;
      float v = acc[i][j] + bias[o];
      v = v > 0.f ? v : slope * v;
      y[(size_t)b * out_f + o] = v;
    }
  }
}

// Small-batch form (the reference's B = 64; any B <= 128), latency bound: one CTA = 8 output features.  Their weight rows
// (<= 32 KB) are loaded into shared memory once, in one round trip; warp w then takes batch rows w, w + 8, ...: the row
// is read with 128-bit loads straight into registers (all loads of TWO rows in flight together), multiplied against the
// 8 weight rows (conflict-free 128-bit shared-memory reads) and reduced with shuffles.  One dependent L2 round trip per
// pair of rows and no barrier in the loop.  (Earlier forms: weight row in registers + batch rows one by one through L1,
// 60 us per layer at B = 64; x staged through shared memory in 64-wide K chunks, 30 us per layer.)
template <int KV>   // 128-bit words per lane and row: ceil(in_f / 128)
__global__ void __launch_bounds__(256) linear_lrelu_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y,
                                                                 int batch, int in_f, int out_f, float slope) {
  __shared__ float4 sw4[8 * KV * 32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int o0 = blockIdx.x * 8;
  for (int idx = threadIdx.x; idx < 8 * KV * 32; idx += 256) {
    const int ol = idx / (KV * 32), k = 4 * (idx - ol * (KV * 32));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < in_f && o0 + ol < out_f) v = __ldg(reinterpret_cast<const float4*>(w + (size_t)(o0 + ol) * in_f + k));
    sw4[idx] = v;
  }
  __syncthreads();
  float bo[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) bo[o] = o0 + o < out_f ? __ldg(bias + o0 + o) : 0.f;
  for (int b = wp; b < batch; b += 16) {
    const int b1 = b + 8;
    float4 xa[KV], xb[KV];
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      const int k = 4 * (lane + 32 * i);
      xa[i] = k < in_f ? __ldg(reinterpret_cast<const float4*>(x + (size_t)b * in_f + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      xb[i] = (k < in_f && b1 < batch) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)b1 * in_f + k))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float ra = 0.f, rb = 0.f;   // lane o keeps output o of both rows
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        const float4 wv = sw4[o * KV * 32 + lane + 32 * i];
        a0 = fmaf(xa[i].x, wv.x, a0); a0 = fmaf(xa[i].y, wv.y, a0); a0 = fmaf(xa[i].z, wv.z, a0); a0 = fmaf(xa[i].w, wv.w, a0);
        a1 = fmaf(xb[i].x, wv.x, a1); a1 = fmaf(xb[i].y, wv.y, a1); a1 = fmaf(xb[i].z, wv.z, a1); a1 = fmaf(xb[i].w, wv.w, a1);
      }
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, sft);
        a1 += __shfl_xor_sync(0xffffffffu, a1, sft);
      }
      if (lane == o) { ra = a0 + bo[o]; rb = a1 + bo[o]; }
    }
    if (lane < 8 && o0 + lane < out_f) {
      y[(size_t)b * out_f + o0 + lane] = ra > 0.f ? ra : slope * ra;
      if (b1 < batch) y[(size_t)b1 * out_f + o0 + lane] = rb > 0.f ? rb : slope * rb;
    }
  }
}

__global__ void __launch_bounds__(256) mlp_head_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int batch, int in_f,
                                                       float* __restrict__ logit, float* __restrict__ prob,
                                                       float* __restrict__ loss) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= batch) return;
  float acc = 0.f;
  for (int i = lane; i < in_f; i += 32) acc = fmaf(h[(size_t)b * in_f + i], w[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += bias[0];
    const float p = 1.0f / (1.0f + expf(-acc));
    if (logit) logit[b] = acc;
    if (prob) prob[b] = p;
    if (loss) loss[b] = -fmaxf(logf(p), -100.0f);
  }
}

}  // namespace mlp
}  // namespace sg

extern "C" {

size_t sg_mlp_workspace_bytes(int64_t max_batch) {
  return (size_t)(max_batch < 1 ? 1 : max_batch) * (1024 + 512 + 256) * sizeof(float);
}

int sg_mlp_score(const float* x, int64_t batch, const float* const* h_params, void* workspace, float* logit,
                 float* prob, float* loss, void* stream) {
  using namespace sg::mlp;
  SG_READY();
  SG_REQUIRE(x && h_params && workspace, "null pointer");
  SG_REQUIRE(batch >= 0 && batch <= (1 << 22), "batch out of range");
  for (int i = 0; i < 8; ++i) SG_REQUIRE(h_params[i] != nullptr, "h_params must hold 8 device pointers (w, b) x 4");
  if (batch == 0) return SG_OK;
  cudaStream_t st = sg::as_stream(stream);
  float* h1 = static_cast<float*>(workspace);
  float* h2 = h1 + (size_t)batch * 1024;
  float* h3 = h2 + (size_t)batch * 512;
  const int dims[4] = {784, 1024, 512, 256};
  const float* in = x;
  float* outs[3] = {h1, h2, h3};
  for (int l = 0; l < 3; ++l) {
    if (batch <= 128 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {   // 128-bit row loads
      const unsigned g = (unsigned)sg::ceil_div(dims[l + 1], 8);
      if (l == 0) linear_lrelu_small_kernel<7><<<g, 256, 0, st>>>(in, h_params[0], h_params[1], outs[0], (int)batch, 784, 1024, 0.2f);
      else if (l == 1) linear_lrelu_small_kernel<8><<<g, 256, 0, st>>>(in, h_params[2], h_params[3], outs[1], (int)batch, 1024, 512, 0.2f);
      else linear_lrelu_small_kernel<4><<<g, 256, 0, st>>>(in, h_params[4], h_params[5], outs[2], (int)batch, 512, 256, 0.2f);
      SG_LAUNCH_CHECK();
      in = outs[l];
      continue;
    }
    const dim3 grid((unsigned)sg::ceil_div(dims[l + 1], 64), (unsigned)sg::ceil_div(batch, 64));
    linear_lrelu_kernel<<<grid, 256, 0, st>>>(in, h_params[2 * l], h_params[2 * l + 1], outs[l], (int)batch, dims[l],
                                              dims[l + 1], 0.2f);
    SG_LAUNCH_CHECK();
    in = outs[l];
  }
  mlp_head_kernel<<<(unsigned)sg::ceil_div(batch, 8), 256, 0, st>>>(h3, h_params[6], h_params[7], (int)batch, 256, logit,
                                                                    prob, loss);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
