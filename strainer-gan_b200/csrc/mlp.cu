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

// Small-batch form (the reference's B = 64): weight-stationary.  One warp per output feature keeps its weight row in
// registers (in_f / 32 values per lane), streams the <= 128 batch rows (shared by the 8 warps of the CTA through L1)
// and reduces each dot product with shuffles.  out_f / 8 CTAs instead of out_f / 64: the chain is latency bound and
// this form has 8x the CTAs and no shared-memory barriers in its K loop.
template <int KPL>   // ceil(in_f / 32) values per lane
__global__ void __launch_bounds__(256) linear_lrelu_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y,
                                                                 int batch, int in_f, int out_f, float slope) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= out_f) return;
  float wr[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int k = lane + 32 * i;
    wr[i] = k < in_f ? __ldg(w + (size_t)o * in_f + k) : 0.f;
  }
  const float bo = __ldg(bias + o);
  for (int b = 0; b < batch; b += 2) {
    const float* x0 = x + (size_t)b * in_f;
    const bool two = b + 1 < batch;
    const float* x1 = two ? x0 + in_f : x0;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      if (k < in_f) {
        a0 = fmaf(x0[k], wr[i], a0);
        a1 = fmaf(x1[k], wr[i], a1);
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, s);
      a1 += __shfl_xor_sync(0xffffffffu, a1, s);
    }
    if (lane == 0) {
      float v = a0 + bo;
      y[(size_t)b * out_f + o] = v > 0.f ? v : slope * v;
      if (two) {
        v = a1 + bo;
        y[(size_t)(b + 1) * out_f + o] = v > 0.f ? v : slope * v;
      }
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
    if (batch <= 128) {
      const unsigned g = (unsigned)sg::ceil_div(dims[l + 1], 8);
      if (l == 0) linear_lrelu_small_kernel<25><<<g, 256, 0, st>>>(in, h_params[0], h_params[1], outs[0], (int)batch, 784, 1024, 0.2f);
      else if (l == 1) linear_lrelu_small_kernel<32><<<g, 256, 0, st>>>(in, h_params[2], h_params[3], outs[1], (int)batch, 1024, 512, 0.2f);
      else linear_lrelu_small_kernel<16><<<g, 256, 0, st>>>(in, h_params[4], h_params[5], outs[2], (int)batch, 512, 256, 0.2f);
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
