// Auto-encoder reconstruction-error scoring ("#autoencoder.py:269-291" forward +
// ":315-316" per-sample MSE), SURVEY.md §2.3 K15.
//
// Six small convolutions (channels 3/16/32/64) in fp32 on the CUDA cores: one CTA per (sample,
// output-channel group) keeps the sample's whole input feature map and the group's weights in
// shared memory; a warp covers 32 output pixels of one 4-channel quad, so the weight fetch is a
// 16-byte broadcast and each input fetch feeds 4 FMAs.  The last layer fuses tanh, the squared
// error against the input image and the per-sample mean (fixed-order block reduction).
// Transposed convolutions are evaluated in gather form:
//   out[oc][y][x] = b[oc] + sum_{ic,ky,kx : (y+p-ky) % s == 0, (x+p-kx) % s == 0}
//                   in[ic][(y+p-ky)/s][(x+p-kx)/s] * w[ic][oc][ky][kx]
// NOT part of the release library: compiled only with -DSG_AB_VARIANTS (SG_AB_VARIANTS=1 python strainer-gan_b200/build.py),
// as the plain-fp32 cross-check of the tensor-core pipeline in csrc/ae_tc.cu.
#ifdef SG_AB_VARIANTS
#include "common.cuh"

namespace sg {
namespace ae {

enum Act { kNone = 0, kRelu = 1, kTanhMse = 2 };

struct LayerDesc {
  int cin, cout, k, stride, pad, hin, hout, transposed, act, ocg;  // ocg: output channels per CTA (mult of 4)
};

__host__ __device__ inline size_t layer_smem_bytes(const LayerDesc& d) {
  return ((size_t)d.cin * d.hin * d.hin + (size_t)d.cin * d.k * d.k * d.ocg) * sizeof(float);
}

template <int K, int S, int P, bool T>
__global__ void __launch_bounds__(256) ae_conv_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                      const float* __restrict__ target, float* __restrict__ err,
                                                      LayerDesc d) {
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;
  float* s_w = smem + d.cin * d.hin * d.hin;
  __shared__ double s_red[256];
  const int n = blockIdx.x;
  const int oc0 = blockIdx.y * d.ocg;
  const int hin = d.hin, hout = d.hout, cin = d.cin, ocg = d.ocg;
  const int in_elems = cin * hin * hin;
  const float* gin = in + (size_t)n * in_elems;
  for (int i = threadIdx.x; i < in_elems; i += 256) s_in[i] = gin[i];
  // weights -> [ic][ky][kx][ocl]
  const int wtot = cin * K * K * ocg;
  for (int i = threadIdx.x; i < wtot; i += 256) {
    const int ocl = i % ocg;
    int r = i / ocg;
    const int kx = r % K; r /= K;
    const int ky = r % K;
    const int ic = r / K;
    const int oc = oc0 + ocl;
    float v = 0.f;
    if (oc < d.cout) v = T ? w[((size_t)(ic * d.cout + oc) * K + ky) * K + kx] : w[((size_t)(oc * cin + ic) * K + ky) * K + kx];
    s_w[i] = v;
  }
  __syncthreads();
  const int npix = hout * hout;
  const int quads = ocg >> 2;
  const int items = quads * npix;
  double sq = 0.0;
  for (int it = threadIdx.x; it < items; it += 256) {
    const int quad = it / npix;
    const int pix = it - quad * npix;
    const int oy = pix / hout, ox = pix - oy * hout;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      int iy;
      if (T) {
        const int t = oy + P - ky;
        if (t < 0 || (t % S) != 0) continue;
        iy = t / S;
      } else {
        iy = oy * S - P + ky;
      }
      if (iy < 0 || iy >= hin) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        int ix;
        if (T) {
          const int t = ox + P - kx;
          if (t < 0 || (t % S) != 0) continue;
          ix = t / S;
        } else {
          ix = ox * S - P + kx;
        }
        if (ix < 0 || ix >= hin) continue;
        const float* pi = s_in + iy * hin + ix;
        const float* pw = s_w + ((ky * K + kx) * ocg) + quad * 4;
        const int wstride = K * K * ocg;
        for (int ic = 0; ic < cin; ++ic) {
          const float v = pi[ic * hin * hin];
          const float4 w4 = *reinterpret_cast<const float4*>(pw + ic * wstride);
          a0 = fmaf(v, w4.x, a0); a1 = fmaf(v, w4.y, a1); a2 = fmaf(v, w4.z, a2); a3 = fmaf(v, w4.w, a3);
        }
      }
    }
    const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int oc = oc0 + quad * 4 + j;
      if (oc >= d.cout) continue;
      float v = acc[j] + bias[oc];
      if (d.act == kRelu) v = fmaxf(v, 0.f);
      if (d.act == kTanhMse) {
        v = tanhf(v);
        const float df = v - target[((size_t)n * d.cout + oc) * npix + pix];
        sq += (double)(df * df);
      }
      if (out) out[((size_t)n * d.cout + oc) * npix + pix] = v;
    }
  }
  if (d.act == kTanhMse) {
    s_red[threadIdx.x] = sq;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) err[n] = (float)(s_red[0] / (double)(d.cout * npix));
  }
}

static const LayerDesc kLayers[6] = {
    {3, 16, 3, 2, 1, 64, 32, 0, kRelu, 16},   {16, 32, 3, 2, 1, 32, 16, 0, kRelu, 32},
    {32, 64, 7, 1, 0, 16, 10, 0, kNone, 16},  {64, 32, 7, 1, 0, 10, 16, 1, kRelu, 8},
    {32, 16, 3, 2, 1, 16, 32, 1, kRelu, 16},  {16, 3, 3, 2, 1, 32, 64, 1, kTanhMse, 4},
};
// fp32 elements per sample of the five intermediate activations
static const size_t kActElems[5] = {16 * 32 * 32, 32 * 16 * 16, 64 * 10 * 10, 32 * 16 * 16, 16 * 32 * 32};

}  // namespace ae
}  // namespace sg

extern "C" {

int sg_ae_init_attributes() {
  using namespace sg::ae;
  SG_CUDA(cudaFuncSetAttribute(ae_conv_kernel<3, 2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SG_CUDA(cudaFuncSetAttribute(ae_conv_kernel<7, 1, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SG_CUDA(cudaFuncSetAttribute(ae_conv_kernel<7, 1, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SG_CUDA(cudaFuncSetAttribute(ae_conv_kernel<3, 2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  return SG_OK;
}

size_t sg_ae_workspace_bytes(int64_t max_batch) {
  size_t per = 0;
  for (int i = 0; i < 5; ++i) per += sg::align_up(sg::ae::kActElems[i] * sizeof(float), 256);
  return per * (size_t)(max_batch < 1 ? 1 : max_batch);
}

int sg_ae_score(const float* x, int64_t batch, const float* const* h_params, void* workspace, float* err_out,
                float* recon_out, void* stream) {
  using namespace sg::ae;
  SG_READY();
  SG_REQUIRE(x && h_params && workspace && err_out, "null pointer");
  SG_REQUIRE(batch >= 0 && batch <= 65535 * 16, "batch out of range");
  for (int i = 0; i < 12; ++i) SG_REQUIRE(h_params[i] != nullptr, "h_params must hold 12 device pointers (w, b) x 6");
  if (batch == 0) return SG_OK;
  cudaStream_t st = sg::as_stream(stream);
  float* acts[5];
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  size_t off = 0;
  for (int i = 0; i < 5; ++i) {
    acts[i] = reinterpret_cast<float*>(wsb + off);
    off += sg::align_up(kActElems[i] * sizeof(float), 256) * (size_t)batch;
  }
  for (int l = 0; l < 6; ++l) {
    const LayerDesc d = kLayers[l];
    const float* in = l == 0 ? x : acts[l - 1];
    float* out = l == 5 ? recon_out : acts[l];
    const dim3 grid((unsigned)batch, (unsigned)((d.cout + d.ocg - 1) / d.ocg));
    const size_t smem = layer_smem_bytes(d);
    const float* w = h_params[2 * l];
    const float* b = h_params[2 * l + 1];
    if (d.k == 3 && !d.transposed)
      ae_conv_kernel<3, 2, 1, false><<<grid, 256, smem, st>>>(in, w, b, out, x, err_out, d);
    else if (d.k == 7 && !d.transposed)
      ae_conv_kernel<7, 1, 0, false><<<grid, 256, smem, st>>>(in, w, b, out, x, err_out, d);
    else if (d.k == 7)
      ae_conv_kernel<7, 1, 0, true><<<grid, 256, smem, st>>>(in, w, b, out, x, err_out, d);
    else
      ae_conv_kernel<3, 2, 1, true><<<grid, 256, smem, st>>>(in, w, b, out, x, err_out, d);
    SG_LAUNCH_CHECK();
  }
  return SG_OK;
}

}  // extern "C"

#else
extern "C" int sg_ae_init_attributes() { return 0; }
#endif  // SG_AB_VARIANTS
