// D64 TRAINING path: Discriminator forward with batch-statistics BatchNorm + the whole backward pass (weight, BatchNorm
// and input gradients) on tcgen05.  Replaces what autograd + cuDNN run for `netD(x)` / `err.backward()` in the D and G
// steps of "#strainer gan.py:586-633" (SURVEY.md 8f item 3).
//
// Everything GEMM-shaped goes through ONE kernel, trgemm_kernel<MODE>: persistent, warp specialised (warp 0 TMA producer,
// warp 1 tcgen05.mma issuer, warps 2-5 TMEM-drain epilogue), 128 x 128 x 64 tiles, fp16 operands, fp32 accumulation in
// TMEM (two accumulator pairs ping-pong).  Operands are never im2col'ed in memory: activations live as zero-bordered NHWC
// fp16 tensors [B][S+2][S+2][C] and every operand tile is one (or two) multi-dimensional TMA boxes of them:
//
//   fprop  Y[(b,oh,ow), co]   = sum_{tap,ci} X[b, 2oh+kh, 2ow+kw, ci] W[co, tap, ci]      A = tap box of X (K-major)
//   dgrad  dX[(b,i,j)_class, ci] = sum_{a,c,co} dY[b, i-a, j-c, co] W[co, ci, ph+2a, pw+2c]   one GEMM per input-pixel
//          parity class (ph, pw): 4 taps each, no multiplications by the zeros of a dilated gradient
//   wgrad  dW[co, (tap,ci)]   = sum_{(b,oh,ow)} dY[b,oh,ow,co] X[b, 2oh+kh, 2ow+kw, ci]     the SAME boxes, consumed as
//          MN-major operands (the contraction runs over the rows of the box), split-K over the batch, fp32 partials
//
// The padded NHWC tensor is addressed by TMA as 5-D (2C, S/2+1, 2, S/2+1, B): (column parity, channel) | column pair |
// row parity | row pair | image, so that the stride-2 tap (kh, kw) of a tile of output pixels is a dense box.
// Gradients are carried in fp16 with a power-of-two loss scale chosen per call from max|dL/dlogit| (head_bwd_prep_kernel)
// and removed when the fp32 results are written.  The raw conv outputs stay fp32: BatchNorm (training mode, deterministic
// two-stage column sums), x-hat and the LeakyReLU gate of the backward pass see unrounded sums.
//
// precision 1 (the fp32-parity arithmetic): every 16-bit tensor carries hi = fp16(x) | lo' = fp16((x - hi) * 2^11) side by
// side, every K step runs three segments -- A_hi.B_hi into one accumulator, A_lo'.B_hi and A_hi.B_lo' into a second one --
// and the epilogue adds the pair with the factor 2^-11: 22 significant bits through the same kernel.
//
// Every kernel is launched with programmatic stream serialisation (griddepcontrol): the ~40 launches of a forward +
// backward pass overlap their prologues with the tail of the kernel before.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sg {
namespace dtr {

using namespace ptx;

constexpr int kErrBase = 80;
constexpr int kNonFiniteMagic = 0x46503136;   // status word 1: a non-finite logit or gradient (the value csrc/d64.cu's head uses)
constexpr int kStageBytes = 32768;            // A 128 x 64 fp16 | B 128 x 64 fp16
constexpr int kStages = 6;
constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
constexpr int kBnBlocks = 148;
constexpr float kSlope = 0.2f;
// split operands: x = hi + lo with lo STORED as fp16((x - hi) * 2^11), so that it stays a normal fp16 number whenever hi is one
// (an unscaled lo of an activation ~0.01 would be subnormal: 6e-6 relative error); the two lo segments accumulate in their own
// TMEM accumulator, which the epilogue adds with the factor 2^-11
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

enum { MODE_FPROP = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };
enum { EPI_RAW16 = 0, EPI_L1PAD = 1, EPI_SCATTER = 2, EPI_F32 = 3 };   // EPI_F32 with splits == 1 also stores the raw conv output

struct TrGemm {
  int epi;
  int classes, m_tiles, n_tiles, k_steps, splits, kps;   // units = classes * m_tiles * n_tiles * splits
  int plain;                   // operands are plain 2-D matrices (layer 1)
  int ow, ohb, bb, tpi;        // row box of a tile: ow x ohb pixels of bb images; tiles per image (M tiles, or K tiles in wgrad)
  int ow_log2, ohb_log2;
  int nchunk;                  // 64-channel chunks on the contraction side (fprop: cin / 64, dgrad: cout / 64)
  int cin;                     // channel pitch per (pixel, column parity) of the activation operand (fprop A, wgrad B)
  int creal;                   // its real channel count (= cin, or cin / 2 when the tensor carries hi | lo halves)
  int nseg;                    // 1, or 3: split-operand arithmetic, segments A_hi.B_hi + A_lo.B_hi + A_hi.B_lo per K step
  int a_lo, b_lo;              // channel offset of the lo half in the A tensor / in the activation B tensor of wgrad
  int batch;
  int m_valid, n_valid;        // rows / columns that exist (EPI_RAW16, EPI_F32); EPI_SCATTER: n_valid only
  int ldo;                     // output pitch in elements
  int out_s;                   // EPI_SCATTER: spatial size of dX (= 2 x dY's)
  void* out;
  const __half* mask_src;      // EPI_SCATTER: zero-bordered activation whose sign gates the gradient (LeakyReLU of layer 1)
  int* err;
};

// Programmatic dependent launch: every kernel of this file is launched with programmatic stream serialisation, lets its
// successor become resident at once (launch_dependents) and waits for its predecessor's completion (and memory flush)
// before it touches global memory -- the ~40 launches of a forward + backward pass overlap their prologues (barrier
// init, TMEM allocation, descriptor prefetch) with the tail of the kernel before.

struct Unit { int cls, mt, nt, split, ks0, ks1; };
__device__ __forceinline__ Unit decode_unit(const TrGemm& p, int u) {
  Unit w;
  w.split = u % p.splits;
  int t = u / p.splits;
  w.nt = t % p.n_tiles;
  t /= p.n_tiles;
  w.mt = t % p.m_tiles;
  w.cls = t / p.m_tiles;
  w.ks0 = w.split * p.kps;
  w.ks1 = min(p.k_steps, w.ks0 + p.kps);
  return w;
}

// MN-major, 128-byte-swizzled operand: 64-element MN atoms LBO = 8 KB apart (one TMA box each), 8-row K groups SBO = 1 KB
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

template <int MODE>
__global__ void __launch_bounds__(192, 1)
trgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TrGemm p) {
  constexpr int S = kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar0 = base + S * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * kStageBytes);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * S + 4);
  volatile int* s_abort = reinterpret_cast<volatile int*>(bars + 2 * S + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.classes * p.m_tiles * p.n_tiles * p.splits;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    *s_abort = 0;
    fence_barrier_init();
  }
  pdl_launch_dependents();
  if (warp == 2) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int u = blockIdx.x; u < total && ok; u += gridDim.x) {
        const Unit w = decode_unit(p, u);
        const int b0 = (w.mt / p.tpi) * p.bb, oh0 = (w.mt % p.tpi) * p.ohb;   // row box of the M tile (fprop, dgrad)
        for (int kss = w.ks0 * p.nseg; kss < w.ks1 * p.nseg; ++kss) {
          const int ks = kss / p.nseg, seg = kss - ks * p.nseg;          // seg 1: A lo x B hi, seg 2: A hi x B lo
          const int aoff = seg == 1 ? p.a_lo : 0, boff = seg == 2 ? p.b_lo : 0;
          if (!mbar_wait(empty_bar(stage), phase ^ 1u, s_abort, p.err, kErrBase + 1)) { ok = false; break; }
          const uint32_t sa = base + stage * kStageBytes, sb = sa + 16384;
          const uint32_t fb = full_bar(stage);
          mbar_arrive_expect_tx(fb, kStageBytes);
          if (MODE == MODE_FPROP) {
            if (p.plain) {
              tma_load_2d(sa, &tmap_a, fb, ks * 64 + aoff, w.mt * 128);
            } else {
              const int tap = ks / p.nchunk, chunk = ks - tap * p.nchunk, kh = tap >> 2, kw = tap & 3;
              tma_load_5d(sa, &tmap_a, fb, (kw & 1) * p.cin + aoff + chunk * 64, kw >> 1, kh & 1, oh0 + (kh >> 1), b0);
            }
            tma_load_2d(sb, &tmap_b, fb, (seg * p.k_steps + ks) * 64, w.nt * 128);      // weights: [hi | hi | lo] along K
          } else if (MODE == MODE_DGRAD) {
            if (p.plain) {
              tma_load_2d(sa, &tmap_a, fb, ks * 64 + aoff, w.mt * 128);
              tma_load_2d(sb, &tmap_b, fb, (seg * p.k_steps + ks) * 64, w.nt * 128);
            } else {
              const int ph = w.cls >> 1, pw = w.cls & 1;
              const int ti = ks / p.nchunk, chunk = ks - ti * p.nchunk, a = ti >> 1, c = ti & 1;
              tma_load_5d(sa, &tmap_a, fb, chunk * 64 + aoff, 2 - pw - c, 0, oh0 + 2 - ph - a, b0);
              const int tap16 = (ph + 2 * a) * 4 + pw + 2 * c;
              tma_load_2d(sb, &tmap_b, fb, ((seg * 16 + tap16) * p.nchunk + chunk) * 64, w.nt * 128);
            }
          } else {
            // wgrad: K tile ks = 64 rows (pixels); both operands MN-major, two 64-wide boxes each
            if (p.plain) {       // 64 real columns each: the second box lies outside the matrix (zero fill)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                tma_load_2d(sa + h * 8192, &tmap_a, fb, h ? (1 << 20) : aoff, ks * 64);
                tma_load_2d(sb + h * 8192, &tmap_b, fb, h ? (1 << 20) : boff, ks * 64);
              }
            } else {
              const int kb0 = (ks / p.tpi) * p.bb, koh0 = (ks % p.tpi) * p.ohb;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                tma_load_5d(sa + h * 8192, &tmap_a, fb, w.mt * 128 + h * 64 + aoff, 1, 0, 1 + koh0, kb0);
                const int n0 = w.nt * 128 + h * 64;
                const int tap = n0 / p.creal, ci0 = n0 - tap * p.creal, kh = tap >> 2, kw = tap & 3;
                tma_load_5d(sb + h * 8192, &tmap_b, fb, (kw & 1) * p.cin + boff + ci0, kw >> 1, kh & 1, koh0 + (kh >> 1), kb0);
              }
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the loop and polls the barriers; elect.sync inside the MMA / commit statements picks the issuing
    // lane.  A stage is four M128 x N128 x K16 MMAs = 256 tensor cycles: a single-lane `if (lane == 0)` issuer (a per-lane
    // serialisation loop around every MMA, ~100 instructions per stage) could not keep up with that.
    {
      constexpr uint32_t idesc = umma_idesc_16(128, 128, true) | (MODE == MODE_WGRAD ? ((1u << 15) | (1u << 16)) : 0u);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      bool ok = true;
      const uint64_t a_desc0 = MODE == MODE_WGRAD ? umma_desc_mn_sw128(base) : umma_desc_sw128(base);
      constexpr uint64_t kBOff = 16384 >> 4, kStep = MODE == MODE_WGRAD ? 128 : 2;
      for (int u = blockIdx.x; u < total && ok; u += gridDim.x) {
        const Unit w = decode_unit(p, u);
        if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1u, s_abort, p.err, kErrBase + 3)) break;
        tc_fence_after();
        // accumulator pair of this unit: columns [0, 128) hi x hi, [128, 256) the two lo segments (split operands only)
        const uint32_t tmem_pair = tmem_base + (uint32_t)(acc * 256);
        uint32_t accum[2] = {0, 0};
        int seg3 = 0;                                    // kss % 3 without a division per stage
        for (int kss = w.ks0 * p.nseg; kss < w.ks1 * p.nseg; ++kss) {
          if (!mbar_wait(full_bar(stage), phase, s_abort, p.err, kErrBase + 2)) { ok = false; break; }
          tc_fence_after();
          const int which = (p.nseg == 3 && seg3 != 0) ? 1 : 0;
          if (++seg3 == 3) seg3 = 0;
          const uint32_t tmem_d = tmem_pair + (uint32_t)(which * 128);
          const uint64_t adesc = a_desc0 + (uint64_t)((uint32_t)(stage * kStageBytes) >> 4), bdesc = adesc + kBOff;
#pragma unroll
          for (int k = 0; k < 4; ++k) { umma_f16_elect(tmem_d, adesc + kStep * k, bdesc + kStep * k, idesc, accum[which]); accum[which] = 1; }
          umma_commit_elect(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (!ok) break;
        umma_commit_elect(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================= epilogue: TMEM lane = tile row =================
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
      const Unit w = decode_unit(p, u);
      bool valid;
      __half* dst16 = nullptr;
      float* dst32 = nullptr;
      const __half* maskp = nullptr;
      if (p.epi == EPI_RAW16) {
        const int64_t row = (int64_t)w.mt * 128 + r;
        valid = row < p.m_valid;
        dst16 = static_cast<__half*>(p.out) + row * p.ldo + w.nt * 128;
      } else if (p.epi == EPI_L1PAD) {           // row = (b, oh, ow) of the 32 x 32 map -> [b][oh + 1][ow + 1][64]
        const int64_t row = (int64_t)w.mt * 128 + r;
        valid = row < p.m_valid;
        const int64_t b = row >> 10;
        const int oh = (int)(row >> 5) & 31, ow = (int)row & 31;
        dst16 = static_cast<__half*>(p.out) + ((b * 34 + oh + 1) * 34 + ow + 1) * p.ldo;
      } else if (p.epi == EPI_SCATTER) {         // row (b, i, j) of parity class (ph, pw) -> pixel (2i + 1 - ph, 2j + 1 - pw)
        const int b0 = (w.mt / p.tpi) * p.bb, oh0 = (w.mt % p.tpi) * p.ohb;
        const int ph = w.cls >> 1, pw = w.cls & 1;
        const int j = r & (p.ow - 1), i = oh0 + ((r >> p.ow_log2) & (p.ohb - 1)), b = b0 + (r >> (p.ow_log2 + p.ohb_log2));
        valid = b < p.batch;
        const int ih = 2 * i + 1 - ph, iw = 2 * j + 1 - pw;
        dst16 = static_cast<__half*>(p.out) + (((int64_t)b * p.out_s + ih) * p.out_s + iw) * p.ldo + w.nt * 128;
        if (p.mask_src)
          maskp = p.mask_src + (((int64_t)b * (p.out_s + 2) + ih + 1) * (p.out_s + 2) + iw + 1) * p.ldo + w.nt * 128;
      } else {                                   // fp32 split-K partial [split][m_valid][ldo]
        const int row = w.mt * 128 + r;
        valid = row < p.m_valid;
        dst32 = static_cast<float*>(p.out) + ((int64_t)w.split * p.m_valid + row) * p.ldo + w.nt * 128;
      }
      const int ncols = min(128, p.n_valid - w.nt * 128);
      if (!mbar_wait(tfull_bar(acc), acc_phase, s_abort, p.err, kErrBase + 4)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 256);
      for (int cb = 0; cb < ncols; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + cb, v);
        if (p.nseg == 3) {          // + 2^-11 x the accumulator of the two lo segments
          uint32_t vl[32];
          tmem_ld_32x32(taddr + 128 + cb, vl);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(fmaf(__uint_as_float(vl[q]), kLoInv, __uint_as_float(v[q])));
        } else {
          tmem_ld_wait();
        }
        if (!valid) continue;
        if (p.epi == EPI_F32) {
          float4* d = reinterpret_cast<float4*>(dst32 + cb);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            d[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                               __uint_as_float(v[4 * q + 3]));
        } else {
          float f[32];
          if (p.epi == EPI_L1PAD) {
#pragma unroll
            for (int q = 0; q < 32; ++q) { const float a = __uint_as_float(v[q]); f[q] = fmaxf(a, kSlope * a); }
          } else if (maskp) {
            uint32_t mk[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(maskp + cb) + q);
              mk[4 * q] = m4.x; mk[4 * q + 1] = m4.y; mk[4 * q + 2] = m4.z; mk[4 * q + 3] = m4.w;
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float2 m = __half22float2(*reinterpret_cast<const __half2*>(&mk[q]));
              f[2 * q] = __uint_as_float(v[2 * q]) * (m.x > 0.f ? 1.f : kSlope);
              f[2 * q + 1] = __uint_as_float(v[2 * q + 1]) * (m.y > 0.f ? 1.f : kSlope);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]);
          }
          uint32_t pk[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const __half2 h = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
            pk[q] = *reinterpret_cast<const uint32_t*>(&h);
          }
          uint4* d = reinterpret_cast<uint4*>(dst16 + cb);
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          if (p.nseg == 3) {      // split operands: lo = fp16((x - hi) * 2^11) goes n_valid channels further
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&pk[q]));
              const __half2 l = __floats2half2_rn((f[2 * q] - hf.x) * kLoScale, (f[2 * q + 1] - hf.y) * kLoScale);
              pk[q] = *reinterpret_cast<const uint32_t*>(&l);
            }
            uint4* dl = reinterpret_cast<uint4*>(dst16 + p.n_valid + cb);
#pragma unroll
            for (int q = 0; q < 4; ++q) dl[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
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
    tmem_dealloc<512>(tmem_base);
  }
}

// ---- weight packing: fp32 [Cout][Cin][4][4] -> fp16 Wf [Cout][tap*Cin + ci] (fprop B) and Wd [Cin][tap*Cout + co] (dgrad B)
struct PackArgs {
  const float* w[5];
  __half *wf1, *wd1, *wf[3], *wd[3];
  float* w5p;
  int nseg;      // 3: every row holds [hi | hi | lo] along K (the B side of A_hi.B_hi + A_lo.B_hi + A_hi.B_lo)
};
__device__ __forceinline__ void put_split(__half* d, size_t seg_stride, int nseg, float v) {
  const __half h = __float2half_rn(v);
  d[0] = h;
  if (nseg == 3) { d[seg_stride] = h; d[2 * seg_stride] = __float2half_rn((v - __half2float(h)) * kLoScale); }
}
__global__ void __launch_bounds__(256) pack_train_kernel(const PackArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int pairs2 = 128 * 64, pairs3 = 256 * 128, pairs4 = 512 * 256, pairs = pairs2 + pairs3 + pairs4;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < 2 * pairs + 4096 + 8192; i += gridDim.x * 256) {
    if (i < 2 * pairs) {
      const bool second = i >= pairs;
      int t = second ? i - pairs : i;
      int l, cin, cout;
      if (t < pairs2) { l = 0; cin = 64; cout = 128; }
      else if (t < pairs2 + pairs3) { l = 1; cin = 128; cout = 256; t -= pairs2; }
      else { l = 2; cin = 256; cout = 512; t -= pairs2 + pairs3; }
      int co, ci;
      if (!second) { co = t / cin; ci = t - co * cin; }        // ci fastest: Wf rows are written coalesced
      else { ci = t / cout; co = t - ci * cout; }              // co fastest: Wd rows
      const float4* src = reinterpret_cast<const float4*>(a.w[l + 1] + ((size_t)co * cin + ci) * 16);
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float4 f = __ldg(src + q); v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w; }
      if (!second) {
        __half* d = a.wf[l] + (size_t)co * 16 * cin * a.nseg + ci;
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) put_split(d + (size_t)tap * cin, (size_t)16 * cin, a.nseg, v[tap]);
      } else {
        __half* d = a.wd[l] + (size_t)ci * 16 * cout * a.nseg + co;
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) put_split(d + (size_t)tap * cout, (size_t)16 * cout, a.nseg, v[tap]);
      }
    } else if (i < 2 * pairs + 4096) {     // layer 1: k = (kh*4 + kw)*4 + c (c == 3: zero), Wf1 [co][k], Wd1 [k][co]
      const int t = i - 2 * pairs, co = t >> 6, k = t & 63, tap = k >> 2, c = k & 3;
      const float v = c < 3 ? a.w[0][(co * 3 + c) * 16 + tap] : 0.f;
      put_split(a.wf1 + co * 64 * a.nseg + k, 64, a.nseg, v);
      put_split(a.wd1 + k * 64 * a.nseg + co, 64, a.nseg, v);
    } else {                               // head filter [1][512][4][4] -> fp32 [p][c]
      const int t = i - 2 * pairs - 4096, pp = t >> 9, c = t & 511;
      a.w5p[t] = a.w[4][c * 16 + pp];
    }
  }
}

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
    f[2 * q] = t.x;
    f[2 * q + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __half2 h = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
    w[q] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// 8 consecutive 16-bit values at p; split operands: x = hi + lo, the lo halves lie lo_off elements further
__device__ __forceinline__ void load_h8(const __half* p, int lo_off, bool split, float (&f)[8]) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
  if (split) {
    float l[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p + lo_off)), l);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(l[j], kLoInv, f[j]);
  }
}
__device__ __forceinline__ void store_h8(__half* p, int lo_off, bool split, const float (&f)[8]) {
  const uint4 h = pack8(f);
  *reinterpret_cast<uint4*>(p) = h;
  if (split) {
    float hf[8], l[8];
    unpack8(h, hf);
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = (f[j] - hf[j]) * kLoScale;
    *reinterpret_cast<uint4*>(p + lo_off) = pack8(l);
  }
}

// ---- layer 1 operand: x fp32 NCHW [B][3][64][64] -> im2col rows [B*1024][64] fp16, k = (kh*4 + kw)*4 + c -------------
__global__ void __launch_bounds__(256) im2col1_kernel(const float* __restrict__ x, int64_t batch, __half* __restrict__ col,
                                                      int* __restrict__ status, int split) {
  pdl_launch_dependents();
  pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) status[2] = 0;      // this call's non-finite flag (read by bn_commit_kernel)
  const int64_t total = batch * 1024 * 4;
  for (int64_t t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int kh = (int)(t & 3);
    const int64_t row = t >> 2;
    const int ow = (int)row & 31, oh = (int)(row >> 5) & 31;
    const int64_t b = row >> 10;
    const int ih = 2 * oh - 1 + kh;
    float v[16];
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int iw = 2 * ow - 1 + kw;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (ih >= 0 && ih < 64 && iw >= 0 && iw < 64) {
        const float* px = x + (b * 3 * 64 + ih) * 64 + iw;
        v0 = __ldg(px); v1 = __ldg(px + 4096); v2 = __ldg(px + 8192);
      }
      v[4 * kw] = v0; v[4 * kw + 1] = v1; v[4 * kw + 2] = v2; v[4 * kw + 3] = 0.f;
    }
    const float (&lo8)[8] = *reinterpret_cast<const float (*)[8]>(v);
    const float (&hi8)[8] = *reinterpret_cast<const float (*)[8]>(v + 8);
    __half* d = col + row * (split ? 128 : 64) + kh * 16;
    store_h8(d, 64, split != 0, lo8);
    store_h8(d + 8, 64, split != 0, hi8);
  }
}

// dx fp32 NCHW = (1 / scale) * col2im(dcol [B*1024][64]): the 2 x 2 taps that touch each input pixel
__global__ void __launch_bounds__(256) col2im1_kernel(const __half* __restrict__ dcol, int64_t batch, const float* __restrict__ scal,
                                                      int split, float* __restrict__ dx) {
  pdl_launch_dependents();
  pdl_wait();
  const float inv = scal[1];
  const int64_t total = batch * 4096;
  for (int64_t t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int iw = (int)t & 63, ih = (int)(t >> 6) & 63;
    const int64_t b = t >> 12;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kh = ((ih + 1) & 1) + 2 * a, oh = (ih + 1 - kh) >> 1;
      if (oh < 0 || oh >= 32) continue;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int kw = ((iw + 1) & 1) + 2 * c, ow = (iw + 1 - kw) >> 1;
        if (ow < 0 || ow >= 32) continue;
        const __half* src = dcol + ((b * 32 + oh) * 32 + ow) * (split ? 128 : 64) + (kh * 4 + kw) * 4;
        for (int part = 0; part <= split; ++part) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src + part * 64));
          const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          const float wgt = part ? kLoInv : 1.f;
          s0 = fmaf(f01.x, wgt, s0); s1 = fmaf(f01.y, wgt, s1); s2 = fmaf(f23.x, wgt, s2);
        }
      }
    }
    float* d = dx + (b * 3 * 64 + ih) * 64 + iw;
    d[0] = s0 * inv; d[4096] = s1 * inv; d[8192] = s2 * inv;
  }
}

// ---- BatchNorm, training mode ------------------------------------------------------------------------------------------
// ss block of a layer: scale | shift | mean | rstd | gamma (512 floats each)
// column sums over rows: forward (sum x, sum x^2) or backward (sum g, sum g * xhat with g = dx * LeakyReLU'(x*scale + shift));
// partial [gridDim.x][2][C], fixed order -> deterministic
template <bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ raw, const __half* __restrict__ dx,
                                                        const float* __restrict__ ss, int64_t rows, int C, int split,
                                                        float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[256][17];
  const int Cp = split ? 2 * C : C;      // pitch of the 16-bit gradient rows (hi | lo)
  const int tpr = C >> 3, rpp = 256 / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float s1[8], s2[8], sc[8], sh[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (BWD) {
    load8(ss + cg * 8, sc); load8(ss + 512 + cg * 8, sh); load8(ss + 1024 + cg * 8, mu); load8(ss + 1536 + cg * 8, rs);
  }
  for (int64_t r = (int64_t)blockIdx.x * rpp + rl; r < rows; r += (int64_t)gridDim.x * rpp) {
    float x[8];
    load8(raw + r * C + cg * 8, x);
    if (!BWD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += x[j]; s2[j] = fmaf(x[j], x[j], s2[j]); }
    } else {
      float g[8];
      load_h8(dx + r * Cp + cg * 8, C, split != 0, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(x[j], sc[j], sh[j]);
        const float gj = z > 0.f ? g[j] : kSlope * g[j];
        s1[j] += gj;
        s2[j] = fmaf(gj, (x[j] - mu[j]) * rs[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s1[j]; red[threadIdx.x][8 + j] = s2[j]; }
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float t = 0.f;
      for (int q = 0; q < rpp; ++q) t += red[q * tpr + cg][j];
      partial[((size_t)blockIdx.x * 2 + (j >> 3)) * C + cg * 8 + (j & 7)] = t;
    }
  }
}

// one warp per channel: lanes stride over the row blocks' partial sums (fixed order -> deterministic)
__device__ __forceinline__ void warp_partial_sums(const float* __restrict__ partial, int blocks, int C, int ch, int lane, double& s,
                                                  double& q) {
  s = 0.0; q = 0.0;
  // eight row blocks per trip, all sixteen loads issued before the first add: the rolled loop paid one L2 round trip per
  // block (14 us for the 512 blocks of layer 2); the order of the adds is unchanged
  for (int b0 = lane; b0 < blocks; b0 += 32 * 8) {
    float a[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = b0 + 32 * u;
      const bool in = b < blocks;
      a[u] = in ? partial[((size_t)b * 2) * C + ch] : 0.f;
      c[u] = in ? partial[((size_t)b * 2 + 1) * C + ch] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { s += (double)a[u]; q += (double)c[u]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
}

__global__ void __launch_bounds__(256) bn_fwd_finalize_kernel(const float* __restrict__ partial, int blocks, int64_t rows, int C,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, float momentum, const float* __restrict__ run_mean,
                                                              const float* __restrict__ run_var, float* __restrict__ pend,
                                                              float* __restrict__ ss) {
  pdl_launch_dependents();
  pdl_wait();
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ch >= C) return;
  double s, q;
  warp_partial_sums(partial, blocks, C, ch, lane, s, q);
  if (lane != 0) return;
  const double mean = s / (double)rows;
  double var = q / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = 1.0f / sqrtf((float)var + eps);
  const float g = gamma[ch];
  ss[ch] = g * rstd;
  ss[512 + ch] = beta[ch] - (float)mean * g * rstd;
  ss[1024 + ch] = (float)mean;
  ss[1536 + ch] = rstd;
  ss[2048 + ch] = g;
  // the new running statistics are parked (pend_mean | pend_var) until the head has shown that the batch stayed finite
  if (run_mean) pend[ch] = (1.f - momentum) * run_mean[ch] + momentum * (float)mean;
  if (run_var) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    pend[512 + ch] = (1.f - momentum) * run_var[ch] + momentum * (float)unbiased;
  }
}

// running_mean / running_var of the three BatchNorm layers become visible only if no logit of this batch was non-finite:
// an overflowing batch (re-scored in split arithmetic by the caller) must update them exactly once
struct CommitArgs { float* dst[6]; };
__global__ void __launch_bounds__(512) bn_commit_kernel(const float* __restrict__ pending, const CommitArgs a, const int* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  if (status[2] != 0) return;      // word 2: this call's flag (word 1 is the sticky one sg_d64_train_check reports)
  const int t = blockIdx.x, c = 128 << (t >> 1);
  if (a.dst[t] && threadIdx.x < c) a.dst[t][threadIdx.x] = pending[t * 512 + threadIdx.x];
}

// y = LeakyReLU(raw * scale + shift) -> interior of the zero-bordered [B][S+2][S+2][C] tensor (pad = 1) or plain rows (pad = 0)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ raw, const float* __restrict__ ss, int64_t rows, int C,
                                                       int s_log2, int pad, int split, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int tpr = C >> 3;
  const int64_t total = rows * tpr;
  const int S = 1 << s_log2;
  for (int64_t t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int cg = (int)(t % tpr);
    const int64_t r = t / tpr;
    float x[8], sc[8], sh[8];
    load8(raw + r * C + cg * 8, x);
    load8(ss + cg * 8, sc);
    load8(ss + 512 + cg * 8, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float z = fmaf(x[j], sc[j], sh[j]); x[j] = fmaxf(z, kSlope * z); }
    int64_t orow = r;
    if (pad) {
      const int ow = (int)r & (S - 1), oh = (int)(r >> s_log2) & (S - 1);
      const int64_t b = r >> (2 * s_log2);
      orow = (b * (S + 2) + oh + 1) * (S + 2) + ow + 1;
    }
    store_h8(out + orow * (split ? 2 * C : C) + cg * 8, C, split != 0, x);
  }
}

// dgamma = S2 / scale, dbeta = S1 / scale; coef = gamma*rstd | S1/rows | S2/rows for the apply pass
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int blocks, int64_t rows, int C,
                                                              const float* __restrict__ ss, const float* __restrict__ scal,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float* __restrict__ coef, int* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ch >= C) return;
  double s1, s2;
  warp_partial_sums(partial, blocks, C, ch, lane, s1, s2);
  if (lane != 0) return;
  const float inv = scal[1];
  const float dg = (float)s2 * inv, db = (float)s1 * inv;
  if (!(fabsf(dg) <= 3.0e38f) || !(fabsf(db) <= 3.0e38f)) atomicExch(status + 1, kNonFiniteMagic);
  if (dgamma) dgamma[ch] = dg;
  if (dbeta) dbeta[ch] = db;
  coef[ch] = ss[2048 + ch] * ss[1536 + ch];
  coef[512 + ch] = (float)(s1 / (double)rows);
  coef[1024 + ch] = (float)(s2 / (double)rows);
}

// dconv = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)) -> interior of the zero-bordered dY tensor [B][S+2][S+2][C]
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ raw, const __half* __restrict__ dx,
                                                           const float* __restrict__ ss, const float* __restrict__ coef, int64_t rows,
                                                           int C, int s_log2, int split, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int tpr = C >> 3;
  const int64_t total = rows * tpr;
  const int S = 1 << s_log2;
  for (int64_t t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int cg = (int)(t % tpr);
    const int64_t r = t / tpr;
    float x[8], g[8], sc[8], sh[8], mu[8], rs[8], k0[8], m1[8], m2[8];
    const int Cp = split ? 2 * C : C;
    load8(raw + r * C + cg * 8, x);
    load_h8(dx + r * Cp + cg * 8, C, split != 0, g);
    load8(ss + cg * 8, sc); load8(ss + 512 + cg * 8, sh); load8(ss + 1024 + cg * 8, mu); load8(ss + 1536 + cg * 8, rs);
    load8(coef + cg * 8, k0); load8(coef + 512 + cg * 8, m1); load8(coef + 1024 + cg * 8, m2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(x[j], sc[j], sh[j]);
      const float gj = z > 0.f ? g[j] : kSlope * g[j];
      const float xh = (x[j] - mu[j]) * rs[j];
      x[j] = k0[j] * (gj - m1[j] - xh * m2[j]);
    }
    const int ow = (int)r & (S - 1), oh = (int)(r >> s_log2) & (S - 1);
    const int64_t b = r >> (2 * s_log2);
    const int64_t orow = (b * (S + 2) + oh + 1) * (S + 2) + ow + 1;
    store_h8(out + orow * Cp + cg * 8, C, split != 0, x);
  }
}

// ---- head: conv 512 -> 1 k4 over the 4 x 4 map = an 8192-long dot per image + sigmoid ----------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const __half* __restrict__ act4, int64_t batch, const float* __restrict__ w5p,
                                                       float* __restrict__ logit, float* __restrict__ prob_ws, float* __restrict__ prob,
                                                       int split, int* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  const int Cp = split ? 1024 : 512;
  const __half* a = act4 + b * 16 * Cp;
  float acc = 0.f;
#pragma unroll
  for (int c = threadIdx.x * 8; c < 8192; c += 2048) {
    float x[8], w[8];
    load_h8(a + (c >> 9) * Cp + (c & 511), 512, split != 0, x);
    load8(w5p + c, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(x[j], w[j], acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    acc = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    if (!(fabsf(acc) <= 3.0e38f)) { atomicExch(status + 1, kNonFiniteMagic); atomicExch(status + 2, 1); }
    const float pr = 1.0f / (1.0f + expf(-acc));
    if (logit) logit[b] = acc;
    prob_ws[b] = pr;
    if (prob) prob[b] = pr;
  }
}

// dlogit = dL/dprob * p (1 - p); loss scale = the power of two that brings max |dlogit| * max |w5| to [4, 8)
__global__ void __launch_bounds__(1024) head_bwd_prep_kernel(const float* __restrict__ gout, const float* __restrict__ prob, int64_t batch,
                                                             const float* __restrict__ w5p, float* __restrict__ dlogit,
                                                             float* __restrict__ scal) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[2][32];
  float m1 = 0.f, m2 = 0.f;
  for (int64_t b = threadIdx.x; b < batch; b += 1024) {
    const float p = prob[b];
    const float d = gout[b] * p * (1.f - p);
    dlogit[b] = d;
    m1 = fmaxf(m1, fabsf(d));
  }
  for (int i = threadIdx.x; i < 8192; i += 1024) m2 = fmaxf(m2, fabsf(w5p[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = m1; red[1][threadIdx.x >> 5] = m2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 32; ++i) { m1 = fmaxf(m1, red[0][i]); m2 = fmaxf(m2, red[1][i]); }
    const float m = m1 * m2;
    float s = 1.f;
    if (m > 0.f && m <= 3.0e38f) {
      int e;
      (void)frexpf(m, &e);              // m = f * 2^e, f in [0.5, 1)
      int k = 3 - e;                    // m * 2^k in [4, 8)
      k = k < -60 ? -60 : (k > 60 ? 60 : k);
      s = ldexpf(1.f, k);
    }
    scal[0] = s;
    scal[1] = 1.f / s;
  }
}

// dX4 [B*16][512] = fp16(scale * dlogit[b] * w5[p][c])
__global__ void __launch_bounds__(256) head_bwd_dx_kernel(const float* __restrict__ dlogit, const float* __restrict__ w5p,
                                                          const float* __restrict__ scal, int64_t batch, int split,
                                                          __half* __restrict__ dx4) {
  pdl_launch_dependents();
  pdl_wait();
  const float s = scal[0];
  const int Cp = split ? 1024 : 512;
  const int64_t total = batch * 1024;
  for (int64_t t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int e = (int)(t & 1023);
    const float d = dlogit[t >> 10] * s;
    float w[8];
    load8(w5p + e * 8, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] *= d;
    const int64_t row = (t >> 10) * 16 + (e >> 6);           // (image, pixel); e & 63: the pixel's 8-channel group
    store_h8(dx4 + row * Cp + (e & 63) * 8, 512, split != 0, w);
  }
}

// dw5 [1][512][4][4] = sum_b dlogit[b] * act4[b][p][c]: 64 columns x 4 batch slices per block, fixed-order slice sum
__global__ void __launch_bounds__(256) head_bwd_dw_kernel(const float* __restrict__ dlogit, const __half* __restrict__ act4, int64_t batch,
                                                          int split, float* __restrict__ dw5, int* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[4][64];
  const int col = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int t = blockIdx.x * 64 + col;   // p*512 + c
  float acc = 0.f;
  const int Cp = split ? 1024 : 512;
  const int at = (t >> 9) * Cp + (t & 511);
  for (int64_t b0 = slice; b0 < batch; b0 += 32) {    // eight images per trip, loads first (same order of the adds)
    float av[8], dv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t b = b0 + 4 * u;
      const bool in = b < batch;
      float a = in ? __half2float(act4[b * 16 * Cp + at]) : 0.f;
      if (split && in) a = fmaf(__half2float(act4[b * 16 * Cp + at + 512]), kLoInv, a);
      av[u] = a;
      dv[u] = in ? dlogit[b] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (b0 + 4 * u < batch) acc = fmaf(dv[u], av[u], acc);
  }
  red[slice][col] = acc;
  __syncthreads();
  if (slice == 0) {
    acc = (red[0][col] + red[1][col]) + (red[2][col] + red[3][col]);
    if (!(fabsf(acc) <= 3.0e38f)) atomicExch(status + 1, kNonFiniteMagic);
    dw5[(t & 511) * 16 + (t >> 9)] = acc;
  }
}

// dW [Cout][Cin][4][4] = (1 / scale) * sum_splits partial[s][co][tap * cstride + ci].  One block per (co, 16 input channels):
// thread (tap, ci) sums its column over the splits (64-byte runs), the 16 x 16 tile is transposed through shared memory and
// written as 256 consecutive floats.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int cout, int cin, int cstride,
                                                           int ldn, const float* __restrict__ scal, float* __restrict__ dw,
                                                           int* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[16][17];
  const int groups = (cin + 15) >> 4;
  const int co = blockIdx.x / groups, ci0 = (blockIdx.x - co * groups) * 16;
  const int tap = threadIdx.x >> 4, cil = threadIdx.x & 15;
  float acc = 0.f;
  if (ci0 + cil < cin) {
    const float* src = partial + (size_t)co * ldn + tap * cstride + ci0 + cil;
    const size_t step = (size_t)cout * ldn;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int s = 0;
    for (; s + 16 <= splits; s += 16) {        // sixteen loads in flight per trip (one L2 round trip instead of four); the
      float v[16];                             // association below is that of the 4-wide loop: bit-identical sums
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = src[(size_t)(s + u) * step];
#pragma unroll
      for (int u = 0; u < 16; u += 4) { a0 += v[u]; a1 += v[u + 1]; a2 += v[u + 2]; a3 += v[u + 3]; }
    }
    for (; s + 4 <= splits; s += 4) {          // fixed association: deterministic
      a0 += src[(size_t)s * step]; a1 += src[(size_t)(s + 1) * step]; a2 += src[(size_t)(s + 2) * step]; a3 += src[(size_t)(s + 3) * step];
    }
    for (; s < splits; ++s) a0 += src[(size_t)s * step];
    acc = (a0 + a1) + (a2 + a3);
  }
  acc *= scal[1];
  if (!(fabsf(acc) <= 3.0e38f)) atomicExch(status + 1, kNonFiniteMagic);
  tile[cil][tap] = acc;
  __syncthreads();
  const int ocil = threadIdx.x >> 4, otap = threadIdx.x & 15;
  if (ci0 + ocil < cin) dw[((size_t)co * cin + ci0 + ocil) * 16 + otap] = tile[ocil][otap];
}

// ---- workspace ---------------------------------------------------------------------------------------------------------
struct PackedTrainLayout { size_t wf1, wd1, wf[3], wd[3], w5p, total; };
struct TrainLayout {
  size_t status, scal;
  size_t col1, act1p, raw[3], actp[2], act4n, ss, bnpart, bnpend, prob, dlogit;
  size_t dx[3], dyp[3], dy1, dcol1, partial, coef;
  size_t zero_begin[6], zero_bytes[6];     // the zero-bordered tensors (cleared once by sg_d64_train_workspace_init)
  size_t total;
};
static const int kC[5] = {3, 64, 128, 256, 512};     // channels after layer l
static const int kS[5] = {64, 32, 16, 8, 4};         // spatial size after layer l

// precision 0: fp16 operands; 1: split operands (x = hi + lo in fp16, three tensor passes: the fp32-parity arithmetic)
static PackedTrainLayout packed_train_layout(int precision) {
  PackedTrainLayout L;
  const size_t ns = precision ? 3 : 1;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += align_up(bytes, 1024); return at; };
  L.wf1 = take(64 * 64 * 2 * ns);
  L.wd1 = take(64 * 64 * 2 * ns);
  for (int l = 0; l < 3; ++l) {
    L.wf[l] = take((size_t)kC[l + 2] * 16 * kC[l + 1] * 2 * ns);
    L.wd[l] = take((size_t)kC[l + 2] * 16 * kC[l + 1] * 2 * ns);
  }
  L.w5p = take(8192 * 4);
  L.total = o;
  return L;
}

static TrainLayout train_layout(int64_t cap, int precision) {
  TrainLayout L;
  const size_t b = (size_t)cap * (precision ? 2 : 1);      // 16-bit tensors carry hi | lo halves in split arithmetic
  const size_t b1 = (size_t)cap;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += align_up(bytes, 1024); return at; };
  L.status = take(1024);
  L.scal = take(1024);
  L.col1 = take(b * 1024 * 64 * 2);
  int z = 0;
  L.act1p = take(b * 34 * 34 * 64 * 2);
  L.zero_begin[z] = L.act1p; L.zero_bytes[z++] = b * 34 * 34 * 64 * 2;
  for (int l = 0; l < 3; ++l) L.raw[l] = take(b1 * kS[l + 2] * kS[l + 2] * kC[l + 2] * 4);   // fp32: BatchNorm sees unrounded sums
  for (int l = 0; l < 2; ++l) {
    const size_t bytes = b * (kS[l + 2] + 2) * (kS[l + 2] + 2) * kC[l + 2] * 2;
    L.actp[l] = take(bytes);
    L.zero_begin[z] = L.actp[l]; L.zero_bytes[z++] = bytes;
  }
  L.act4n = take(b * 16 * 512 * 2);
  L.ss = take(3 * 5 * 512 * 4);
  L.bnpart = take((size_t)kBnBlocks * 2 * 512 * 4);
  L.bnpend = take(6 * 512 * 4);
  L.prob = take(b1 * 4);
  L.dlogit = take(b1 * 4);
  for (int l = 0; l < 3; ++l) {
    L.dx[l] = take(b * kS[l + 2] * kS[l + 2] * kC[l + 2] * 2);
    const size_t bytes = b * (kS[l + 2] + 2) * (kS[l + 2] + 2) * kC[l + 2] * 2;
    L.dyp[l] = take(bytes);
    L.zero_begin[z] = L.dyp[l]; L.zero_bytes[z++] = bytes;
  }
  L.dy1 = take(b * 1024 * 64 * 2);
  L.dcol1 = take(b * 1024 * 64 * 2);
  L.partial = take((size_t)256 * 128 * 128 * 4);
  L.coef = take(3 * 3 * 512 * 4);
  L.total = o;
  return L;
}

// zero-bordered NHWC fp16 [B][S+2][S+2][C] as (2C, S/2+1, 2, S/2+1, B); box = 64 channels x ow x 1 x ohb x bb
static int encode_act_map(CUtensorMap* m, const void* ptr, int64_t batch, int S, int C, int ow, int ohb, int bb) {
  const int P = S + 2;
  cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)P / 2, 2, (cuuint64_t)P / 2, (cuuint64_t)batch};
  cuuint64_t strides[4] = {(cuuint64_t)2 * C * 2, (cuuint64_t)P * C * 2, (cuuint64_t)2 * P * C * 2, (cuuint64_t)P * P * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)ow, 1, (cuuint32_t)ohb, (cuuint32_t)bb};
  return encode_tmap(m, 5, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}
// zero-bordered gradient [B][S+2][S+2][C] as (C, S+2, 1, S+2, B)
static int encode_dy_map(CUtensorMap* m, const void* ptr, int64_t batch, int S, int C, int ow, int ohb, int bb) {
  const int P = S + 2;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)P, 1, (cuuint64_t)P, (cuuint64_t)batch};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2, (cuuint64_t)P * C * 2, (cuuint64_t)P * P * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)ow, 1, (cuuint32_t)ohb, (cuuint32_t)bb};
  return encode_tmap(m, 5, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}
static int encode_mat_map(CUtensorMap* m, const void* ptr, int64_t cols, int64_t rows, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  return encode_tmap(m, 2, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// row-box geometry of `rows` pixels (128: M tile, 64: wgrad K tile) of an S x S map
static void row_box(int S, int rows, TrGemm* p, int64_t batch, int* tiles) {
  p->ow = S;
  if (S * S >= rows) { p->ohb = rows / S; p->bb = 1; p->tpi = S / p->ohb; *tiles = (int)batch * p->tpi; }
  else { p->ohb = S; p->bb = rows / (S * S); p->tpi = 1; *tiles = (int)ceil_div(batch, p->bb); }
  p->ow_log2 = ilog2(p->ow);
  p->ohb_log2 = ilog2(p->ohb);
}

#define SG_PDL(...)                              \
  do {                                           \
    const int _pr = sg::launch_pdl(__VA_ARGS__); \
    if (_pr != SG_OK) return _pr;                \
  } while (0)

template <int MODE>
static int launch_trgemm(const CUtensorMap& ta, const CUtensorMap& tb, const TrGemm& p, cudaStream_t st) {
  const int64_t total = (int64_t)p.classes * p.m_tiles * p.n_tiles * p.splits;
  const int grid = (int)(total < state().sm_count ? total : state().sm_count);
  return sg::launch_pdl(trgemm_kernel<MODE>, dim3((unsigned)grid), dim3(192), (size_t)kSmemBytes, st, ta, tb, p);
}

static int ew_blocks(int64_t threads) {
  int64_t b = ceil_div(threads, 256);
  const int64_t cap = (int64_t)state().sm_count * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace dtr
}  // namespace sg

extern "C" {

int sg_d64_train_init_attributes() {
  using namespace sg::dtr;
  SG_CUDA(cudaFuncSetAttribute(trgemm_kernel<MODE_FPROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(trgemm_kernel<MODE_DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(trgemm_kernel<MODE_WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  return SG_OK;
}

size_t sg_d64_train_workspace_bytes(int64_t max_batch, int precision) {
  return sg::dtr::train_layout(max_batch < 1 ? 1 : max_batch, precision).total;
}

int sg_d64_train_workspace_init(void* workspace, int64_t max_batch, int precision, void* stream) {
  using namespace sg::dtr;
  SG_READY();
  SG_REQUIRE(workspace && ((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");
  SG_REQUIRE(max_batch >= 1 && max_batch <= 4096, "max_batch in [1, 4096]");
  SG_REQUIRE(precision == 0 || precision == 1, "precision: 0 (fp16 operands) or 1 (split operands)");
  const TrainLayout L = train_layout(max_batch, precision);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  cudaStream_t st = sg::as_stream(stream);
  SG_CUDA(cudaMemsetAsync(ws + L.status, 0, 2048, st));
  for (int z = 0; z < 6; ++z) SG_CUDA(cudaMemsetAsync(ws + L.zero_begin[z], 0, L.zero_bytes[z], st));
  return SG_OK;
}

size_t sg_d64_train_packed_bytes(int precision) { return sg::dtr::packed_train_layout(precision).total; }

int sg_d64_train_pack(const float* const* h_weights, int precision, void* packed, void* stream) {
  using namespace sg::dtr;
  SG_READY();
  SG_REQUIRE(h_weights && packed && ((uintptr_t)packed & 1023) == 0, "h_weights / 1024-byte aligned packed block");
  SG_REQUIRE(precision == 0 || precision == 1, "precision: 0 (fp16 operands) or 1 (split operands)");
  for (int i = 0; i < 5; ++i) SG_REQUIRE(h_weights[i] != nullptr, "h_weights: conv1..conv5 weight");
  const PackedTrainLayout PL = packed_train_layout(precision);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  PackArgs pa;
  for (int i = 0; i < 5; ++i) pa.w[i] = h_weights[i];
  pa.wf1 = reinterpret_cast<__half*>(pk + PL.wf1);
  pa.wd1 = reinterpret_cast<__half*>(pk + PL.wd1);
  for (int l = 0; l < 3; ++l) { pa.wf[l] = reinterpret_cast<__half*>(pk + PL.wf[l]); pa.wd[l] = reinterpret_cast<__half*>(pk + PL.wd[l]); }
  pa.w5p = reinterpret_cast<float*>(pk + PL.w5p);
  pa.nseg = precision ? 3 : 1;
  SG_PDL(pack_train_kernel, (unsigned)(sg::state().sm_count * 4), 256u, (size_t)0, sg::as_stream(stream), pa);
  return SG_OK;
}

int sg_d64_train_forward(const float* x, int64_t batch, int64_t max_batch, int precision, const void* packed,
                         const float* const* h_bn_params, float* const* h_running_stats, float momentum, float bn_eps,
                         void* workspace, float* prob, float* logit, void* stream) {
  using namespace sg::dtr;
  SG_READY();
  SG_REQUIRE(x && packed && h_bn_params && workspace, "null pointer");
  SG_REQUIRE(((uintptr_t)workspace & 1023) == 0 && ((uintptr_t)packed & 1023) == 0, "workspace / packed must be 1024-byte aligned");
  SG_REQUIRE(max_batch >= 1 && max_batch <= 4096 && batch >= 2 && batch <= max_batch, "2 <= batch <= max_batch <= 4096");
  SG_REQUIRE(precision == 0 || precision == 1, "precision: 0 (fp16 operands) or 1 (split operands)");
  for (int i = 0; i < 6; ++i) SG_REQUIRE(h_bn_params[i] != nullptr, "h_bn_params: gamma2, beta2, gamma3, beta3, gamma4, beta4");
  const TrainLayout L = train_layout(max_batch, precision);
  const PackedTrainLayout PL = packed_train_layout(precision);
  const int split = precision ? 1 : 0, m = split + 1, nseg = split ? 3 : 1;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  cudaStream_t st = sg::as_stream(stream);
  int* status = reinterpret_cast<int*>(ws + L.status);
  auto h16 = [&](size_t off) { return reinterpret_cast<__half*>(ws + off); };
  auto f32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto p16 = [&](size_t off) { return reinterpret_cast<const __half*>(pk + off); };

  SG_PDL(im2col1_kernel, (unsigned)(ew_blocks(batch * 4096)), 256u, (size_t)0, st, x, batch, h16(L.col1), status, split);

  CUtensorMap ta, tb;
  int r;
  {  // layer 1: [B*1024, 64] x [64, 64]^T, LeakyReLU, into the zero-bordered act1
    TrGemm p{};
    p.epi = EPI_L1PAD; p.classes = 1; p.m_tiles = (int)(batch * 8); p.n_tiles = 1; p.k_steps = 1; p.splits = 1; p.kps = 1;
    p.plain = 1; p.tpi = 1; p.bb = 1; p.ohb = 1; p.ow = 1; p.nchunk = 1; p.cin = 64 * m; p.creal = 64; p.nseg = nseg; p.a_lo = 64;
    p.batch = (int)batch; p.m_valid = (int)(batch * 1024); p.n_valid = 64; p.ldo = 64 * m; p.out = h16(L.act1p); p.err = status;
    if ((r = encode_mat_map(&ta, h16(L.col1), 64 * m, batch * 1024, 128)) != SG_OK) return r;
    if ((r = encode_mat_map(&tb, p16(PL.wf1), 64 * nseg, 64, 128)) != SG_OK) return r;
    if ((r = launch_trgemm<MODE_FPROP>(ta, tb, p, st)) != SG_OK) return r;
  }
  for (int l = 0; l < 3; ++l) {   // layers 2..4
    const int cin = kC[l + 1], cout = kC[l + 2], S = kS[l + 2];
    const int64_t rows = batch * S * S;
    const void* in = (l == 0) ? h16(L.act1p) : h16(L.actp[l - 1]);
    TrGemm p{};
    p.epi = EPI_F32; p.classes = 1; p.n_tiles = cout / 128; p.k_steps = 16 * (cin / 64); p.splits = 1; p.kps = p.k_steps;
    p.nchunk = cin / 64; p.cin = cin * m; p.creal = cin; p.nseg = nseg; p.a_lo = cin; p.batch = (int)batch; p.m_valid = (int)rows;
    p.n_valid = cout; p.ldo = cout; p.out = f32(L.raw[l]); p.err = status;
    row_box(S, 128, &p, batch, &p.m_tiles);
    if ((r = encode_act_map(&ta, in, batch, 2 * S, cin * m, p.ow, p.ohb, p.bb)) != SG_OK) return r;
    if ((r = encode_mat_map(&tb, p16(PL.wf[l]), 16 * cin * nseg, cout, 128)) != SG_OK) return r;
    if ((r = launch_trgemm<MODE_FPROP>(ta, tb, p, st)) != SG_OK) return r;

    float* part = reinterpret_cast<float*>(ws + L.bnpart);
    float* ss = reinterpret_cast<float*>(ws + L.ss) + (size_t)l * 5 * 512;
    const int rpp = 256 / (cout / 8);
    int blocks = (int)sg::ceil_div(rows, rpp * 4);
    if (blocks > kBnBlocks) blocks = kBnBlocks;
    SG_PDL(bn_reduce_kernel<false>, (unsigned)(blocks), 256u, (size_t)0, st, f32(L.raw[l]), nullptr, nullptr, rows, cout, 0, part);
    float* rm = h_running_stats ? h_running_stats[2 * l] : nullptr;
    float* rv = h_running_stats ? h_running_stats[2 * l + 1] : nullptr;
    SG_PDL(bn_fwd_finalize_kernel, (unsigned)((cout + 7) / 8), 256u, (size_t)0, st, part, blocks, rows, cout, h_bn_params[2 * l],
           h_bn_params[2 * l + 1], bn_eps, momentum, rm, rv, reinterpret_cast<float*>(ws + L.bnpend) + (size_t)l * 1024, ss);
    __half* out = (l < 2) ? h16(L.actp[l]) : h16(L.act4n);
    SG_PDL(bn_apply_kernel, (unsigned)(ew_blocks(rows * (cout / 8))), 256u, (size_t)0, st, f32(L.raw[l]), ss, rows, cout, ilog2(S),
           l < 2 ? 1 : 0, split, out);
  }
  SG_PDL(head_fwd_kernel, (unsigned)((unsigned)batch), 256u, (size_t)0, st, h16(L.act4n), batch, reinterpret_cast<const float*>(pk + PL.w5p),
         logit, reinterpret_cast<float*>(ws + L.prob), prob, split, status);
  if (h_running_stats) {
    CommitArgs ca;
    for (int i = 0; i < 6; ++i) ca.dst[i] = h_running_stats[i];
    SG_PDL(bn_commit_kernel, 6u, 512u, (size_t)0, st, reinterpret_cast<const float*>(ws + L.bnpend), ca, status);
  }
  return SG_OK;
}

int sg_d64_train_backward(const float* grad_prob, int64_t batch, int64_t max_batch, int precision, const void* packed,
                          void* workspace, float* const* h_grads, float* grad_x, void* stream) {
  using namespace sg::dtr;
  SG_READY();
  SG_REQUIRE(grad_prob && packed && workspace, "null pointer");
  SG_REQUIRE(((uintptr_t)workspace & 1023) == 0 && ((uintptr_t)packed & 1023) == 0, "workspace / packed must be 1024-byte aligned");
  SG_REQUIRE(max_batch >= 1 && max_batch <= 4096 && batch >= 2 && batch <= max_batch, "2 <= batch <= max_batch <= 4096");
  SG_REQUIRE(precision == 0 || precision == 1, "precision: 0 (fp16 operands) or 1 (split operands)");
  if (h_grads) for (int i = 0; i < 11; ++i) SG_REQUIRE(h_grads[i] != nullptr, "h_grads: dw1..dw5, dgamma2, dbeta2, ... (or h_grads = NULL)");
  const TrainLayout L = train_layout(max_batch, precision);
  const int split = precision ? 1 : 0, m = split + 1, nseg = split ? 3 : 1;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  cudaStream_t st = sg::as_stream(stream);
  int* status = reinterpret_cast<int*>(ws + L.status);
  auto h16 = [&](size_t off) { return reinterpret_cast<__half*>(ws + off); };
  float* scal = reinterpret_cast<float*>(ws + L.scal);
  float* dlogit = reinterpret_cast<float*>(ws + L.dlogit);
  const PackedTrainLayout PL = packed_train_layout(precision);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  auto p16 = [&](size_t off) { return reinterpret_cast<const __half*>(pk + off); };
  const float* w5p = reinterpret_cast<const float*>(pk + PL.w5p);
  float* partial = reinterpret_cast<float*>(ws + L.partial);
  const bool want_w = h_grads != nullptr;
  const int sms = sg::state().sm_count;
  CUtensorMap ta, tb;
  int r;

  SG_PDL(head_bwd_prep_kernel, (unsigned)(1), 1024u, (size_t)0, st, grad_prob, reinterpret_cast<const float*>(ws + L.prob), batch, w5p, dlogit, scal);
  SG_PDL(head_bwd_dx_kernel, (unsigned)(ew_blocks(batch * 1024)), 256u, (size_t)0, st, dlogit, w5p, scal, batch, split, h16(L.dx[2]));
  if (want_w) {
    SG_PDL(head_bwd_dw_kernel, (unsigned)(128), 256u, (size_t)0, st, dlogit, h16(L.act4n), batch, split, h_grads[4], status);
  }

  auto splits_for = [&](int tiles, int k_steps, int* kps) {
    int s = sms / tiles;
    if (s < 1) s = 1;
    if (s > k_steps) s = k_steps;
    *kps = (k_steps + s - 1) / s;
    return (k_steps + *kps - 1) / *kps;
  };

  for (int l = 2; l >= 0; --l) {   // layers 4, 3, 2
    const int cin = kC[l + 1], cout = kC[l + 2], S = kS[l + 2];
    const int64_t rows = batch * S * S;
    float* part = reinterpret_cast<float*>(ws + L.bnpart);
    const float* raw = reinterpret_cast<const float*>(ws + L.raw[l]);
    const float* ss = reinterpret_cast<const float*>(ws + L.ss) + (size_t)l * 5 * 512;
    float* coef = reinterpret_cast<float*>(ws + L.coef) + (size_t)l * 3 * 512;
    const int rpp = 256 / (cout / 8);
    int blocks = (int)sg::ceil_div(rows, rpp * 4);
    if (blocks > kBnBlocks) blocks = kBnBlocks;
    SG_PDL(bn_reduce_kernel<true>, (unsigned)(blocks), 256u, (size_t)0, st, raw, h16(L.dx[l]), ss, rows, cout, split, part);
    SG_PDL(bn_bwd_finalize_kernel, (unsigned)((cout + 7) / 8), 256u, (size_t)0, st, part, blocks, rows, cout, ss, scal,
           want_w ? h_grads[5 + 2 * l] : nullptr, want_w ? h_grads[6 + 2 * l] : nullptr, coef, status);
    SG_PDL(bn_bwd_apply_kernel, (unsigned)(ew_blocks(rows * (cout / 8))), 256u, (size_t)0, st, raw, h16(L.dx[l]), ss, coef, rows, cout,
           ilog2(S), split, h16(L.dyp[l]));
    const void* in = (l == 0) ? h16(L.act1p) : h16(L.actp[l - 1]);
    if (want_w) {   // dW_l [cout][16 cin] = dY^T . im2col(act_{l-1}), split-K over the pixels
      TrGemm p{};
      p.epi = EPI_F32; p.classes = 1; p.m_tiles = cout / 128; p.n_tiles = 16 * cin / 128; p.nchunk = 1; p.cin = cin * m; p.creal = cin;
      p.nseg = nseg; p.a_lo = cout; p.b_lo = cin; p.batch = (int)batch;
      p.m_valid = cout; p.n_valid = 16 * cin; p.ldo = 16 * cin; p.out = partial; p.err = status;
      row_box(S, 64, &p, batch, &p.k_steps);
      p.splits = splits_for(p.m_tiles * p.n_tiles, p.k_steps, &p.kps);
      if ((r = encode_dy_map(&ta, h16(L.dyp[l]), batch, S, cout * m, p.ow, p.ohb, p.bb)) != SG_OK) return r;
      if ((r = encode_act_map(&tb, in, batch, 2 * S, cin * m, p.ow, p.ohb, p.bb)) != SG_OK) return r;
      if ((r = launch_trgemm<MODE_WGRAD>(ta, tb, p, st)) != SG_OK) return r;
      SG_PDL(wgrad_reduce_kernel, (unsigned)(cout * (cin / 16)), 256u, (size_t)0, st, partial, p.splits, cout, cin, cin, 16 * cin, scal, h_grads[l + 1], status);
    }
    {   // dX_{l-1}: one GEMM per input-pixel parity class; layer 2's also applies layer 1's LeakyReLU gate -> dY1
      TrGemm p{};
      p.epi = EPI_SCATTER; p.classes = 4; p.n_tiles = (cin + 127) / 128; p.nchunk = cout / 64; p.k_steps = 4 * p.nchunk; p.splits = 1;
      p.kps = p.k_steps; p.cin = cin; p.creal = cin; p.nseg = nseg; p.a_lo = cout; p.batch = (int)batch; p.n_valid = cin; p.ldo = cin * m;
      p.out_s = 2 * S;
      p.out = (l == 0) ? h16(L.dy1) : h16(L.dx[l - 1]);
      p.mask_src = (l == 0) ? h16(L.act1p) : nullptr;
      p.err = status;
      row_box(S, 128, &p, batch, &p.m_tiles);
      if ((r = encode_dy_map(&ta, h16(L.dyp[l]), batch, S, cout * m, p.ow, p.ohb, p.bb)) != SG_OK) return r;
      if ((r = encode_mat_map(&tb, p16(PL.wd[l]), 16 * cout * nseg, cin, 128)) != SG_OK) return r;
      if ((r = launch_trgemm<MODE_DGRAD>(ta, tb, p, st)) != SG_OK) return r;
    }
  }
  if (want_w) {   // dW1 [64][48] = dY1^T . col1
    TrGemm p{};
    p.epi = EPI_F32; p.classes = 1; p.m_tiles = 1; p.n_tiles = 1; p.plain = 1; p.tpi = 1; p.bb = 1; p.ohb = 1; p.ow = 1; p.nchunk = 1;
    p.cin = 64 * m; p.creal = 64; p.nseg = nseg; p.a_lo = 64; p.b_lo = 64;
    p.batch = (int)batch; p.m_valid = 64; p.n_valid = 64; p.ldo = 64; p.out = partial; p.err = status;
    p.k_steps = (int)(batch * 16);
    p.splits = splits_for(1, p.k_steps, &p.kps);
    if ((r = encode_mat_map(&ta, h16(L.dy1), 64 * m, batch * 1024, 64)) != SG_OK) return r;
    if ((r = encode_mat_map(&tb, h16(L.col1), 64 * m, batch * 1024, 64)) != SG_OK) return r;
    if ((r = launch_trgemm<MODE_WGRAD>(ta, tb, p, st)) != SG_OK) return r;
    SG_PDL(wgrad_reduce_kernel, (unsigned)(64), 256u, (size_t)0, st, partial, p.splits, 64, 3, 4, 64, scal, h_grads[0], status);
  }
  if (grad_x) {   // dcol [B*1024][64] = dY1 . W1, then the col2im gather
    TrGemm p{};
    p.epi = EPI_RAW16; p.classes = 1; p.m_tiles = (int)(batch * 8); p.n_tiles = 1; p.k_steps = 1; p.splits = 1; p.kps = 1;
    p.plain = 1; p.tpi = 1; p.bb = 1; p.ohb = 1; p.ow = 1; p.nchunk = 1; p.cin = 64 * m; p.creal = 64; p.nseg = nseg; p.a_lo = 64;
    p.batch = (int)batch; p.m_valid = (int)(batch * 1024); p.n_valid = 64; p.ldo = 64 * m; p.out = h16(L.dcol1); p.err = status;
    if ((r = encode_mat_map(&ta, h16(L.dy1), 64 * m, batch * 1024, 128)) != SG_OK) return r;
    if ((r = encode_mat_map(&tb, p16(PL.wd1), 64 * nseg, 64, 128)) != SG_OK) return r;
    if ((r = launch_trgemm<MODE_DGRAD>(ta, tb, p, st)) != SG_OK) return r;
    SG_PDL(col2im1_kernel, (unsigned)(ew_blocks(batch * 4096)), 256u, (size_t)0, st, h16(L.dcol1), batch, scal, split, grad_x);
  }
  return SG_OK;
}

int sg_d64_train_check(void* workspace, void* stream) {
  SG_READY();
  SG_REQUIRE(workspace != nullptr, "null workspace");
  cudaStream_t st = sg::as_stream(stream);
  int32_t h[2] = {0, 0};
  SG_CUDA(cudaMemcpyAsync(h, workspace, 8, cudaMemcpyDeviceToHost, st));
  SG_CUDA(cudaStreamSynchronize(st));
  if (h[0] != 0 || h[1] != 0) SG_CUDA(cudaMemsetAsync(workspace, 0, 8, st));
  if (h[0] != 0) {
    sg::set_error("the training GEMM pipeline timed out waiting on an mbarrier (role code %d)", h[0]);
    return SG_ECUDA;
  }
  if (h[1] != 0) {
    sg::set_error("a non-finite logit or gradient left the fp16 training path");
    return SG_EINVAL;
  }
  return SG_OK;
}

// debugging / tests: one tensor of the workspace as fp32.  what: 1 act1 [B,64,32,32], 2..4 raw conv output of layer 2..4
// [B,C,S,S], 5..6 normalised activation of layer 2..3, 7 act4 [B,512,4,4]
int sg_d64_train_read(const void* workspace, int64_t batch, int64_t max_batch, int precision, int what, float* out, void* stream);

}  // extern "C"

namespace sg {
namespace dtr {
__global__ void __launch_bounds__(256) read_nhwc_kernel(const __half* __restrict__ src, int64_t batch, int S, int C, int pad, int is_f32,
                                                        int split, float* __restrict__ out) {
  const int64_t total = batch * C * S * S;
  const int P = S + 2 * pad;
  for (int64_t t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int w = (int)(t % S), h = (int)((t / S) % S), c = (int)((t / ((int64_t)S * S)) % C);
    const int64_t b = t / ((int64_t)S * S * C);
    const int64_t at = ((b * P + h + pad) * P + w + pad) * (is_f32 || !split ? C : 2 * C) + c;
    out[t] = is_f32 ? reinterpret_cast<const float*>(src)[at] : __half2float(src[at]) + (split ? __half2float(src[at + C]) * kLoInv : 0.f);
  }
}
}  // namespace dtr
}  // namespace sg

extern "C" int sg_d64_train_read(const void* workspace, int64_t batch, int64_t max_batch, int precision, int what, float* out,
                                 void* stream) {
  using namespace sg::dtr;
  SG_READY();
  SG_REQUIRE(workspace && out && what >= 1 && what <= 7 && batch >= 1 && batch <= max_batch, "workspace / out / what / batch");
  SG_REQUIRE(precision == 0 || precision == 1, "precision");
  const TrainLayout L = train_layout(max_batch, precision);
  const uint8_t* ws = static_cast<const uint8_t*>(workspace);
  size_t off;
  int S, C, pad;
  if (what == 1) { off = L.act1p; S = 32; C = 64; pad = 1; }
  else if (what <= 4) { off = L.raw[what - 2]; S = kS[what]; C = kC[what]; pad = 0; }
  else if (what <= 6) { off = L.actp[what - 5]; S = kS[what - 3]; C = kC[what - 3]; pad = 1; }
  else { off = L.act4n; S = 4; C = 512; pad = 0; }
  read_nhwc_kernel<<<ew_blocks(batch * C * S * S), 256, 0, sg::as_stream(stream)>>>(reinterpret_cast<const __half*>(ws + off), batch, S, C, pad, (what >= 2 && what <= 4) ? 1 : 0, precision, out);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
