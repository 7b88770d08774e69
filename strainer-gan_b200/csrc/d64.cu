// D64 scoring path: Discriminator.forward (eval-mode BN) + sigmoid + BCE vs label 1.
// Replaces "#strainer gan.py:230-256" + ":369-375" (K1..K6 of SURVEY.md §2.3).
//
//   L1  conv 3->64 k4s2p1 + LeakyReLU        conv1_fused_kernel: fp32 NCHW box by TMA, bf16 operand built in the kernel,
//                                            tcgen05 M=128 N=64 K=4x16, TMA-store epilogue
//   L2  conv 64->128  + BN + LeakyReLU       conv2_swap2_kernel: channel-major accumulator (weights = A), plane reuse
//   L3  conv 128->256 + BN + LeakyReLU  \    conv_pair2_kernel: CTA pairs (cta_group::2, 256x256 tiles), plane reuse;
//   L4  conv 256->512 + BN + LeakyReLU  /    fp32 accumulators in TMEM, fused scale/shift/LeakyReLU epilogue
//       (the per-tap-streaming forms of round 1 -- conv_umma_kernel / conv_pair_kernel / conv2_swap_kernel -- are in
//        the git history; DESIGN.md section 4 records what each step bought)
//   L5  conv 512->1 k4s1p0 (8192-dot) + sigmoid + BCE: one warp per sample
//
// Activation layout between layers ("parity planes"): a 4x4/stride-2/pad-1 conv reads input pixel
// (2*oh-1+kh, 2*ow-1+kw).  Storing the SxS activation as [n][hp][wp][S/2][S/2][C] (hp = h&1,
// wp = w&1) makes the im2col slice of one filter tap a dense box of one plane:
// rows (oh+dh, ow+dw), dh = (kh-1)>>1, dw = (kw-1)>>1, plane ((kh-1)&1, (kw-1)&1).  That box is ONE
// 5-D tiled TMA load (out-of-range rows/cols are the zero padding), landing in shared memory as the
// K-major SW128 operand tcgen05.mma wants: 128 output pixels x 64 channels.
//
// fp32-parity mode (SG_CONV_BF16X3): every activation/weight is kept as bf16 hi + bf16 lo
// (hi = rn(x), lo = rn(x - hi)); the GEMM runs three K-segments per (tap, channel chunk):
// A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, i.e. a 3x longer K loop through the same kernel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sg {
namespace d64 {

using namespace ptx;

constexpr int kErrProducer = 1, kErrMma = 2, kErrMmaAcc = 3, kErrEpilogue = 4;
constexpr int kFp16OverflowMagic = 0x46503136;   // workspace word 1 (sticky until sg_d64_check reads it)

// 16-bit operand format of the single-segment modes: bf16 (SG_CONV_BF16) or fp16 (SG_CONV_FP16: 11-bit significand,
// losses within 1e-3 of fp32 in ONE tensor pass; activations must stay below 65504).  Same kernels, same shared
// memory / TMA layouts; only the fp32 <-> 16-bit conversions and the MMA instruction descriptor differ.
template <bool HALF> __device__ __forceinline__ uint16_t act_pack1(float a) {
  return HALF ? __half_as_ushort(__float2half_rn(a)) : __bfloat16_as_ushort(__float2bfloat16_rn(a));
}
template <bool HALF> __device__ __forceinline__ uint32_t act_pack2(float a, float b) {   // a in the low half
  if (HALF) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float act_unpack(uint16_t u, bool half) {
  return half ? __half2float(__ushort_as_half(u)) : __uint_as_float((uint32_t)u << 16);
}
__device__ __forceinline__ uint16_t act_pack_rt(float a, bool half) {
  return half ? __half_as_ushort(__float2half_rn(a)) : __bfloat16_as_ushort(__float2bfloat16_rn(a));
}
constexpr int kBnBlocks = 256;              // fixed row partition of the train-mode BN reduction

// ------------------------------------------------------------------------------------------
// Packed parameter block layout (bytes), shared by sg_d64_pack / sg_d64_score
// ------------------------------------------------------------------------------------------
struct PackedLayout {
  size_t w1, w1t, w2, w3, w4, w5, ss2, ss3, ss4, ident, gb2, gb3, gb4, total;
  int nseg;
};
static PackedLayout packed_layout(int mode) {
  PackedLayout L;
  L.nseg = (mode == SG_CONV_BF16X3) ? 3 : 1;
  size_t o = 0;
  L.w1 = o; o += align_up(48 * 64 * 4, 1024);
  L.w2 = o; o += align_up((size_t)128 * 16 * 64 * L.nseg * 2, 1024);
  L.w3 = o; o += align_up((size_t)256 * 16 * 128 * L.nseg * 2, 1024);
  L.w4 = o; o += align_up((size_t)512 * 16 * 256 * L.nseg * 2, 1024);
  L.w5 = o; o += align_up(16 * 512 * 4, 1024);
  L.w1t = o; o += align_up(64 * 64 * 2 * 2, 1024);  // conv1 weights, bf16 [64][hi|lo][kh*16 + kw*4 + c]
  L.ss2 = o; o += align_up(2 * 128 * 4, 1024);
  L.ss3 = o; o += align_up(2 * 256 * 4, 1024);
  L.ss4 = o; o += align_up(2 * 512 * 4, 1024);
  L.ident = o; o += align_up(2 * 512 * 4, 1024);  // scale 1 | shift 0 (raw conv output for train-mode BN)
  L.gb2 = o; o += align_up(2 * 128 * 4, 1024);    // gamma | beta
  L.gb3 = o; o += align_up(2 * 256 * 4, 1024);
  L.gb4 = o; o += align_up(2 * 512 * 4, 1024);
  L.total = o;
  return L;
}

struct WorkspaceLayout {
  size_t flag, act1, act2, act3, act4, bnpart, bnss, bnrun, total;
  int sega;
};
static WorkspaceLayout workspace_layout(int64_t batch, int mode) {
  WorkspaceLayout L;
  L.sega = (mode == SG_CONV_BF16X3) ? 2 : 1;
  size_t o = 0;
  L.flag = o; o += 1024;
  L.act1 = o; o += align_up((size_t)batch * 32 * 32 * 64 * L.sega * 2, 1024);
  L.act2 = o; o += align_up((size_t)batch * 16 * 16 * 128 * L.sega * 2, 1024);
  L.act3 = o; o += align_up((size_t)batch * 8 * 8 * 256 * L.sega * 2, 1024);
  L.act4 = o; o += align_up((size_t)batch * 16 * 512 * L.sega * 2, 1024);
  L.bnpart = o; o += align_up((size_t)kBnBlocks * 512 * 2 * sizeof(double), 1024);  // train-mode BN partial sums
  L.bnss = o; o += align_up(2 * 512 * 4, 1024);                                      // batch-stat scale | shift
  L.bnrun = o; o += align_up(6 * 512 * 4, 1024);   // fp16 mode: pending running_mean | running_var of BN 2..4 (committed
                                                   // by bn_commit_kernel only if no activation overflowed)
  L.total = o;
  return L;
}

// ------------------------------------------------------------------------------------------
// Weight packing.  w [Cout][Cin][4][4] fp32 -> bf16 [Cout][((tap*nchunk + chunk)*nseg + seg)*64 + j], c = chunk*64 + j;
// seg 0: hi (pairs with A hi), seg 1: hi (pairs with A lo), seg 2: lo (pairs with A hi).
// w1 [64][3][4][4] -> fp32 [k = c*16 + kh*4 + kw][64] and bf16 [co][hi|lo][kh*16 + kw*4 + c];
// w5 [1][512][4][4] -> fp32 [p = kh*4+kw][512];  eval-mode BatchNorm2d folded to y = x * scale + shift.
// ------------------------------------------------------------------------------------------
// All of sg_d64_pack in ONE launch (a training loop repacks after every optimiser step: 7 launches -> 1).
// blockIdx ranges: [0, 1024) conv weights (grid-stride over w2 | w3 | w4), [1024, 1056) small tensors, 1056..1058 BN folds.
struct PackAllArgs {
  const float *w1, *w2, *w3, *w4, *w5;
  const float *g[3], *b[3], *m[3], *v[3];
  __nv_bfloat16 *p2, *p3, *p4, *o1t;
  float *o1, *o5, *ss[3], *gb[3], *ident;
  float eps;
  int nseg;
  int half;   // SG_CONV_FP16: 16-bit weights are fp16 (nseg == 1)
};
__device__ __forceinline__ void pack_conv_elem(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin, int nseg,
                                               int64_t i, bool half) {
  const int nchunk = cin >> 6;
  const int64_t kprime = (int64_t)16 * nchunk * nseg * 64;
  const int co = (int)(i / kprime);
  int64_t r = i - (int64_t)co * kprime;
  const int j = (int)(r & 63); r >>= 6;
  const int seg = (int)(r % nseg); r /= nseg;
  const int chunk = (int)(r % nchunk);
  const int tap = (int)(r / nchunk);
  const float v = w[((int64_t)co * cin + chunk * 64 + j) * 16 + tap];
  if (half) { reinterpret_cast<uint16_t*>(out)[i] = __half_as_ushort(__float2half_rn(v)); return; }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  out[i] = (seg == 2) ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
}
__global__ void __launch_bounds__(256) pack_all_kernel(const PackAllArgs a) {
  const int bid = blockIdx.x;
  if (bid < 1024) {
    const int64_t n2 = (int64_t)128 * 16 * 64 * a.nseg, n3 = (int64_t)256 * 16 * 128 * a.nseg, n4 = (int64_t)512 * 16 * 256 * a.nseg;
    for (int64_t i = bid * 256ll + threadIdx.x; i < n2 + n3 + n4; i += 1024ll * 256) {
      if (i < n2) pack_conv_elem(a.w2, a.p2, 64, a.nseg, i, a.half);
      else if (i < n2 + n3) pack_conv_elem(a.w3, a.p3, 128, a.nseg, i - n2, a.half);
      else pack_conv_elem(a.w4, a.p4, 256, a.nseg, i - n2 - n3, a.half);
    }
  } else if (bid < 1056) {
    const int i = (bid - 1024) * 256 + threadIdx.x;
    if (i < 64 * 128) {  // o1t [co][seg][kh*16 + kw*4 + c], c == 3 is the zero pad channel
      const int co = i >> 7, seg = (i >> 6) & 1, k = i & 63;
      const int kh = k >> 4, kw = (k >> 2) & 3, c = k & 3;
      const float v = (c < 3) ? a.w1[co * 48 + c * 16 + kh * 4 + kw] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      if (a.half) reinterpret_cast<uint16_t*>(a.o1t)[i] = seg ? (uint16_t)0 : __half_as_ushort(__float2half_rn(v));
      else a.o1t[i] = seg ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
    }
    if (i < 48 * 64) a.o1[i] = a.w1[(i & 63) * 48 + (i >> 6)];
    if (i < 16 * 512) a.o5[i] = a.w5[(i & 511) * 16 + (i >> 9)];
  } else {
    const int l = bid - 1056, c = 128 << l;
    for (int i = threadIdx.x; i < c; i += 256) {
      const float sc = a.g[l][i] / sqrtf(a.v[l][i] + a.eps);
      a.ss[l][i] = sc;
      a.ss[l][c + i] = a.b[l][i] - a.m[l][i] * sc;
      a.gb[l][i] = a.g[l][i];
      a.gb[l][c + i] = a.b[l][i];
      if (l == 2) { a.ident[i] = 1.f; a.ident[512 + i] = 0.f; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// L2..L4: implicit-GEMM conv on tcgen05.  Warp-specialised persistent kernel:
//   warp 0   : TMA producer (one elected lane)         smem ring: kStages x (A 128x64 | B BLOCK_N x 64) bf16
//   warp 1   : tcgen05.mma issuer (one elected lane)   accumulators: 2 x BLOCK_N TMEM columns (ping-pong)
//   warps 2-5: epilogue, TMEM -> registers -> scale/shift + LeakyReLU -> bf16 (hi[,lo]) -> global
// ------------------------------------------------------------------------------------------
struct ConvParams {
  int total_tiles, n_tiles;   // tiles = m_tiles * n_tiles, n fastest
  int n_img;                  // images in this launch
  int ow_log2, bh_log2;       // OW = 1<<ow_log2 (= OH), box rows BH = 1<<bh_log2
  int bimg;                   // images per M tile (1, 2, 8)
  int tiles_per_img;          // OH / BH (2 for L2, else 1)
  int c_in;                   // input channels per segment
  int nchunk, nseg, k_steps;  // c_in/64, 1|3, 16*nchunk*nseg
  int c_out;                  // output channels per segment
  int out_sega;               // 1: hi only, 2: hi|lo
  int out_planes;             // 1: parity-plane layout for the next conv, 0: plain [n][oh*OW+ow][C]
  float slope;                // negative-side slope: 0.2 = LeakyReLU, 1.0 = identity (raw conv output)
  const float* scale;         // [c_out]
  const float* shift;         // [c_out]
  __nv_bfloat16* out;
  int* err;
};

// ------------------------------------------------------------------------------------------
// L3 / L4 on CTA pairs with PLANE REUSE.  conv_pair_kernel streams one activation box per filter tap: every input
// pixel is fetched from L2 four times (once per tap that touches it), and the L2 -> SM fabric, not the tensor pipe,
// sets the pace.  Here the K loop is reordered by parity plane: the four taps (2 row shifts x 2
// column shifts) that read plane (ph, pw) share TWO shared-memory copies of that plane, one per column shift
// (loaded with the TMA start column at dw, so the shift costs nothing), each holding W+1 plane rows (one halo
// row, TMA zero fill).  Tile rows are ordered (oh, image, ow): a row shift is then a whole number of 8-row core
// groups, i.e. just a 1024-byte-aligned descriptor offset.  Activation bytes from L2 halve; the per-tap weight
// tiles keep their own (deeper) ring.
//   unit  = (plane, 64-channel chunk, activation segment): 2 copies, 4 taps x {1, 2} weight tiles
// ------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct Pair2Cfg {
  static constexpr int kCopyBytes = 20 * 1024;                 // (H+1) x IMG x W rows of 128 B: 18 KB (L2, L3), 20 KB (L4)
  static constexpr int kUnitBytes = 2 * kCopyBytes;
  static constexpr int kUnits = 3;                             // activation ring
  static constexpr int kBBytes = (BLOCK_N / 2) * 64 * 2;       // this CTA's half of the weight tile
  static constexpr int kBStages = 6;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSsBytes = 2 * 512 * 4;
  static constexpr int kSmemBytes = kUnits * kUnitBytes + kBStages * kBBytes + 256 + kSsBytes + 1024;
  static constexpr int kThreads = 192;
};

template <int SEGA, int BLOCK_N, bool HALF = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv_pair2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const ConvParams p) {
  using Cfg = Pair2Cfg<BLOCK_N>;
  constexpr int UA = Cfg::kUnits, SB = Cfg::kBStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t b_base = base + UA * Cfg::kUnitBytes;
  const uint32_t bar0 = b_base + SB * Cfg::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (bar0 - base));
  auto afull_bar = [&](int s) { return bar0 + 8u * s; };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (UA + s); };
  auto bfull_bar = [&](int s) { return bar0 + 8u * (2 * UA + s); };
  auto bempty_bar = [&](int s) { return bar0 + 8u * (2 * UA + SB + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * UA + 2 * SB + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * UA + 2 * SB + 2 + a); };
  constexpr int kNb = 2 * UA + 2 * SB + 4;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + kNb);
  volatile int* s_abort = reinterpret_cast<volatile int*>(bars + kNb + 1);
  float* s_ss = reinterpret_cast<float*>(smem + (bar0 - base) + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int W = 1 << p.ow_log2;             // output (= plane) width: 16 (L2), 8 (L3), 4 (L4)
  const int H = 1 << p.bh_log2;             // output rows per CTA tile: 8 (L2: the pair splits an image), W otherwise
  const int IMG = p.bimg;                   // images per CTA tile: 1 (L2), 2 (L3), 8 (L4)
  const int oh0 = (p.tiles_per_img > 1) ? (int)rank * H : 0;
  const uint32_t copy_bytes = (uint32_t)((H + 1) * IMG * W * 128);
  const uint32_t shift_bytes = (uint32_t)(IMG * W * 128);   // one plane row of every image = W*IMG/8 core groups
  const int units = 4 * p.nchunk * SEGA;    // (plane, chunk, activation segment)

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    for (int s = 0; s < UA; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(bfull_bar(s), 1); mbar_init(bempty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    *s_abort = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  for (int i = threadIdx.x; i < p.c_out; i += Cfg::kThreads) { s_ss[i] = p.scale[i]; s_ss[512 + i] = p.shift[i]; }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  // unit u -> (plane, chunk, aseg); tap t in 0..3 of a plane -> (kh, kw); weight K-step of (tap, chunk, seg)
  //   plane (ph, pw): kh in {1 - ph + 2*i... }: ph = 1 -> kh = 0 (dh -1), 2 (dh 0); ph = 0 -> kh = 1 (dh 0), 3 (dh +1)
  //   the first of the two has row offset 0 inside the copy, the second one plane row further (see header)
  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int aslot = 0, bstage = 0;
      uint32_t aphase = 0, bphase = 0;
      bool ok = true;
      for (int tile = pair; tile < p.total_tiles && ok; tile += npairs) {
        const int nt = tile % p.n_tiles;
        const int img0 = (p.tiles_per_img > 1) ? tile / p.n_tiles : (2 * (tile / p.n_tiles) + (int)rank) * IMG;
        for (int u = 0; u < units && ok; ++u) {
          const int aseg = u % SEGA;
          const int chunk = (u / SEGA) % p.nchunk;
          const int plane = u / (SEGA * p.nchunk);
          const int ph = plane >> 1, pw = plane & 1;
          if (!mbar_wait(aempty_bar(aslot), aphase ^ 1u, s_abort, p.err, kErrProducer + 30)) { ok = false; break; }
          const uint32_t lead_afull = mapa_shared(afull_bar(aslot), 0);
          if (leader) mbar_arrive_expect_tx(afull_bar(aslot), 4 * copy_bytes);   // 2 copies x 2 CTAs
          const int cc = chunk * 64 + aseg * p.c_in;
          const uint32_t ua = base + aslot * Cfg::kUnitBytes;
          // copy 0: the tap with the smaller kw of this plane; copy 1: the larger.  pw = 1: dw = -1, 0; pw = 0: dw = 0, +1
          tma_load_5d_pair(ua, &tmap_a, lead_afull, cc, pw ? -1 : 0, img0, oh0 + (ph ? -1 : 0), plane);
          tma_load_5d_pair(ua + Cfg::kCopyBytes, &tmap_a, lead_afull, cc, pw ? 0 : 1, img0, oh0 + (ph ? -1 : 0), plane);
          if (++aslot == UA) { aslot = 0; aphase ^= 1u; }
          const int nb = (SEGA == 2 && aseg == 0) ? 2 : 1;   // A_hi pairs with B_hi and B_lo, A_lo with B_hi only
          for (int t = 0; t < 4 && ok; ++t) {
            const int kh = (t >> 1) * 2 + (1 - ph), kw = (t & 1) * 2 + (1 - pw);
            const int tap = kh * 4 + kw;
            for (int b = 0; b < nb; ++b) {
              const int seg = (SEGA == 1) ? 0 : (aseg == 1 ? 1 : (b == 0 ? 0 : 2));
              const int ks = (tap * p.nchunk + chunk) * p.nseg + seg;
              if (!mbar_wait(bempty_bar(bstage), bphase ^ 1u, s_abort, p.err, kErrProducer + 31)) { ok = false; break; }
              const uint32_t lead_bfull = mapa_shared(bfull_bar(bstage), 0);
              if (leader) mbar_arrive_expect_tx(bfull_bar(bstage), 2 * Cfg::kBBytes);
              tma_load_2d_pair(b_base + bstage * Cfg::kBBytes, &tmap_b, lead_bfull, ks * 64, nt * BLOCK_N + (int)rank * (BLOCK_N / 2));
              if (++bstage == SB) { bstage = 0; bphase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    // (the whole warp runs the loop and polls the barriers; elect.sync inside the MMA / commit statements)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_16(256, BLOCK_N, HALF);
      int aslot = 0, bstage = 0, acc = 0;
      uint32_t aphase = 0, bphase = 0, acc_phase = 0;
      bool ok = true;
      // The issuer's instruction stream between two MMAs has to stay well below the MMA's 128 tensor cycles: the four taps of
      // a unit and the four K = 16 steps of a stage are fully unrolled, every operand is a base descriptor plus a constant
      // (the address field holds bytes >> 4 and never carries out of its 14 bits).
      const uint64_t a_desc0 = umma_desc_sw128(base), b_desc0 = umma_desc_sw128(b_base);
      const uint32_t shift16 = shift_bytes >> 4;
      for (int tile = pair; tile < p.total_tiles && ok; tile += npairs) {
        if (!mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u, s_abort, p.err, kErrMmaAcc + 30)) break;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        uint32_t first = 0;
        for (int u = 0; u < units && ok; ++u) {
          const int aseg = u % SEGA;
          if (!mbar_wait(afull_bar(aslot), aphase, s_abort, p.err, kErrMma + 30)) { ok = false; break; }
          tc_fence_after();
          const uint64_t ua = a_desc0 + (uint64_t)((uint32_t)(aslot * Cfg::kUnitBytes) >> 4);
          const int nb = (SEGA == 2 && aseg == 0) ? 2 : 1;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            // t >> 1: second row tap of the plane (one plane row further), t & 1: second column tap (copy 1)
            const uint64_t adesc = ua + (uint64_t)(((t & 1) * Cfg::kCopyBytes) >> 4) + (uint64_t)((t >> 1) * shift16);
            for (int b = 0; b < nb; ++b) {
              if (!mbar_wait(bfull_bar(bstage), bphase, s_abort, p.err, kErrMma + 31)) { ok = false; break; }
              tc_fence_after();
              const uint64_t bdesc = b_desc0 + (uint64_t)((uint32_t)(bstage * Cfg::kBBytes) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16_pair_elect(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, first);
                first = 1u;
              }
              umma_commit_pair_elect(bempty_bar(bstage), 3);
              if (++bstage == SB) { bstage = 0; bphase ^= 1u; }
            }
            if (!ok) break;
          }
          if (!ok) break;
          umma_commit_pair_elect(aempty_bar(aslot), 3);
          if (++aslot == UA) { aslot = 0; aphase ^= 1u; }
        }
        if (!ok) break;
        umma_commit_pair_elect(tfull_bar(acc), 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5 of both CTAs); tile row = (oh, image, ow) =================
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int ct = p.c_out * SEGA;
    const float slope = p.slope;
    const int ow = row & (W - 1);
    const int img_l = (row >> p.ow_log2) % IMG;
    const int oh = oh0 + (row >> p.ow_log2) / IMG;
    for (int tile = pair; tile < p.total_tiles; tile += npairs) {
      const int nt = tile % p.n_tiles;
      const int img = ((p.tiles_per_img > 1) ? tile / p.n_tiles : (2 * (tile / p.n_tiles) + (int)rank) * IMG) + img_l;
      const bool valid = img < p.n_img;
      size_t off;
      if (p.out_planes) {
        const int half = W >> 1;
        off = ((((size_t)img * 4 + ((oh & 1) * 2 + (ow & 1))) * half + (oh >> 1)) * half + (ow >> 1)) * ct;
      } else {
        off = (((size_t)img * W + oh) * W + ow) * ct;
      }
      __nv_bfloat16* dst = p.out + off + (size_t)nt * BLOCK_N;
      const float4* sc4 = reinterpret_cast<const float4*>(s_ss + nt * BLOCK_N);
      const float4* sh4 = reinterpret_cast<const float4*>(s_ss + 512 + nt * BLOCK_N);
      if (!mbar_wait(tfull_bar(acc), acc_phase, s_abort, p.err, kErrEpilogue + 30)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 2
      for (int cb = 0; cb < BLOCK_N; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + cb, v);
        tmem_ld_wait();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = sc4[(cb >> 2) + q], h4 = sh4[(cb >> 2) + q];
          float a0 = fmaf(__uint_as_float(v[4 * q]), s4.x, h4.x), a1 = fmaf(__uint_as_float(v[4 * q + 1]), s4.y, h4.y);
          float a2 = fmaf(__uint_as_float(v[4 * q + 2]), s4.z, h4.z), a3 = fmaf(__uint_as_float(v[4 * q + 3]), s4.w, h4.w);
          a0 = fmaxf(a0, slope * a0); a1 = fmaxf(a1, slope * a1);
          a2 = fmaxf(a2, slope * a2); a3 = fmaxf(a3, slope * a3);
          if (HALF) { hi[2 * q] = act_pack2<true>(a0, a1); hi[2 * q + 1] = act_pack2<true>(a2, a3); continue; }
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(a0, a1), h23 = __floats2bfloat162_rn(a2, a3);
          hi[2 * q] = *reinterpret_cast<const uint32_t*>(&h01);
          hi[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h23);
          if (SEGA == 2) {
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(a0 - f01.x, a1 - f01.y);
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(a2 - f23.x, a3 - f23.y);
            lo[2 * q] = *reinterpret_cast<const uint32_t*>(&l01);
            lo[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&l23);
          }
        }
        if (valid) {
          uint4* d = reinterpret_cast<uint4*>(dst + cb);
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          if (SEGA == 2) {
            uint4* dl = reinterpret_cast<uint4*>(dst + p.c_out + cb);
#pragma unroll
            for (int q = 0; q < 4; ++q) dl[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

struct Conv2RCfg {
  static constexpr int kWBytes = 128 * 64 * 2;            // weights of one tap [128 cout x 64 k]
  static constexpr int kCopyBytes = 17 * 16 * 128;        // one shifted copy of a parity plane: 17 rows x 16 px x 64 ch
  static constexpr int kUnitBytes = kCopyBytes;            // ring slot = ONE copy, shared by the two row taps of its column shift
  static constexpr int kUnits = 3;
  static constexpr int kWStages = 5;
  static constexpr int kTmemCols = 512;
  static constexpr int kStageOut = 128 * 128 * 2;
  static constexpr int kSmemBytes = kUnits * kUnitBytes + kWStages * kWBytes + kStageOut + 256 + 1024;
  static constexpr int kThreads = 192;
};

// conv2_swap_kernel with PLANE REUSE of the pixel operand (see conv_pair2_kernel): the four taps of a parity plane
// read two shared-memory copies of it (one per column shift, 17 plane rows x 16 columns each); a row shift is a
// 2048-byte offset of the B descriptor.  Pixel bytes from L2 halve; the 16 KB weight tiles keep a ring of their own.
template <int SEGA, bool HALF = false>
__global__ void __launch_bounds__(192, 1)
conv2_swap2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_o, const ConvParams p) {
  using Cfg = Conv2RCfg;
  constexpr int S = Cfg::kWStages, UA = Cfg::kUnits;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t w_base = base + UA * Cfg::kUnitBytes;
  const uint32_t stg = w_base + S * Cfg::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (stg - base) + Cfg::kStageOut);
  const uint32_t bar0 = stg + Cfg::kStageOut;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };              // weight ring
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto xfull_bar = [&](int s) { return bar0 + 8u * (2 * S + s); };    // plane-unit ring
  auto xempty_bar = [&](int s) { return bar0 + 8u * (2 * S + UA + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 * UA + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 * UA + 2 + a); };
  constexpr int kNb = 2 * S + 2 * UA + 4;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + kNb);
  volatile int* s_abort = reinterpret_cast<volatile int*>(bars + kNb + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_x);
    prefetch_tensormap(&tmap_w);
    if (SEGA == 1) prefetch_tensormap(&tmap_o);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < UA; ++s) { mbar_init(xfull_bar(s), 1); mbar_init(xempty_bar(s), 1); }
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
      int stage = 0, xslot = 0;
      uint32_t phase = 0, xphase = 0;
      bool ok = true;
      for (int img = blockIdx.x; img < p.n_img && ok; img += gridDim.x) {
        for (int u = 0; u < 8 * SEGA && ok; ++u) {
          const int xseg = u % SEGA, cp = (u / SEGA) & 1, plane = u / (2 * SEGA), ph = plane >> 1, pw = plane & 1;
          if (!mbar_wait(xempty_bar(xslot), xphase ^ 1u, s_abort, p.err, kErrProducer + 40)) { ok = false; break; }
          const uint32_t ux = base + xslot * Cfg::kUnitBytes;
          mbar_arrive_expect_tx(xfull_bar(xslot), Cfg::kCopyBytes);
          // copy cp of plane (ph, pw): column shift dw = cp - pw, first plane row -ph
          tma_load_5d(ux, &tmap_x, xfull_bar(xslot), xseg * 64, cp - pw, ph ? -1 : 0, plane, img);
          if (++xslot == UA) { xslot = 0; xphase ^= 1u; }
          const int nb = (SEGA == 2 && xseg == 0) ? 2 : 1;   // x_hi pairs with w_hi and w_lo, x_lo with w_hi only
          for (int t = 0; t < 2 && ok; ++t) {
            const int kh = t * 2 + (1 - ph), kw = cp * 2 + (1 - pw);
            for (int b = 0; b < nb; ++b) {
              const int seg = (SEGA == 1) ? 0 : (xseg == 1 ? 1 : (b == 0 ? 0 : 2));
              const int ks = (kh * 4 + kw) * p.nseg + seg;
              if (!mbar_wait(empty_bar(stage), phase ^ 1u, s_abort, p.err, kErrProducer + 41)) { ok = false; break; }
              mbar_arrive_expect_tx(full_bar(stage), Cfg::kWBytes);
              tma_load_2d(w_base + stage * Cfg::kWBytes, &tmap_w, full_bar(stage), ks * 64, 0);
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp runs the loop and polls the barriers, elect.sync inside the MMA / commit statements; units
    // and taps are unrolled and every operand is a base descriptor plus an offset, so that the instruction stream between
    // two MMAs stays well below an MMA's 128 tensor cycles (a rolled single-lane loop needed ~100 instructions per stage of
    // four MMAs and left the tensor pipe waiting)
    {
      constexpr uint32_t idesc = umma_idesc_16(128, 256, HALF);
      int stage = 0, xslot = 0, acc = 0;
      uint32_t phase = 0, xphase = 0, acc_phase = 0;
      bool ok = true;
      const uint64_t x_desc0 = umma_desc_sw128(base), w_desc0 = umma_desc_sw128(w_base);
      for (int img = blockIdx.x; img < p.n_img && ok; img += gridDim.x) {
        if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1u, s_abort, p.err, kErrMmaAcc + 40)) break;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
#pragma unroll
        for (int u = 0; u < 8 * SEGA; ++u) {
          const int xseg = u % SEGA;
          if (!mbar_wait(xfull_bar(xslot), xphase, s_abort, p.err, kErrMma + 40)) { ok = false; break; }
          tc_fence_after();
          const uint64_t ux = x_desc0 + (uint64_t)((uint32_t)(xslot * Cfg::kUnitBytes) >> 4);
          const int nb = (SEGA == 2 && xseg == 0) ? 2 : 1;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            // t: second row tap of this copy (one plane row = 16 pixels x 128 B further)
            const uint64_t xdesc = ux + (uint64_t)((t * 2048) >> 4);
#pragma unroll
            for (int b = 0; b < nb; ++b) {
              if (!mbar_wait(full_bar(stage), phase, s_abort, p.err, kErrMma + 41)) { ok = false; break; }
              tc_fence_after();
              const uint64_t wdesc = w_desc0 + (uint64_t)((uint32_t)(stage * Cfg::kWBytes) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_elect(tmem_d, wdesc + 2 * k, xdesc + 2 * k, idesc, (uint32_t)((u | t | b | k) != 0));
              umma_commit_elect(empty_bar(stage));
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            if (!ok) break;
          }
          if (!ok) break;
          umma_commit_elect(xempty_bar(xslot));
          if (++xslot == UA) { xslot = 0; xphase ^= 1u; }
        }
        if (!ok) break;
        umma_commit_elect(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int lg = warp & 3;
    const int ch = lg * 32 + lane;          // this thread's output channel
    const float sc = __ldg(p.scale + ch), sh = __ldg(p.shift + ch);
    const int ct = 128 * p.out_sega;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int img = blockIdx.x; img < p.n_img; img += gridDim.x) {
      if (!mbar_wait(tfull_bar(acc), acc_phase, s_abort, p.err, kErrEpilogue + 20)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 256);
      __nv_bfloat16* out_img = p.out + (size_t)img * 4 * 64 * ct + ch;
      if (SEGA == 1) {
        const bool issuer = (threadIdx.x == 64);
        const float slope = p.slope;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          if (issuer) tma_store_wait_read<0>();       // the previous store has finished reading the staging tile
          named_bar_sync(1, 128);
#pragma unroll 1
          for (int pq = 0; pq < 4; ++pq) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + half * 128 + pq * 32, v);
            tmem_ld_wait();
            // pixel (oh_l = 2*pq + (j >> 4), ow = j & 15) of this half -> staging row (plane, oh_l >> 1, ow >> 1)
            const uint32_t rbase = stg + (uint32_t)(pq * 8 * 256 + ch * 2);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float a = fmaf(__uint_as_float(v[j]), sc, sh);
              a = fmaxf(a, slope * a);
              const int row = (((j >> 4) & 1) * 2 + (j & 1)) * 32 + ((j & 15) >> 1);
              st_shared_u16(rbase + (uint32_t)(row * 256), act_pack1<HALF>(a));
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (issuer) {
            tma_store_5d(&tmap_o, stg, 0, 0, half * 4, 0, img);
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int pb = 0; pb < 256; pb += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + pb, v);
          tmem_ld_wait();
  #pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int px = pb + j, oh = px >> 4, ow = px & 15;
            float a = fmaf(__uint_as_float(v[j]), sc, sh);
            a = a > 0.f ? a : p.slope * a;
            const __nv_bfloat16 ah = __float2bfloat16_rn(a);
            __nv_bfloat16* d = out_img + ((size_t)((oh & 1) * 2 + (ow & 1)) * 64 + (oh >> 1) * 8 + (ow >> 1)) * ct;
            *d = ah;
            if (p.out_sega == 2) d[128] = __float2bfloat16_rn(a - __bfloat162float(ah));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (SEGA == 1 && threadIdx.x == 64) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// L1 fused with the input conversion: the fp32 NCHW image is the only HBM read (49 152 B/sample), no
// bf16 staging copy.  One tile = 4 output rows x 32 cols of one image:
//   warp 0    : TMA producer, fp32 box [3 ch][10 input rows][64 cols] (rows -1 / 64 zero-filled by TMA)
//   warps 6-9 : converters, one output pixel per thread: 4 x (LDS.64 + 2 shuffles) per channel give the
//               4x4x3 patch, packed to the K-major SWIZZLE_32B bf16 operand (K = kh | kw, c4) of the 4 kh slices
//   warp 1    : tcgen05.mma issuer (M = 128, N = 64, 4 x K = 16; 12 MMAs in fp32-parity mode)
//   warps 2-5 : epilogue, TMEM -> LeakyReLU -> bf16 -> swizzled staging -> TMA store into act1 planes
// ------------------------------------------------------------------------------------------
template <int SEGA>
struct Conv1FCfg {
  static constexpr int kRawBytes = 3 * 10 * 64 * 4;   // one fp32 input box
  static constexpr int kRawStride = 8192;
  static constexpr int kRawStages = 3;
  static constexpr int kSliceA = 128 * 32;
  static constexpr int kSliceB = 64 * 32;
  static constexpr int kAStage = SEGA * 4 * kSliceA;
  static constexpr int kAStages = 2;
  static constexpr int kBBytes = SEGA * 4 * kSliceB;
  static constexpr int kTmemCols = 128;
  static constexpr int kStageOut = SEGA * 128 * 128;
  static constexpr int kSmemBytes = kAStages * kAStage + kBBytes + 2 * kStageOut + kRawStages * kRawStride + 256 + 1024;
  static constexpr int kThreads = 320;
};

template <int SEGA, bool HALF = false>
__global__ void __launch_bounds__(320, SEGA == 1 ? 2 : 1)
conv1_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_o, int total_tiles, int* err) {
  using Cfg = Conv1FCfg<SEGA>;
  constexpr int SA = Cfg::kAStages, SR = Cfg::kRawStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t b_base = base + SA * Cfg::kAStage;
  const uint32_t o_base = b_base + Cfg::kBBytes;
  const uint32_t r_base = o_base + 2 * Cfg::kStageOut;
  const uint32_t bar0 = r_base + SR * Cfg::kRawStride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (bar0 - base));
  auto rfull_bar = [&](int s) { return bar0 + 8u * s; };
  auto rempty_bar = [&](int s) { return bar0 + 8u * (SR + s); };
  auto afull_bar = [&](int s) { return bar0 + 8u * (2 * SR + s); };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (2 * SR + SA + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * SR + 2 * SA + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * SR + 2 * SA + 2 + a); };
  constexpr int kNb = 2 * SR + 2 * SA + 4;
  const uint32_t wbar = bar0 + 8u * kNb;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + kNb + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(bars + kNb + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_x);
    prefetch_tensormap(&tmap_b);
    prefetch_tensormap(&tmap_o);
    for (int s = 0; s < SR; ++s) { mbar_init(rfull_bar(s), 1); mbar_init(rempty_bar(s), 128); }
    for (int s = 0; s < SA; ++s) { mbar_init(afull_bar(s), 128); mbar_init(aempty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    mbar_init(wbar, 1);
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
      mbar_arrive_expect_tx(wbar, Cfg::kBBytes);
      for (int sg_ = 0; sg_ < SEGA; ++sg_)
        for (int kh = 0; kh < 4; ++kh)
          tma_load_2d(b_base + (sg_ * 4 + kh) * Cfg::kSliceB, &tmap_b, wbar, sg_ * 64 + kh * 16, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile >> 3, oh0 = (tile & 7) << 2;
        if (!mbar_wait(rempty_bar(stage), phase ^ 1u, s_abort, err, kErrProducer + 10)) break;
        mbar_arrive_expect_tx(rfull_bar(stage), Cfg::kRawBytes);
        tma_load_4d(r_base + stage * Cfg::kRawStride, &tmap_x, rfull_bar(stage), 0, 2 * oh0 - 1, 0, n);
        if (++stage == SR) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_16(128, 64, HALF);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      bool ok = mbar_wait(wbar, 0, s_abort, err, kErrMma + 10);
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1u, s_abort, err, kErrMmaAcc + 10)) break;
        if (!mbar_wait(afull_bar(stage), phase, s_abort, err, kErrMma + 10)) break;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 64);
        const uint32_t sa = base + stage * Cfg::kAStage;
#pragma unroll
        for (int kh = 0; kh < 4; ++kh) {
          const uint64_t a_hi = umma_desc_sw32(sa + kh * Cfg::kSliceA);
          const uint64_t b_hi = umma_desc_sw32(b_base + kh * Cfg::kSliceB);
          umma_f16(tmem_d, a_hi, b_hi, idesc, (uint32_t)(kh != 0));
          if (SEGA == 2) {
            const uint64_t a_lo = umma_desc_sw32(sa + (4 + kh) * Cfg::kSliceA);
            const uint64_t b_lo = umma_desc_sw32(b_base + (4 + kh) * Cfg::kSliceB);
            umma_f16(tmem_d, a_lo, b_hi, idesc, 1u);
            umma_f16(tmem_d, a_hi, b_lo, idesc, 1u);
          }
        }
        umma_commit(aempty_bar(stage));
        umma_commit(tfull_bar(acc));
        if (++stage == SA) { stage = 0; phase ^= 1u; }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 6) {
    // ================= converters: fp32 patch -> bf16 K-major SW32 operand rows =================
    // thread = (output row ohl, pixel PAIR 2p / 2p+1, filter-row half): one 128-bit load per channel covers the
    // input columns 4p .. 4p+3, the two outer columns come from the neighbouring pairs by shuffle -- 41 % fewer
    // shared-memory / shuffle instructions than one pixel per thread (the kernel was MIO bound, profiles/r1d).
    const int ohl = warp - 6, pr = lane & 15, khh = lane >> 4;
    const int m0 = ohl * 32 + 2 * pr;                         // operand rows m0, m0 + 1 (same swizzle phase)
    const uint32_t row_off = (uint32_t)m0 * 32u;
    const uint32_t c0off = (uint32_t)(((m0 >> 2) & 1) << 4);   // SWIZZLE_32B: 16-B chunk ^= address bit 7
    int rs = 0, as = 0;
    uint32_t rphase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (!mbar_wait(rfull_bar(rs), rphase, s_abort, err, kErrProducer + 30)) break;
      if (!mbar_wait(aempty_bar(as), aphase ^ 1u, s_abort, err, kErrProducer + 31)) break;
      const uint32_t raw = r_base + rs * Cfg::kRawStride + (uint32_t)(pr * 16);
      const uint32_t dst = base + as * Cfg::kAStage + row_off;
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int kh = 2 * khh + k2;
        float px[2][4][3];   // [pixel][kw][c]
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float4 v = ld_shared_f4(raw + (uint32_t)(((c * 10 + 2 * ohl + kh) * 64) * 4));
          float l = __shfl_up_sync(0xffffffffu, v.w, 1, 16);
          float r = __shfl_down_sync(0xffffffffu, v.x, 1, 16);
          if (pr == 0) l = 0.f;      // input column -1
          if (pr == 15) r = 0.f;     // input column 64
          px[0][0][c] = l;   px[0][1][c] = v.x; px[0][2][c] = v.y; px[0][3][c] = v.z;
          px[1][0][c] = v.y; px[1][1][c] = v.z; px[1][2][c] = v.w; px[1][3][c] = r;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int kw = 0; kw < 4; ++kw) {
            if (HALF || SEGA == 1) {   // packed two-value conversions (the kernel is issue bound: 0.315 -> 0.27 ms)
              hi[2 * kw] = act_pack2<HALF>(px[q][kw][0], px[q][kw][1]);
              hi[2 * kw + 1] = (uint32_t)act_pack1<HALF>(px[q][kw][2]);
              continue;
            }
            const __nv_bfloat16 h0 = __float2bfloat16_rn(px[q][kw][0]), h1 = __float2bfloat16_rn(px[q][kw][1]),
                                h2 = __float2bfloat16_rn(px[q][kw][2]);
            hi[2 * kw] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            hi[2 * kw + 1] = (uint32_t)__bfloat16_as_ushort(h2);
            if (SEGA == 2) {
              const __nv_bfloat16 l0 = __float2bfloat16_rn(px[q][kw][0] - __bfloat162float(h0));
              const __nv_bfloat16 l1 = __float2bfloat16_rn(px[q][kw][1] - __bfloat162float(h1));
              const __nv_bfloat16 l2 = __float2bfloat16_rn(px[q][kw][2] - __bfloat162float(h2));
              lo[2 * kw] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              lo[2 * kw + 1] = (uint32_t)__bfloat16_as_ushort(l2);
            }
          }
          const uint32_t d = dst + (uint32_t)(kh * Cfg::kSliceA + q * 32);
          st_shared_v4(d + c0off, hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(d + (c0off ^ 16u), hi[4], hi[5], hi[6], hi[7]);
          if (SEGA == 2) {
            st_shared_v4(d + 4 * Cfg::kSliceA + c0off, lo[0], lo[1], lo[2], lo[3]);
            st_shared_v4(d + 4 * Cfg::kSliceA + (c0off ^ 16u), lo[4], lo[5], lo[6], lo[7]);
          }
        }
      }
      mbar_arrive(rempty_bar(rs));       // every LDS of this box has been consumed by the stores above
      fence_proxy_async_smem();          // generic-proxy operand writes -> visible to tcgen05.mma
      mbar_arrive(afull_bar(as));
      if (++rs == SR) { rs = 0; rphase ^= 1u; }
      if (++as == SA) { as = 0; aphase ^= 1u; }
    }
  } else {
    // ================= epilogue (warps 2..5): as conv1_umma_kernel =================
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int ohl = row >> 5, ow = row & 31;
    const int prow = (((ohl & 1) * 2 + (ow & 1)) << 5) + ((ohl >> 1) << 4) + (ow >> 1);
    const uint32_t swz = (uint32_t)(prow & 7);
    const bool issuer = (threadIdx.x == 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n = tile >> 3, oh0 = (tile & 7) << 2;
      const uint32_t stg = o_base + (uint32_t)(it & 1) * Cfg::kStageOut;
      if (!mbar_wait(tfull_bar(acc), acc_phase, s_abort, err, kErrEpilogue + 10)) break;
      tc_fence_after();
      if (issuer) tma_store_wait_read<1>();
      named_bar_sync(1, 128);
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 64);
#pragma unroll
      for (int cb = 0; cb < 64; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + cb, v);
        tmem_ld_wait();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
          a = a > 0.f ? a : 0.2f * a;
          b = b > 0.f ? b : 0.2f * b;
          if (HALF || SEGA == 1) { hi[j] = act_pack2<HALF>(a, b); continue; }
          const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh2 = __float2bfloat16_rn(b);
          hi[j] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh2) << 16);
          if (SEGA == 2) {
            const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
            const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh2));
            lo[j] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
          }
        }
        const uint32_t rbase = stg + (uint32_t)prow * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t chunk = (uint32_t)((cb >> 3) + q) ^ swz;
          st_shared_v4(rbase + chunk * 16u, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          if (SEGA == 2)
            st_shared_v4(rbase + 128u * 128u + chunk * 16u, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer) {
#pragma unroll
        for (int sg_ = 0; sg_ < SEGA; ++sg_)
#pragma unroll
          for (int pl = 0; pl < 4; ++pl)
            tma_store_5d(&tmap_o, stg + (uint32_t)(sg_ * 128 * 128 + pl * 4096), sg_ * 64, 0, oh0 >> 1, pl, n);
        tma_store_commit();
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// Train-mode BatchNorm (in-batch strain, "# 상위 10% 제거해서 fake image에 concate.py:244-245": the
// reference scores with netD in TRAIN mode under no_grad, so BN normalises with batch statistics
// and updates running_mean / running_var).  The conv kernels write the raw conv
// output (identity epilogue); these three kernels then reduce per-channel moments over all
// rows in a fixed order (fp64), fold them into scale/shift (+ running-stat update, momentum,
// unbiased variance) and apply scale/shift + LeakyReLU in place.  All layouts are channel-innermost,
// so the kernels are layout agnostic: rows x (C*sega).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ act, int64_t rows, int c,
                                                       int sega, int half, double* __restrict__ part) {
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * per, r1 = min(r0 + per, rows);
  const int ct = c * sega;
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    double s = 0.0, q = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
      float v = act_unpack(reinterpret_cast<const uint16_t*>(act)[r * ct + ch], half != 0);
      if (sega == 2) v += __bfloat162float(act[r * ct + c + ch]);
      s += (double)v;
      q = fma((double)v, (double)v, q);
    }
    part[((size_t)blockIdx.x * c + ch) * 2] = s;
    part[((size_t)blockIdx.x * c + ch) * 2 + 1] = q;
  }
}

// One warp per channel: lane l adds the partials of blocks l, l+32, ... and the 32 lane sums are combined by a
// fixed butterfly (deterministic; a single thread walking 256 strided doubles took 55 us per layer).
__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ part, int blocks, int64_t rows, int c,
                                                          const float* __restrict__ gb, float eps, float momentum,
                                                          const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                                          float* __restrict__ new_mean, float* __restrict__ new_var,
                                                          float* __restrict__ ss) {
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  double s = 0.0, q = 0.0;
  for (int b = lane; b < blocks; b += 32) {
    s += part[((size_t)b * c + ch) * 2];
    q += part[((size_t)b * c + ch) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane != 0) return;
  const double mean = s / (double)rows;
  double var = q / (double)rows - mean * mean;   // biased (normalisation)
  if (var < 0.0) var = 0.0;
  const float scale = gb[ch] / sqrtf((float)var + eps);
  ss[ch] = scale;
  ss[512 + ch] = gb[c + ch] - (float)mean * scale;
  if (running_mean) new_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    new_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

// fp16 mode: the running statistics of the three BatchNorm layers become visible only if the whole batch stayed finite
// (status[1] == 0 after the head): an overflowing batch is scored again in another conv mode by the caller and must
// update the statistics exactly once.
struct BnCommitArgs { float* dst[6]; };
__global__ void __launch_bounds__(512) bn_commit_kernel(const float* __restrict__ pending, const BnCommitArgs a,
                                                        const int* __restrict__ status) {
  if (status[1] != 0) return;
  const int t = blockIdx.x, c = 128 << (t >> 1);
  if (a.dst[t] && threadIdx.x < c) a.dst[t][threadIdx.x] = pending[t * 512 + threadIdx.x];
}

__global__ void __launch_bounds__(256) bn_apply_kernel(__nv_bfloat16* __restrict__ act, int64_t rows, int c, int sega,
                                                       int half, const float* __restrict__ ss) {
  const int groups = c >> 3;                                  // 8 channels (16 B) per thread
  const int64_t total = rows * groups;
  const int ct = c * sega;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / groups;
    const int ch = (int)(i - r * groups) << 3;
    uint4* ph = reinterpret_cast<uint4*>(act + r * ct + ch);
    uint4* pl = reinterpret_cast<uint4*>(act + r * ct + c + ch);
    const uint4 h = *ph;
    uint4 l = make_uint4(0, 0, 0, 0);
    if (sega == 2) l = *pl;
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a, b;
      if (half) {
        a = act_unpack((uint16_t)(hw[k] & 0xFFFFu), true);
        b = act_unpack((uint16_t)(hw[k] >> 16), true);
      } else {
        a = __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
        b = __uint_as_float(hw[k] & 0xFFFF0000u) + __uint_as_float(lw[k] & 0xFFFF0000u);
      }
      a = fmaf(a, ss[ch + 2 * k], ss[512 + ch + 2 * k]);
      b = fmaf(b, ss[ch + 2 * k + 1], ss[512 + ch + 2 * k + 1]);
      a = a > 0.f ? a : 0.2f * a;
      b = b > 0.f ? b : 0.2f * b;
      if (half) { oh[k] = act_pack2<true>(a, b); ol[k] = 0u; continue; }
      const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
      oh[k] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
      const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
      const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh));
      ol[k] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
    }
    *ph = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    if (sega == 2) *pl = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

// ------------------------------------------------------------------------------------------
// L5 head: logit = <act4[n], w5>, prob = sigmoid(logit), loss = -max(log prob, -100).
// One CTA per sample, thread t owns 8 channels of 4 pixels; fixed summation order (thread chains, xor-shuffle tree,
// fixed-order sum of the 8 warp partials).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_kernel(const __nv_bfloat16* __restrict__ act4, const float* __restrict__ w5p,
                                                         int sega, int half, float* __restrict__ logit, float* __restrict__ prob,
                                                         float* __restrict__ loss, int* __restrict__ err) {
  __shared__ float red[8];
  const int64_t n = blockIdx.x;
  const int ct = 512 * sega;
  const __nv_bfloat16* a = act4 + (size_t)n * 16 * ct;
  const int c = (threadIdx.x & 63) * 8, p0 = threadIdx.x >> 6;
  float accv = 0.f;
#pragma unroll
  for (int pi = 0; pi < 4; ++pi) {
    const int pxl = p0 * 4 + pi;
    const __nv_bfloat16* ap = a + (size_t)pxl * ct;
    const float* wp = w5p + pxl * 512;
    const uint4 raw = *reinterpret_cast<const uint4*>(ap + c);
    const float4 w0 = *reinterpret_cast<const float4*>(wp + c);
    const float4 w1 = *reinterpret_cast<const float4*>(wp + c + 4);
    const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
    float xv[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (half) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[q]));
        xv[2 * q] = f.x;
        xv[2 * q + 1] = f.y;
      } else {
        xv[2 * q] = __uint_as_float(rw[q] << 16);
        xv[2 * q + 1] = __uint_as_float(rw[q] & 0xFFFF0000u);
      }
    }
    if (sega == 2) {
      const uint4 rl = *reinterpret_cast<const uint4*>(ap + 512 + c);
      const uint32_t rlw[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        xv[2 * q] += __uint_as_float(rlw[q] << 16);
        xv[2 * q + 1] += __uint_as_float(rlw[q] & 0xFFFF0000u);
      }
    }
    accv = fmaf(xv[0], w0.x, accv); accv = fmaf(xv[1], w0.y, accv);
    accv = fmaf(xv[2], w0.z, accv); accv = fmaf(xv[3], w0.w, accv);
    accv = fmaf(xv[4], w1.x, accv); accv = fmaf(xv[5], w1.y, accv);
    accv = fmaf(xv[6], w1.z, accv); accv = fmaf(xv[7], w1.w, accv);
  }
  accv = warp_sum(accv);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = accv;
  __syncthreads();
  if (threadIdx.x == 0) {
    accv = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    if (half && !(fabsf(accv) <= 3.4e38f)) atomicExch(err + 1, kFp16OverflowMagic);
    const float pr = 1.0f / (1.0f + expf(-accv));
    if (logit) logit[n] = accv;
    if (prob) prob[n] = pr;
    if (loss) loss[n] = -fmaxf(logf(pr), -100.0f);
  }
}

// debug/test: parity-plane (or plain) bf16 activation -> fp32 NCHW (hi + lo)
__global__ void read_activation_kernel(const __nv_bfloat16* __restrict__ act, int64_t batch, int s, int c, int sega,
                                       int half_fmt, int planes, float* __restrict__ out) {
  const int64_t total = batch * c * s * s;
  const int ct = c * sega;
  const int half = s >> 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int w = (int)(r % s); r /= s;
    const int h = (int)(r % s); r /= s;
    const int ch = (int)(r % c);
    const int64_t n = r / c;
    size_t off;
    if (planes) off = ((((size_t)n * 4 + ((h & 1) * 2 + (w & 1))) * half + (h >> 1)) * half + (w >> 1)) * ct;
    else off = (((size_t)n * s + h) * s + w) * ct;
    float v = act_unpack(reinterpret_cast<const uint16_t*>(act)[off + ch], half_fmt != 0);
    if (sega == 2) v += __bfloat162float(act[off + c + ch]);
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int encode(CUtensorMap* m, int rank, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
                  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  return encode_tmap(m, rank, ptr, dims, strides, box, swz, dtype);
}

// plane-reuse pair kernel for L2 (BLOCK_N = 128: the pair splits one image by rows), L3, L4 (BLOCK_N = 256)
template <int BLOCK_N>
static int launch_conv_pair2(const __nv_bfloat16* act_in, const __nv_bfloat16* wpk, const float* scale, const float* shift,
                             __nv_bfloat16* act_out, int64_t batch, int s_in, int c_in, int c_out, int nseg, int sega,
                             int out_planes, float slope, int* err, cudaStream_t stream, bool half = false) {
  using Cfg = Pair2Cfg<BLOCK_N>;
  const int ow = s_in / 2;                 // output width = parity-plane width
  const bool split = ow * ow > 128;        // L2: 256 output pixels per image -> one image per CTA pair
  const int h_tile = split ? 128 / ow : ow;
  const int bimg = split ? 1 : 128 / (ow * ow);
  const int ct_in = c_in * sega;
  CUtensorMap ta, tb;
  {
    // parity planes [img][plane][h][w][c] addressed as (c, w, img, h, plane): a copy is w fastest, then image, then
    // plane row, so that a row shift of the window is a whole number of 8-row core groups
    cuuint64_t dims[5] = {(cuuint64_t)ct_in, (cuuint64_t)ow, (cuuint64_t)batch, (cuuint64_t)ow, 4};
    cuuint64_t strides[4] = {(cuuint64_t)ct_in * 2, (cuuint64_t)4 * ow * ow * ct_in * 2, (cuuint64_t)ow * ct_in * 2,
                             (cuuint64_t)ow * ow * ct_in * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)ow, (cuuint32_t)bimg, (cuuint32_t)(h_tile + 1), 1};
    int r = encode(&ta, 5, act_in, dims, strides, box);
    if (r != SG_OK) return r;
  }
  const int nchunk = c_in / 64;
  const int k_steps = 16 * nchunk * nseg;
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_steps * 64, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)k_steps * 64 * 2};
    cuuint32_t box[2] = {64, BLOCK_N / 2};
    int r = encode(&tb, 2, wpk, dims, strides, box);
    if (r != SG_OK) return r;
  }
  ConvParams p = {};
  p.n_tiles = c_out / BLOCK_N;
  const int64_t pair_m_tiles = split ? batch : ceil_div(ceil_div(batch, bimg), 2);
  p.total_tiles = (int)(pair_m_tiles * p.n_tiles);
  p.n_img = (int)batch;
  p.ow_log2 = __builtin_ctz(ow);
  p.bh_log2 = __builtin_ctz(h_tile);
  p.bimg = bimg;
  p.tiles_per_img = split ? 2 : 1;
  p.c_in = c_in;
  p.nchunk = nchunk;
  p.nseg = nseg;
  p.k_steps = k_steps;
  p.c_out = c_out;
  p.out_sega = sega;
  p.out_planes = out_planes;
  p.slope = slope;
  p.scale = scale;
  p.shift = shift;
  p.out = act_out;
  p.err = err;
  int pairs = state().sm_count / 2;
  if (p.total_tiles < pairs) pairs = p.total_tiles;
  if (sega == 2) conv_pair2_kernel<2, BLOCK_N><<<2 * pairs, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  else if (half) conv_pair2_kernel<1, BLOCK_N, true><<<2 * pairs, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  else conv_pair2_kernel<1, BLOCK_N><<<2 * pairs, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

static int launch_conv2_swap(const __nv_bfloat16* act1, const __nv_bfloat16* wpk, const float* scale, const float* shift,
                             __nv_bfloat16* act2,
                             int64_t batch, int nseg, int sega, float slope, int* err, cudaStream_t stream, bool half = false) {
  CUtensorMap tw;
  const int ct_in = 64 * sega;
  const int k_steps = 16 * nseg;
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_steps * 64, 128};
    cuuint64_t strides[1] = {(cuuint64_t)k_steps * 64 * 2};
    cuuint32_t box[2] = {64, 128};
    int r = encode(&tw, 2, wpk, dims, strides, box);
    if (r != SG_OK) return r;
  }
  ConvParams p = {};
  p.n_img = (int)batch;
  p.nseg = nseg;
  p.k_steps = k_steps;
  p.c_in = 64;
  p.c_out = 128;
  p.out_sega = sega;
  p.slope = slope;
  p.scale = scale;
  p.shift = shift;
  p.out = act2;
  p.err = err;
  const int grid = (int)(batch < state().sm_count ? batch : state().sm_count);
  CUtensorMap to;
  {
    // act2 parity planes [img][4][8][8][128*sega]; one store = 4 plane rows of all 4 planes of one image (128 px x 128 ch)
    const cuuint64_t ct = 128 * sega;
    cuuint64_t dims[5] = {ct, 8, 8, 4, (cuuint64_t)batch};
    cuuint64_t strides[4] = {ct * 2, 8 * ct * 2, 64 * ct * 2, 256 * ct * 2};
    cuuint32_t box[5] = {128, 8, 4, 4, 1};
    int r = encode(&to, 5, act2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r != SG_OK) return r;
  }
  CUtensorMap txr;   // one shifted copy of a parity plane: 17 plane rows x 16 columns x 64 channels
  {
    cuuint64_t dims[5] = {(cuuint64_t)ct_in, 16, 16, 4, (cuuint64_t)batch};
    cuuint64_t strides[4] = {(cuuint64_t)ct_in * 2, (cuuint64_t)16 * ct_in * 2, (cuuint64_t)256 * ct_in * 2,
                             (cuuint64_t)1024 * ct_in * 2};
    cuuint32_t box[5] = {64, 16, 17, 1, 1};
    int r = encode(&txr, 5, act1, dims, strides, box);
    if (r != SG_OK) return r;
  }
  if (sega == 2) conv2_swap2_kernel<2><<<grid, Conv2RCfg::kThreads, Conv2RCfg::kSmemBytes, stream>>>(txr, tw, to, p);
  else if (half) conv2_swap2_kernel<1, true><<<grid, Conv2RCfg::kThreads, Conv2RCfg::kSmemBytes, stream>>>(txr, tw, to, p);
  else conv2_swap2_kernel<1><<<grid, Conv2RCfg::kThreads, Conv2RCfg::kSmemBytes, stream>>>(txr, tw, to, p);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

template <int SEGA, bool HALF = false>
static int launch_conv1_fused(const float* x, const __nv_bfloat16* w1t, __nv_bfloat16* act1, int64_t batch, int* err,
                              cudaStream_t stream) {
  using Cfg = Conv1FCfg<SEGA>;
  CUtensorMap tx, tb, to;
  {
    // fp32 NCHW input: (w, h, c, n); box = all 64 columns x 10 rows x 3 channels of one image
    cuuint64_t dims[4] = {64, 64, 3, (cuuint64_t)batch};
    cuuint64_t strides[3] = {64 * 4, 64 * 64 * 4, 3 * 64 * 64 * 4};
    cuuint32_t box[4] = {64, 10, 3, 1};
    int r = encode(&tx, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (r != SG_OK) return r;
  }
  {
    cuuint64_t dims[2] = {128, 64};
    cuuint64_t strides[1] = {128 * 2};
    cuuint32_t box[2] = {16, 64};
    int r = encode(&tb, 2, w1t, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_32B);
    if (r != SG_OK) return r;
  }
  {
    const cuuint64_t ct = 64 * SEGA;
    cuuint64_t dims[5] = {ct, 16, 16, 4, (cuuint64_t)batch};
    cuuint64_t strides[4] = {ct * 2, 16 * ct * 2, 256 * ct * 2, 1024 * ct * 2};
    cuuint32_t box[5] = {64, 16, 2, 1, 1};
    int r = encode(&to, 5, act1, dims, strides, box);
    if (r != SG_OK) return r;
  }
  const int64_t tiles = batch * 8;
  const int64_t ctas = (int64_t)state().sm_count * (SEGA == 1 ? 2 : 1);
  int grid = (int)(tiles < ctas ? tiles : ctas);
  conv1_fused_kernel<SEGA, HALF><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tx, tb, to, (int)tiles, err);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // namespace d64
}  // namespace sg

extern "C" {

int sg_d64_init_attributes() {
  using namespace sg::d64;
  SG_CUDA(cudaFuncSetAttribute(conv2_swap2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2RCfg::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(conv2_swap2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2RCfg::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(conv_pair2_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pair2Cfg<256>::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(conv_pair2_kernel<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pair2Cfg<256>::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(conv1_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Conv1FCfg<1>::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(conv1_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Conv1FCfg<2>::kSmemBytes));
  // fp16 conv mode (SG_CONV_FP16): the default kernels with fp16 operands
  SG_CUDA(cudaFuncSetAttribute(conv1_fused_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Conv1FCfg<1>::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(conv2_swap2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2RCfg::kSmemBytes));
  SG_CUDA(cudaFuncSetAttribute(conv_pair2_kernel<1, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Pair2Cfg<256>::kSmemBytes));
  return SG_OK;
}

size_t sg_d64_packed_bytes(int conv_mode) { return sg::d64::packed_layout(conv_mode).total; }

size_t sg_d64_workspace_bytes(int64_t max_batch, int conv_mode) {
  if (max_batch < 1) max_batch = 1;
  return sg::d64::workspace_layout(max_batch, conv_mode).total;
}

int sg_d64_pack(const float* w1, const float* w2, const float* w3, const float* w4, const float* w5,
                const float* bn2_gamma, const float* bn2_beta, const float* bn2_mean, const float* bn2_var,
                const float* bn3_gamma, const float* bn3_beta, const float* bn3_mean, const float* bn3_var,
                const float* bn4_gamma, const float* bn4_beta, const float* bn4_mean, const float* bn4_var,
                float bn_eps, int conv_mode, void* packed, void* stream) {
  using namespace sg::d64;
  SG_READY();
  SG_REQUIRE(conv_mode == SG_CONV_BF16 || conv_mode == SG_CONV_BF16X3 || conv_mode == SG_CONV_FP16, "conv_mode");
  SG_REQUIRE(w1 && w2 && w3 && w4 && w5 && packed, "null weight pointer");
  SG_REQUIRE(((uintptr_t)packed & 1023) == 0, "packed buffer must be 1024-byte aligned");
  const PackedLayout L = packed_layout(conv_mode);
  cudaStream_t st = sg::as_stream(stream);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  PackAllArgs a;
  a.w1 = w1; a.w2 = w2; a.w3 = w3; a.w4 = w4; a.w5 = w5;
  const float* gs[3] = {bn2_gamma, bn3_gamma, bn4_gamma};
  const float* bs[3] = {bn2_beta, bn3_beta, bn4_beta};
  const float* ms[3] = {bn2_mean, bn3_mean, bn4_mean};
  const float* vs[3] = {bn2_var, bn3_var, bn4_var};
  const size_t sso[3] = {L.ss2, L.ss3, L.ss4}, gbo[3] = {L.gb2, L.gb3, L.gb4};
  for (int l = 0; l < 3; ++l) {
    SG_REQUIRE(gs[l] && bs[l] && ms[l] && vs[l], "null BatchNorm pointer");
    a.g[l] = gs[l]; a.b[l] = bs[l]; a.m[l] = ms[l]; a.v[l] = vs[l];
    a.ss[l] = reinterpret_cast<float*>(pk + sso[l]);
    a.gb[l] = reinterpret_cast<float*>(pk + gbo[l]);
  }
  a.p2 = reinterpret_cast<__nv_bfloat16*>(pk + L.w2);
  a.p3 = reinterpret_cast<__nv_bfloat16*>(pk + L.w3);
  a.p4 = reinterpret_cast<__nv_bfloat16*>(pk + L.w4);
  a.o1t = reinterpret_cast<__nv_bfloat16*>(pk + L.w1t);
  a.o1 = reinterpret_cast<float*>(pk + L.w1);
  a.o5 = reinterpret_cast<float*>(pk + L.w5);
  a.ident = reinterpret_cast<float*>(pk + L.ident);
  a.eps = bn_eps;
  a.nseg = L.nseg;
  a.half = conv_mode == SG_CONV_FP16;
  pack_all_kernel<<<1059, 256, 0, st>>>(a);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// layer 1..5 of the scoring pipeline.  bn_train != 0: layers 2..4 use batch statistics (train-mode BN),
// updating running_stats[2*(layer-2)] / [2*(layer-2)+1] (may be NULL) with `momentum`.
static int run_layer_impl(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode, int layer,
                          float* logit, float* prob, float* loss, int bn_train, float* const* running_stats,
                          float momentum, float eps, int32_t* status, void* stream) {
  using namespace sg::d64;
  SG_READY();
  SG_REQUIRE(conv_mode == SG_CONV_BF16 || conv_mode == SG_CONV_BF16X3 || conv_mode == SG_CONV_FP16, "conv_mode");
  SG_REQUIRE(packed && workspace, "null pointer");
  SG_REQUIRE(layer >= 1 && layer <= 5, "layer must be 1..5");
  SG_REQUIRE(batch >= 0 && batch <= (1 << 22), "batch out of range");
  SG_REQUIRE(((uintptr_t)packed & 1023) == 0 && ((uintptr_t)workspace & 1023) == 0,
             "packed/workspace must be 1024-byte aligned");
  if (batch == 0) return SG_OK;
  const PackedLayout P = packed_layout(conv_mode);
  const WorkspaceLayout W = workspace_layout(batch, conv_mode);
  cudaStream_t st = sg::as_stream(stream);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  // status words: [0] pipeline time-out role code, [1] fp16 overflow marker; the caller's (sticky, caller-cleared)
  // or the workspace's own (cleared by sg_d64_check)
  int* err = status ? reinterpret_cast<int*>(status) : reinterpret_cast<int*>(ws + W.flag);
  __nv_bfloat16* act1 = reinterpret_cast<__nv_bfloat16*>(ws + W.act1);
  __nv_bfloat16* act2 = reinterpret_cast<__nv_bfloat16*>(ws + W.act2);
  __nv_bfloat16* act3 = reinterpret_cast<__nv_bfloat16*>(ws + W.act3);
  __nv_bfloat16* act4 = reinterpret_cast<__nv_bfloat16*>(ws + W.act4);
  auto wq = [&](size_t off) { return reinterpret_cast<const __nv_bfloat16*>(pk + off); };
  auto fq = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  const float slope = bn_train ? 1.0f : 0.2f;
  const bool half = conv_mode == SG_CONV_FP16;
  // eval: folded BN scale | shift; train: identity (ones | zeros at +512) so the raw conv output is stored
  auto sc = [&](size_t off, int) { return fq(bn_train ? P.ident : off); };
  auto sh = [&](size_t off, int c) { return bn_train ? fq(P.ident) + 512 : fq(off) + c; };
  int r = SG_OK;
  switch (layer) {
    case 1:
      SG_REQUIRE(x != nullptr && ((uintptr_t)x & 15) == 0, "x must be a 16-byte aligned device pointer");
      if (half) return launch_conv1_fused<1, true>(x, wq(P.w1t), act1, batch, err, st);
      return (W.sega == 2) ? launch_conv1_fused<2>(x, wq(P.w1t), act1, batch, err, st)
                           : launch_conv1_fused<1>(x, wq(P.w1t), act1, batch, err, st);
    case 2:
      r = launch_conv2_swap(act1, wq(P.w2), sc(P.ss2, 128), sh(P.ss2, 128), act2, batch, P.nseg, W.sega, slope, err, st, half);
      break;
    case 3:
      r = launch_conv_pair2<256>(act2, wq(P.w3), sc(P.ss3, 256), sh(P.ss3, 256), act3, batch, 16, 128, 256, P.nseg, W.sega, 1,
                                 slope, err, st, half);
      break;
    case 4:
      r = launch_conv_pair2<256>(act3, wq(P.w4), sc(P.ss4, 512), sh(P.ss4, 512), act4, batch, 8, 256, 512, P.nseg, W.sega, 0,
                                 slope, err, st, half);
      break;
    default:
      // one CTA per sample at every batch size (one summation order everywhere: shards, chunks and single batches agree bit
      // for bit); the warp-per-sample form took 17.8 us at B = 128 against 4 us
      head_kernel<<<(unsigned)batch, 256, 0, st>>>(act4, fq(P.w5), W.sega, half, logit, prob, loss, err);
      SG_LAUNCH_CHECK();
      return SG_OK;
  }
  if (r != SG_OK || !bn_train) return r;
  // train-mode BN on the raw conv output of layer 2..4
  const int c = layer == 2 ? 128 : layer == 3 ? 256 : 512;
  const int64_t rows = batch * (layer == 2 ? 256 : layer == 3 ? 64 : 16);
  __nv_bfloat16* act = layer == 2 ? act2 : layer == 3 ? act3 : act4;
  const size_t gb = layer == 2 ? P.gb2 : layer == 3 ? P.gb3 : P.gb4;
  double* part = reinterpret_cast<double*>(ws + W.bnpart);
  float* ss = reinterpret_cast<float*>(ws + W.bnss);
  int blocks = (int)(rows < kBnBlocks ? rows : kBnBlocks);
  bn_stats_kernel<<<blocks, 256, 0, st>>>(act, rows, c, W.sega, half, part);
  SG_LAUNCH_CHECK();
  float* rm = running_stats ? running_stats[2 * (layer - 2)] : nullptr;
  float* rv = running_stats ? running_stats[2 * (layer - 2) + 1] : nullptr;
  // fp16 mode: the new statistics are parked in the workspace until the head has shown that nothing overflowed
  float* pend = reinterpret_cast<float*>(ws + W.bnrun) + (size_t)2 * (layer - 2) * 512;
  bn_finalize_kernel<<<(c + 7) / 8, 256, 0, st>>>(part, blocks, rows, c, fq(gb), eps, momentum, rm, rv, half ? pend : rm,
                                                  half ? pend + 512 : rv, ss);
  SG_LAUNCH_CHECK();
  int64_t ab = sg::ceil_div(rows * (c / 8), 256);
  if (ab > (int64_t)sg::state().sm_count * 16) ab = (int64_t)sg::state().sm_count * 16;
  bn_apply_kernel<<<(unsigned)ab, 256, 0, st>>>(act, rows, c, W.sega, half, ss);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_d64_run_layer(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode, int layer,
                     float* logit, float* prob, float* loss, void* stream) {
  return run_layer_impl(x, batch, packed, workspace, conv_mode, layer, logit, prob, loss, 0, nullptr, 0.f, 0.f, nullptr, stream);
}

int sg_d64_score_train_status(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode,
                              float* bn2_running_mean, float* bn2_running_var, float* bn3_running_mean,
                              float* bn3_running_var, float* bn4_running_mean, float* bn4_running_var, float momentum,
                              float bn_eps, float* logit, float* prob, float* loss, int32_t* status2, void* stream) {
  using namespace sg::d64;
  SG_READY();
  SG_REQUIRE(x && packed && workspace, "null pointer");
  SG_REQUIRE(batch >= 1, "train-mode BatchNorm needs at least one sample");
  float* rs[6] = {bn2_running_mean, bn2_running_var, bn3_running_mean, bn3_running_var, bn4_running_mean, bn4_running_var};
  for (int layer = 1; layer <= 5; ++layer) {
    const int r = run_layer_impl(x, batch, packed, workspace, conv_mode, layer, logit, prob, loss, 1, rs, momentum, bn_eps,
                                 status2, stream);
    if (r != SG_OK) return r;
  }
  if (conv_mode == SG_CONV_FP16) {
    const WorkspaceLayout W = workspace_layout(batch, conv_mode);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    BnCommitArgs a;
    for (int i = 0; i < 6; ++i) a.dst[i] = rs[i];
    bn_commit_kernel<<<6, 512, 0, sg::as_stream(stream)>>>(reinterpret_cast<const float*>(ws + W.bnrun), a,
                                                           status2 ? status2 : reinterpret_cast<const int*>(ws + W.flag));
    SG_LAUNCH_CHECK();
  }
  return SG_OK;
}

int sg_d64_score_train(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode,
                       float* bn2_running_mean, float* bn2_running_var, float* bn3_running_mean,
                       float* bn3_running_var, float* bn4_running_mean, float* bn4_running_var, float momentum,
                       float bn_eps, float* logit, float* prob, float* loss, void* stream) {
  return sg_d64_score_train_status(x, batch, packed, workspace, conv_mode, bn2_running_mean, bn2_running_var,
                                   bn3_running_mean, bn3_running_var, bn4_running_mean, bn4_running_var, momentum, bn_eps,
                                   logit, prob, loss, nullptr, stream);
}

int sg_d64_score_status(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode, float* logit,
                        float* prob, float* loss, int32_t* status2, void* stream) {
  SG_READY();
  SG_REQUIRE(x && packed && workspace, "null pointer");
  if (batch == 0) return SG_OK;
  for (int layer = 1; layer <= 5; ++layer) {
    const int r = run_layer_impl(x, batch, packed, workspace, conv_mode, layer, logit, prob, loss, 0, nullptr, 0.f, 0.f,
                                 status2, stream);
    if (r != SG_OK) return r;
  }
  return SG_OK;
}

int sg_d64_score(const float* x, int64_t batch, const void* packed, void* workspace, int conv_mode, float* logit,
                 float* prob, float* loss, void* stream) {
  return sg_d64_score_status(x, batch, packed, workspace, conv_mode, logit, prob, loss, nullptr, stream);
}

int sg_d64_check(const void* workspace, void* stream) {
  SG_READY();
  SG_REQUIRE(workspace != nullptr, "workspace");
  int flags[2] = {0, 0};
  SG_CUDA(cudaMemcpyAsync(flags, workspace, 8, cudaMemcpyDeviceToHost, sg::as_stream(stream)));
  SG_CUDA(cudaStreamSynchronize(sg::as_stream(stream)));
  if (flags[0] != 0 || flags[1] != 0)   // both words are sticky until reported here
    (void)cudaMemsetAsync(const_cast<void*>(workspace), 0, 8, sg::as_stream(stream));
  if (flags[0] != 0) {
    sg::set_error("conv pipeline timed out waiting on an mbarrier (role code %d: 1 producer, 2 mma, 3 mma-acc, 4 epilogue)", flags[0]);
    return SG_ECUDA;
  }
  if (flags[1] == sg::d64::kFp16OverflowMagic) {
    sg::set_error("non-finite logit in the fp16 conv mode: an activation exceeded 65504 (or the input is not finite); "
                  "score this discriminator with SG_CONV_BF16X3 ('fp32') or SG_CONV_BF16");
    return SG_EINVAL;
  }
  return SG_OK;
}

int sg_d64_read_activation(const void* workspace, int64_t batch, int conv_mode, int layer, float* out, void* stream) {
  using namespace sg::d64;
  SG_READY();
  SG_REQUIRE(workspace && out && layer >= 1 && layer <= 4 && batch > 0, "arguments");
  const WorkspaceLayout W = workspace_layout(batch, conv_mode);
  const uint8_t* ws = static_cast<const uint8_t*>(workspace);
  const size_t offs[5] = {0, W.act1, W.act2, W.act3, W.act4};
  const int s[5] = {0, 32, 16, 8, 4};
  const int c[5] = {0, 64, 128, 256, 512};
  const int64_t total = batch * c[layer] * s[layer] * s[layer];
  int blocks = (int)sg::ceil_div(total, 256);
  if (blocks > 65535) blocks = 65535;
  read_activation_kernel<<<blocks, 256, 0, sg::as_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(ws + offs[layer]), batch, s[layer], c[layer], W.sega, conv_mode == SG_CONV_FP16, layer != 4, out);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
