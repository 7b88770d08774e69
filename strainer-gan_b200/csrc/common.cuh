// Shared host/device helpers for libstrainer_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/strainer_b200.h"

extern "C" int sg_d64_init_attributes();  // internal: raises the conv kernels' dynamic smem limit
extern "C" int sg_ae_init_attributes();
extern "C" int sg_select_init_attributes();
extern "C" int sg_ae_tc_init_attributes();
extern "C" int sg_dbscan_init_attributes();
extern "C" int sg_sort_init_attributes();
extern "C" int sg_gemm_init_attributes();
extern "C" int sg_d64_train_init_attributes();

namespace sg {

// ---- host side state / error plumbing -------------------------------------------------
struct DeviceState {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  void* encode_tiled = nullptr;  // PFN cuTensorMapEncodeTiled
};
DeviceState& state();
void set_error(const char* fmt, ...);
void count_launch();  // bumps the process-wide kernel-launch counter read by sg_launch_count()
int check_ready();  // SG_OK or SG_ENOINIT / SG_EARCH

#define SG_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      sg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SG_ECUDA;                                                             \
    }                                                                              \
  } while (0)

#define SG_REQUIRE(cond, msg)                                      \
  do {                                                             \
    if (!(cond)) {                                                 \
      sg::set_error("invalid argument: %s (%s)", msg, #cond);      \
      return SG_EINVAL;                                            \
    }                                                              \
  } while (0)

#define SG_READY()                      \
  do {                                  \
    int _r = sg::check_ready();         \
    if (_r != SG_OK) return _r;         \
  } while (0)

#define SG_LAUNCH_CHECK()                                                          \
  do {                                                                             \
    sg::count_launch();                                                            \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      sg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SG_ECUDA;                                                             \
    }                                                                              \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: the kernel may become resident while its predecessor in the stream still runs; it must
// execute pdl_wait() before it touches global memory the predecessor writes (or that a successor of the predecessor reads).
template <typename... KArgs, typename... Args>
static inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  count_launch();
  if (e != cudaSuccess) {
    set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e), __FILE__, __LINE__);
    return SG_ECUDA;
  }
  return SG_OK;
}
#define SG_LAUNCH_PDL(...)                        \
  do {                                            \
    const int _pr = sg::launch_pdl(__VA_ARGS__);  \
    if (_pr != SG_OK) return _pr;                 \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- device helpers ----------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
// Monotone fp32 -> uint32 radix key.  NaN (either sign) -> 0xFFFFFFFF (sorts last, as numpy /
// torch sort do), -0.0 is canonicalised to +0.0.
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;
  if (b == 0x80000000u) b = 0u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  if (k == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(b);
}

__device__ __forceinline__ bool cmp_apply(float v, float thr, int cmp) {
  bool r;
  switch (cmp & 3) {
    case SG_LT: r = v < thr; break;
    case SG_LE: r = v <= thr; break;
    case SG_GE: r = v >= thr; break;
    default: r = v > thr; break;
  }
  return (cmp & SG_NOT) ? !r : r;
}

// The two interpolation rules of the reference's quantile calls, with their exact rounding steps.
__device__ __forceinline__ float lerp_rule(float a, float b, float w, int kind) {
  const float d = __fsub_rn(b, a);
  if (kind == SG_LERP_NUMPY) {
    // numpy _lerp with fp32 operands: separate roundings, no contraction
    float r = __fadd_rn(a, __fmul_rn(d, w));
    if (w >= 0.5f) r = __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, w)));
    return r;
  }
  // torch.lerp: fused forms
  return (fabsf(w) < 0.5f) ? __fmaf_rn(w, d, a) : __fmaf_rn(-d, __fsub_rn(1.0f, w), b);
}

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_stream_i4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_i4(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Streams a fp32 vector through `f(value, index)` with 128-bit loads, U independent loads in flight per
// thread and consecutive threads on consecutive 16-B chunks (HBM-bound kernels: enough bytes in flight
// to cover the DRAM latency).  Visit ORDER is unspecified: only for order-independent reductions.
template <int U, typename F>
__device__ __forceinline__ void stream_f32_grid(const float* __restrict__ v, int64_t n, int64_t gtid, int64_t gsz, F&& f) {
  if ((reinterpret_cast<uintptr_t>(v) & 15) == 0) {
    const int64_t n4 = n >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(v);
    int64_t i = gtid;
    for (; i + (U - 1) * gsz < n4; i += U * gsz) {
      float4 q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) q[u] = ldg_stream4(v4 + i + u * gsz);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t e = (i + u * gsz) << 2;
        f(q[u].x, e); f(q[u].y, e + 1); f(q[u].z, e + 2); f(q[u].w, e + 3);
      }
    }
    for (; i < n4; i += gsz) {
      const float4 q = ldg_stream4(v4 + i);
      const int64_t e = i << 2;
      f(q.x, e); f(q.y, e + 1); f(q.z, e + 2); f(q.w, e + 3);
    }
    for (int64_t e = (n4 << 2) + gtid; e < n; e += gsz) f(v[e], e);
  } else {
    for (int64_t e = gtid; e < n; e += gsz) f(v[e], e);
  }
}
template <int U, typename F>
__device__ __forceinline__ void stream_f32(const float* __restrict__ v, int64_t n, F&& f) {
  stream_f32_grid<U>(v, n, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, f);
}

// Shared-memory counters in issue-bound streaming loops: atomicAdd on a generic pointer makes the compiler re-derive the
// shared window (S2UR SR_CgaCtaId / ULEA / IMAD, 4-5 instructions) at every call.  A 32-bit shared address laundered
// through an asm stays in a register, and red.shared needs no generic -> shared conversion.
__device__ __forceinline__ uint32_t smem_addr_reg(const void* p) {
  uint32_t a;
  asm volatile("mov.u32 %0, %1;" : "=r"(a) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return a;
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace sg
