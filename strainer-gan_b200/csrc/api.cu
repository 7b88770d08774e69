// Library state, sg_init, error strings, synthetic data generator.
#include <stdarg.h>

#include "common.cuh"

namespace sg {

static thread_local char g_err[512] = "ok";
// One state per CUDA device: every entry point works on the CURRENT device of the calling thread (the Python layer
// makes the tensor's device current around each call), so a process may drive several GPUs.
constexpr int kMaxDevices = 64;
static DeviceState g_states[kMaxDevices];
static DeviceState g_none;

DeviceState& state() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { (void)cudaGetLastError(); return g_none; }
  return g_states[dev];
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_ready() {
  if (!state().ready) {
    set_error("sg_init(device) has not succeeded for the current CUDA device of this thread (no sm_100 device bound)");
    return SG_ENOINIT;
  }
  return SG_OK;
}

// ---- synthetic images: bit-identical to oracle/strainer_oracle.py::synth_images ----------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

__global__ void synth_images_kernel(float* __restrict__ out, int64_t start, int64_t count, uint32_t seed) {
  // one thread per 4 consecutive x of one (sample, c, y) row; 64x64x3 images
  const int64_t total4 = count * (3 * 64 * 64 / 4);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total4;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = t / 3072;
    const uint32_t r = (uint32_t)(t - s * 3072);
    const uint32_t c = r >> 10, y = (r >> 4) & 63u, x0 = (r & 15u) << 2;
    const uint64_t idx = (uint64_t)(start + s);
    const uint32_t key = mix32(seed ^ mix32((uint32_t)idx) ^ ((uint32_t)(idx >> 32) * 0x9E3779B1u));
    const bool noisy = (mix32(key ^ 0xA5A5A5A5u) % 5u) == 0u;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t x = x0 + j;
      uint32_t q;
      if (noisy) {
        const uint32_t e = (c * 64u + y) * 64u + x;
        q = mix32(key + 0x30000000u + e * 0x9E3779B9u) >> 8;
      } else {
        const uint32_t e8 = (c * 8u + (y >> 3)) * 8u + (x >> 3);
        const uint32_t e2 = (c * 32u + (y >> 1)) * 32u + (x >> 1);
        const uint64_t a = mix32(key + 0x10000000u + e8 * 0x9E3779B9u) >> 8;
        const uint64_t b = mix32(key + 0x20000000u + e2 * 0x9E3779B9u) >> 8;
        q = (uint32_t)((3ull * a + b) >> 2);
      }
      v[j] = __fsub_rn(__fmul_rn((float)q, 1.1920928955078125e-07f), 1.0f);  // q * 2^-23 - 1 (exact)
    }
    reinterpret_cast<float4*>(out)[t] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace sg

extern "C" {

int sg_version(void) { return 100; }

const char* sg_last_error_string(void) { return sg::g_err; }

long long sg_launch_count(void) { return (long long)__atomic_load_n(&sg::g_launches, __ATOMIC_RELAXED); }

int sg_sm_count(void) { return sg::state().sm_count; }

int sg_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    sg::set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    (void)cudaGetLastError();
    return SG_EARCH;
  }
  SG_REQUIRE(device >= 0 && device < count, "device index out of range");
  cudaDeviceProp prop;
  SG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    sg::set_error("device %d is sm_%d%d; libstrainer_b200 is built for sm_100a only", device, prop.major,
                  prop.minor);
    return SG_EARCH;
  }
  SG_REQUIRE(device < sg::kMaxDevices, "device index out of range");
  // the caller's current device is left as it was: attributes are set with `device` current, then it is restored
  int prev = -1;
  SG_CUDA(cudaGetDevice(&prev));
  struct Restore { int d; ~Restore() { if (d >= 0) (void)cudaSetDevice(d); } } restore{prev == device ? -1 : prev};
  SG_CUDA(cudaSetDevice(device));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    sg::set_error("cuTensorMapEncodeTiled not available from the driver");
    return SG_ECUDA;
  }
  sg::DeviceState& st = sg::g_states[device];
  st.device = device;
  st.sm_count = prop.multiProcessorCount;
  st.encode_tiled = fn;
  st.ready = true;
  int r = sg_d64_init_attributes();
  if (r == SG_OK) r = sg_ae_init_attributes();
  if (r == SG_OK) r = sg_select_init_attributes();
  if (r == SG_OK) r = sg_ae_tc_init_attributes();
  if (r == SG_OK) r = sg_dbscan_init_attributes();
  if (r == SG_OK) r = sg_sort_init_attributes();
  if (r == SG_OK) r = sg_gemm_init_attributes();
  if (r == SG_OK) r = sg_d64_train_init_attributes();
  if (r != SG_OK) { st.ready = false; return r; }
  return SG_OK;
}

int sg_synth_images(float* out, int64_t start, int64_t count, uint32_t seed, void* stream) {
  SG_READY();
  SG_REQUIRE(out != nullptr && count >= 0 && start >= 0, "out/start/count");
  if (count == 0) return SG_OK;
  const int64_t total4 = count * 3072;
  int blocks = (int)sg::ceil_div(total4, 256);
  const int maxb = sg::state().sm_count * 16;
  if (blocks > maxb) blocks = maxb;
  sg::synth_images_kernel<<<blocks, 256, 0, sg::as_stream(stream)>>>(out, start, count, seed);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
