// Probe: does a CACHE-RESIDENT staging ring speed up the host-packed PCIe copy of fp32 host datasets?
//   mode A (the library's current form): a whole 8192-image chunk is rounded to fp16 with non-temporal stores into one of
//           three 201 MB pinned buffers, then copied with one cudaMemcpyAsync (host DRAM: 403 MB read + 201 MB written by
//           the pack + 201 MB read by the DMA per chunk);
//   mode B: pieces of P images are rounded with ordinary (cacheable) stores into a small pinned ring that fits the LLC and
//           are copied at once, so that the DMA reads can be served from the cache.
// Build: nvcc -O3 -Xcompiler -fopenmp,-mavx512f,-mf16c -o /tmp/ring_probe tools/host_pack_ring_probe.cu
// Run:   /tmp/ring_probe [images = 65536] [threads = 16]
#include <cuda_runtime.h>
#include <immintrin.h>
#include <omp.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } \
  } while (0)

static const int64_t kImg = 3 * 64 * 64;

static void convert(const float* s, uint16_t* d, int64_t n, bool nt) {
  for (int64_t i = 0; i < n; i += 32) {
    const __m256i a = _mm512_cvtps_ph(_mm512_loadu_ps(s + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    const __m256i b = _mm512_cvtps_ph(_mm512_loadu_ps(s + i + 16), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    if (nt) {
      _mm256_stream_si256((__m256i*)(d + i), a);
      _mm256_stream_si256((__m256i*)(d + i + 16), b);
    } else {
      _mm256_storeu_si256((__m256i*)(d + i), a);
      _mm256_storeu_si256((__m256i*)(d + i + 16), b);
    }
  }
  if (nt) _mm_sfence();
}

static void pack(const float* s, uint16_t* d, int64_t n, bool nt, int threads) {
  const int64_t blk = 16384;     // elements per work item
  const int64_t nb = (n + blk - 1) / blk;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t b = 0; b < nb; ++b) {
    const int64_t i0 = b * blk, len = n - i0 < blk ? n - i0 : blk;
    convert(s + i0, d + i0, len, nt);
  }
}

int main(int argc, char** argv) {
  const int64_t n_img = argc > 1 ? atoll(argv[1]) : 65536;
  const int threads = argc > 2 ? atoi(argv[2]) : 16;
  float* src;
  CK(cudaHostAlloc(&src, n_img * kImg * 4, cudaHostAllocDefault));
#pragma omp parallel for num_threads(16)
  for (int64_t i = 0; i < n_img * kImg; ++i) src[i] = (float)((i * 2654435761u) & 0xffff) / 65536.f - 0.5f;
  uint16_t* dev;
  const int64_t chunk = 8192;
  CK(cudaMalloc(&dev, 2 * chunk * kImg * 2));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };

  // raw fp32 copy, for reference
  {
    float* dev32;
    CK(cudaMalloc(&dev32, chunk * kImg * 4));
    const double t0 = now();
    for (int64_t c = 0; c < n_img / chunk; ++c) CK(cudaMemcpyAsync(dev32, src + c * chunk * kImg, chunk * kImg * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    const double dt = now() - t0;
    printf("{\"mode\": \"raw fp32 copy\", \"samples_per_s\": %.0f, \"gbs\": %.1f}\n", n_img / dt, n_img * kImg * 4 / dt / 1e9);
    CK(cudaFree(dev32));
  }
  // mode A
  for (int rep = 0; rep < 2; ++rep) {
    uint16_t* pin[3];
    cudaEvent_t ev[3];
    for (int k = 0; k < 3; ++k) {
      CK(cudaHostAlloc(&pin[k], chunk * kImg * 2, cudaHostAllocDefault));
      CK(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    }
    const double t0 = now();
    for (int64_t c = 0; c < n_img / chunk; ++c) {
      const int k = c % 3;
      if (c >= 3) CK(cudaEventSynchronize(ev[k]));
      pack(src + c * chunk * kImg, pin[k], chunk * kImg, true, threads);
      CK(cudaMemcpyAsync(dev + (c & 1) * chunk * kImg, pin[k], chunk * kImg * 2, cudaMemcpyHostToDevice, st));
      CK(cudaEventRecord(ev[k], st));
    }
    CK(cudaStreamSynchronize(st));
    const double dt = now() - t0;
    printf("{\"mode\": \"A: chunk, non-temporal stores, 3 x 201 MB\", \"threads\": %d, \"samples_per_s\": %.0f}\n", threads, n_img / dt);
    for (int k = 0; k < 3; ++k) { CK(cudaFreeHost(pin[k])); CK(cudaEventDestroy(ev[k])); }
  }
  // mode B
  const int pieces[] = {128, 256, 512, 1024, 2048};
  const int slots_opt[] = {3, 4, 8};
  for (int nt = 0; nt < 2; ++nt)
    for (int piece : pieces)
      for (int slots : slots_opt) {
        if ((int64_t)piece * slots * kImg * 2 > (96ll << 20)) continue;
        uint16_t* ring;
        CK(cudaHostAlloc(&ring, (int64_t)slots * piece * kImg * 2, cudaHostAllocDefault));
        std::vector<cudaEvent_t> ev(slots);
        for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const double t0 = now();
        const int64_t np = n_img / piece;
        for (int64_t p = 0; p < np; ++p) {
          const int k = p % slots;
          if (p >= slots) CK(cudaEventSynchronize(ev[k]));
          uint16_t* h = ring + (int64_t)k * piece * kImg;
          pack(src + p * piece * kImg, h, piece * kImg, nt != 0, threads);
          CK(cudaMemcpyAsync(dev + (p % (2 * chunk / piece)) * piece * kImg, h, piece * kImg * 2, cudaMemcpyHostToDevice, st));
          CK(cudaEventRecord(ev[k], st));
        }
        CK(cudaStreamSynchronize(st));
        const double dt = now() - t0;
        printf("{\"mode\": \"B: ring\", \"stores\": \"%s\", \"piece_images\": %d, \"slots\": %d, \"ring_mb\": %.1f, \"threads\": %d, "
               "\"samples_per_s\": %.0f}\n", nt ? "non-temporal" : "cacheable", piece, slots, slots * piece * kImg * 2 / 1048576.0,
               threads, n_img / dt);
        fflush(stdout);
        for (auto& e : ev) CK(cudaEventDestroy(e));
        CK(cudaFreeHost(ring));
      }
  return 0;
}
