// host fp32 -> fp16 conversion rate with T threads (F16C), buffers of 403 MB like one scoring chunk
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
typedef struct { const float* s; uint16_t* d; size_t n; } job_t;
__attribute__((target("avx2,f16c"))) static void* work(void* p) {
  job_t* j = (job_t*)p;
  size_t i = 0;
  for (; i + 32 <= j->n; i += 32) {
    __m256 a = _mm256_loadu_ps(j->s + i), b = _mm256_loadu_ps(j->s + i + 8), c = _mm256_loadu_ps(j->s + i + 16), d = _mm256_loadu_ps(j->s + i + 24);
    _mm_stream_si128((__m128i*)(j->d + i), _mm256_cvtps_ph(a, 0));
    _mm_stream_si128((__m128i*)(j->d + i + 8), _mm256_cvtps_ph(b, 0));
    _mm_stream_si128((__m128i*)(j->d + i + 16), _mm256_cvtps_ph(c, 0));
    _mm_stream_si128((__m128i*)(j->d + i + 24), _mm256_cvtps_ph(d, 0));
  }
  return 0;
}
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv) {
  size_t n = (size_t)8192 * 12288;
  float* s = aligned_alloc(4096, n * 4 * 2);
  uint16_t* d = aligned_alloc(4096, n * 2);
  for (size_t i = 0; i < 2 * n; ++i) s[i] = (float)(i % 1000) * 1e-3f - 0.5f;
  memset(d, 0, n * 2);
  printf("f16c=%d avx2=%d avx512f=%d\n", __builtin_cpu_supports("f16c"), __builtin_cpu_supports("avx2"), __builtin_cpu_supports("avx512f"));
  int Ts[] = {1, 2, 4, 8, 12, 16, 24, 32};
  for (int ti = 0; ti < 8; ++ti) {
    int T = Ts[ti];
    pthread_t th[64]; job_t jb[64];
    double best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      const float* src = s + (rep & 1) * n;
      double t0 = now();
      for (int t = 0; t < T; ++t) { size_t a = n / T * t / 32 * 32, b = (t == T - 1) ? n : n / T * (t + 1) / 32 * 32; jb[t] = (job_t){src + a, d + a, b - a}; pthread_create(&th[t], 0, work, &jb[t]); }
      for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
      double dt = now() - t0; if (dt < best) best = dt;
    }
    printf("threads %2d: %.2f ms per 8192-image chunk, %.1f GB/s fp32 read\n", T, best * 1e3, n * 4 / best * 1e-9);
  }
  return 0;
}
