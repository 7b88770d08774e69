// Host-side staging for the PCIe copy of a fp32 host dataset (end-to-end leg of `refine_dataset_by_loss`,
// "#strainer gan.py:364-392": the reference's DataLoader hands fp32 batches from host memory to the device).
//
// The default conv mode rounds every input pixel to fp16 (round to nearest even) when conv1 builds its tensor-core
// operand.  Doing that same rounding on the host BEFORE the copy halves the bytes that cross PCIe (24 576 instead of
// 49 152 B per 64x64 RGB sample), and the scores stay bit-identical: cvt.rn.f16.f32 on the device and VCVTPS2PH (RN) on
// the host are the same IEEE conversion (sg_f16_expand widens the fp16 bits back exactly, conv1 re-rounds to the same
// bits).  This is data movement for the copy, not a CPU form of any kernel: nothing is scored on the host.
//
// One 8 192-image chunk (403 MB of fp32) converts in ~4.2 ms on 12-16 host threads (97 GB/s of fp32 read, measured on the
// B200 box: tools/host_f16_probe.c), against 7.3 ms for its fp32 PCIe copy at 55 GB/s.
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/strainer_b200.h"

namespace sg {
void set_error(const char* fmt, ...);
}

namespace {

constexpr int64_t kBlock = 1 << 16;   // elements per work item (256 KB of fp32): dst + i * kBlock keeps dst's alignment

// scalar fp32 -> fp16, round to nearest even, overflow to inf, subnormals exact (the IEEE conversion the vector units do)
inline uint16_t f2h_scalar(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
  x &= 0x7fffffffu;
  if (x > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);          // NaN
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);         // >= 65520 rounds to infinity
  if (x < 0x38800000u) {                                            // below 2^-14: fp16 subnormal or zero
    float a;
    memcpy(&a, &x, 4);
    a += 0.5f;                                                      // the fp32 adder rounds to the fp16 subnormal grid
    uint32_t r;
    memcpy(&r, &a, 4);
    return (uint16_t)(sign | (r - 0x3f000000u));
  }
  const uint32_t odd = (x >> 13) & 1u;
  x += 0xc8000fffu + odd;                                           // rebias the exponent, round half to even
  return (uint16_t)(sign | (x >> 13));
}

void convert_scalar(const float* s, uint16_t* d, int64_t n) {
  for (int64_t i = 0; i < n; ++i) d[i] = f2h_scalar(s[i]);
}

__attribute__((target("avx2,f16c"))) void convert_avx2(const float* s, uint16_t* d, int64_t n) {
  int64_t i = 0;
  const bool nt = ((uintptr_t)d & 15) == 0;
  for (; i + 16 <= n; i += 16) {
    const __m128i a = _mm256_cvtps_ph(_mm256_loadu_ps(s + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    const __m128i b = _mm256_cvtps_ph(_mm256_loadu_ps(s + i + 8), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    if (nt) {
      _mm_stream_si128((__m128i*)(d + i), a);
      _mm_stream_si128((__m128i*)(d + i + 8), b);
    } else {
      _mm_storeu_si128((__m128i*)(d + i), a);
      _mm_storeu_si128((__m128i*)(d + i + 8), b);
    }
  }
  _mm_sfence();
  convert_scalar(s + i, d + i, n - i);
}

__attribute__((target("avx512f"))) void convert_avx512(const float* s, uint16_t* d, int64_t n) {
  int64_t i = 0;
  const bool nt = ((uintptr_t)d & 31) == 0;
  for (; i + 32 <= n; i += 32) {
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
  _mm_sfence();
  convert_scalar(s + i, d + i, n - i);
}

using ConvertFn = void (*)(const float*, uint16_t*, int64_t);

ConvertFn pick_convert(int isa) {
  // isa: 0 = best the CPU has, 1 = scalar, 2 = AVX2 + F16C, 3 = AVX-512F (tests force each form)
  __builtin_cpu_init();
  const bool has512 = __builtin_cpu_supports("avx512f");
  const bool has256 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("f16c");
  if (isa == 1) return convert_scalar;
  if (isa == 2) return has256 ? convert_avx2 : nullptr;
  if (isa == 3) return has512 ? convert_avx512 : nullptr;
  return has512 ? convert_avx512 : has256 ? convert_avx2 : convert_scalar;
}

// A small persistent pool: the calling thread works too; work items are claimed with one atomic add, so a thread that is
// descheduled (the Python main thread shares these cores) does not hold the others back.
struct Job {
  const float* src = nullptr;
  uint16_t* dst = nullptr;
  int64_t n = 0;
  ConvertFn fn = nullptr;
  std::atomic<int64_t> next{0};
};

struct Pool {
  std::mutex m;
  std::condition_variable start, done;
  std::vector<std::thread> workers;
  Job* job = nullptr;
  uint64_t generation = 0;
  int wanted = 0;      // workers that should join the current job
  int running = 0;     // workers that have not finished the current job yet
  pid_t pid = 0;
};

Pool* g_pool = nullptr;
std::mutex g_pool_mutex;   // one conversion at a time per process

void run_job(Job* j) {
  const int64_t blocks = (j->n + kBlock - 1) / kBlock;
  for (;;) {
    const int64_t b = j->next.fetch_add(1, std::memory_order_relaxed);
    if (b >= blocks) break;
    const int64_t i0 = b * kBlock, len = (j->n - i0 < kBlock) ? j->n - i0 : kBlock;
    j->fn(j->src + i0, j->dst + i0, len);
  }
}

void worker_main(Pool* p, int index) {
  uint64_t seen = 0;
  for (;;) {
    Job* j = nullptr;
    {
      std::unique_lock<std::mutex> lk(p->m);
      p->start.wait(lk, [&] { return p->generation != seen; });
      seen = p->generation;
      if (index < p->wanted) j = p->job;
    }
    if (j == nullptr) continue;
    run_job(j);
    {
      std::lock_guard<std::mutex> lk(p->m);
      if (--p->running == 0) p->done.notify_one();
    }
  }
}

int usable_cpus() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int c = CPU_COUNT(&set);
    if (c > 0) return c;
  }
  const long c = sysconf(_SC_NPROCESSORS_ONLN);
  return c > 0 ? (int)c : 1;
}

}  // namespace

extern "C" int sg_host_threads(void) { return usable_cpus(); }

extern "C" int sg_host_f32_to_f16(const float* h_src, int64_t count, uint16_t* h_dst, int threads, int isa) {
  if (count < 0 || (count > 0 && (h_src == nullptr || h_dst == nullptr))) {
    sg::set_error("invalid argument: sg_host_f32_to_f16 needs host pointers and count >= 0");
    return SG_EINVAL;
  }
  if (isa < 0 || isa > 3) {
    sg::set_error("invalid argument: isa must be 0 (auto), 1 (scalar), 2 (AVX2 + F16C) or 3 (AVX-512F)");
    return SG_EINVAL;
  }
  const ConvertFn fn = pick_convert(isa);
  if (fn == nullptr) {
    sg::set_error("sg_host_f32_to_f16: this CPU does not have the requested instruction set (isa = %d)", isa);
    return SG_EINVAL;
  }
  if (count == 0) return SG_OK;
  const int64_t blocks = (count + kBlock - 1) / kBlock;
  int t = threads > 0 ? threads : usable_cpus();
  if (t > 64) t = 64;
  if ((int64_t)t > blocks) t = (int)blocks;
  Job job;
  job.src = h_src;
  job.dst = h_dst;
  job.n = count;
  job.fn = fn;
  if (t <= 1) {
    run_job(&job);
    return SG_OK;
  }
  std::lock_guard<std::mutex> serial(g_pool_mutex);
  if (g_pool == nullptr || g_pool->pid != getpid()) g_pool = new Pool();   // after a fork the parent's threads are gone
  Pool* p = g_pool;
  p->pid = getpid();
  try {
    while ((int)p->workers.size() < t - 1) {
      const int index = (int)p->workers.size();
      p->workers.emplace_back(worker_main, p, index);
      p->workers.back().detach();
    }
  } catch (...) {
    // thread creation refused (cgroup limit): carry on with the workers that exist
  }
  const int helpers = (int)p->workers.size() < t - 1 ? (int)p->workers.size() : t - 1;
  {
    std::lock_guard<std::mutex> lk(p->m);
    p->job = &job;
    p->wanted = helpers;
    p->running = helpers;
    ++p->generation;
  }
  p->start.notify_all();
  run_job(&job);
  {
    std::unique_lock<std::mutex> lk(p->m);
    p->done.wait(lk, [&] { return p->running == 0; });
    p->job = nullptr;
  }
  return SG_OK;
}
