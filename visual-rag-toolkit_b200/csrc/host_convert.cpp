// See host_convert.h.
#include "host_convert.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace vrag {

// Scalar conversion with round-to-nearest-even: normal numbers get the rounding bias added below the 10 kept mantissa
// bits (+1 when the kept part is odd), numbers below the fp16 normal range are rounded by one float addition that aligns
// them to the fp16 denormal grid (the FPU's own round-to-nearest-even does the work).
static inline uint16_t f32_bits_to_f16(uint32_t u) {
  const uint32_t sign = u & 0x80000000u;
  u ^= sign;
  uint16_t o;
  if (u >= ((127u + 16u) << 23)) {             // >= 65536, inf or nan
    o = u > (255u << 23) ? 0x7e00u : 0x7c00u;
  } else if (u < (113u << 23)) {               // below 2^-14: fp16 denormal or zero
    float f;
    memcpy(&f, &u, 4);
    const uint32_t magic_u = ((127u - 15u) + (23u - 10u) + 1u) << 23;
    float magic;
    memcpy(&magic, &magic_u, 4);
    f += magic;
    uint32_t r;
    memcpy(&r, &f, 4);
    o = static_cast<uint16_t>(r - magic_u);
  } else {
    const uint32_t odd = (u >> 13) & 1u;
    u += (static_cast<uint32_t>(15 - 127) << 23) + 0xfffu;
    u += odd;
    o = static_cast<uint16_t>(u >> 13);
  }
  return static_cast<uint16_t>(o | (sign >> 16));
}

static void convert_scalar(const float* src, uint16_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, src + i, 4);
    dst[i] = f32_bits_to_f16(u);
  }
}

__attribute__((target("avx,f16c"))) static void convert_f16c(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm256_cvtps_ph(a, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm256_cvtps_ph(b, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  }
  for (; i + 8 <= n; i += 8)
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i),
                     _mm256_cvtps_ph(_mm256_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  if (i < n) convert_scalar(src + i, dst + i, n - i);
}

void host_f32_to_f16(const float* src, uint16_t* dst, size_t n, bool force_scalar) {
  static const bool have_f16c = __builtin_cpu_supports("f16c") && __builtin_cpu_supports("avx");
  if (have_f16c && !force_scalar) convert_f16c(src, dst, n);
  else convert_scalar(src, dst, n);
}

// ------------------------------------------------------------------------------------------------ worker pool
namespace {
struct Pool {
  std::mutex call_mu;   // one parallel_for at a time
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::vector<std::thread> workers;
  const std::function<void(int)>* job = nullptr;
  int n_tasks = 0, next = 0, pending = 0;
  unsigned long long gen = 0;
  int n_threads = 1;

  Pool() {
    int n = static_cast<int>(std::thread::hardware_concurrency());
    if (n <= 0) n = 1;
    n = std::min(n, 8);
    if (const char* e = getenv("VRAG_HOST_THREADS")) n = std::max(1, std::min(64, atoi(e)));
    n_threads = n;
    for (int i = 1; i < n; ++i) workers.emplace_back([this] { loop(); });
    for (auto& t : workers) t.detach();   // the pool lives as long as the process
  }
  bool take(int* i) {   // mu held
    if (next >= n_tasks) return false;
    *i = next++;
    return true;
  }
  void run_tasks(std::unique_lock<std::mutex>& lk) {
    int i;
    while (take(&i)) {
      const std::function<void(int)>* f = job;
      lk.unlock();
      (*f)(i);
      lk.lock();
      if (--pending == 0) cv_done.notify_all();
    }
  }
  void loop() {
    std::unique_lock<std::mutex> lk(mu);
    unsigned long long seen = 0;
    for (;;) {
      cv_work.wait(lk, [&] { return gen != seen; });
      seen = gen;
      run_tasks(lk);
    }
  }
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    if (n == 1 || n_threads == 1) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    std::lock_guard<std::mutex> call(call_mu);
    std::unique_lock<std::mutex> lk(mu);
    job = &fn;
    n_tasks = n;
    next = 0;
    pending = n;
    ++gen;
    cv_work.notify_all();
    run_tasks(lk);
    cv_done.wait(lk, [&] { return pending == 0; });
    job = nullptr;
    n_tasks = next = 0;
  }
};
// The pool is created on first use and never destroyed (its detached workers may outlive static destruction). A forked
// child has none of the parent's threads: it starts over with a fresh pool on its first use (the old object is leaked; its
// mutexes may have been held by threads that no longer exist).
std::atomic<Pool*> g_pool{nullptr};
std::once_flag g_atfork_once;
std::mutex g_pool_mu;
void forget_pool_in_child() {
  g_pool.store(nullptr, std::memory_order_relaxed);
  new (&g_pool_mu) std::mutex();
}
Pool& pool() {
  Pool* p = g_pool.load(std::memory_order_acquire);
  if (p) return *p;
  std::call_once(g_atfork_once, [] { pthread_atfork(nullptr, nullptr, forget_pool_in_child); });
  std::lock_guard<std::mutex> lk(g_pool_mu);
  p = g_pool.load(std::memory_order_acquire);
  if (!p) {
    p = new Pool();
    g_pool.store(p, std::memory_order_release);
  }
  return *p;
}
}  // namespace

void host_parallel_for(int n_tasks, const std::function<void(int)>& fn) { pool().parallel_for(n_tasks, fn); }
int host_pool_threads() { return pool().n_threads; }

}  // namespace vrag
