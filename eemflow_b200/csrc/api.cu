// Library-level entry points: version, error text, device query.
#include <atomic>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace eem {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// Per-device cache; no CUDA call happens before the first kernel entry point is used, so
// loading the library in a forkserver/fork DataLoader worker does not create a context.
int sm_count() {
  static std::mutex mu;
  static int cache[64];
  static bool have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lock(mu);
  if (!have[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cache[dev] = n;
    have[dev] = true;
  }
  return cache[dev];
}

}  // namespace eem

extern "C" {

int eem_version(void) { return 100; }

const char* eem_last_error_string(void) { return eem::g_err; }

long long eem_launch_count(void) { return eem::g_launches.load(std::memory_order_relaxed); }

int eem_sm_count(void) {
  int n = eem::sm_count();
  if (n < 0) return eem::fail(EEM_ERR_CUDA, "eem_sm_count: no usable CUDA device");
  return n;
}

}  // extern "C"
