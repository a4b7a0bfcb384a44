// Error reporting for the C ABI: thread-local last-error string.
#include "common.cuh"

#include <stdarg.h>

#include <atomic>

static thread_local char g_last_error[512] = "";

void tmc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

static std::atomic<long> g_launches{0};
void tmc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// kernels launched by this library since load (monotonic)
TMC_API long tmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

TMC_API const char* tmc_last_error(void) { return g_last_error; }

TMC_API int tmc_version(void) { return 100; }  // 0.1.0

// number of SMs of the current device (for persistent-grid sizing on the host side)
TMC_API int tmc_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}
