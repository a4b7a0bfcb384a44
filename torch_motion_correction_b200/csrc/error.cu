// Error reporting for the C ABI: thread-local last-error string.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

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

// ---- per-kernel device timing (bench.py's roofline: CUDA-event duration of individual launches) ----------------------
#include <map>
#include <mutex>
#include <string>
#include <vector>
namespace {
struct TimedLaunch {
  const char* kernel;
  cudaEvent_t start, stop;
};
std::atomic<int> g_timing_on{0};
std::mutex g_timing_mutex;
std::vector<TimedLaunch> g_timed;
thread_local cudaEvent_t tl_start = nullptr;
bool capturing(cudaStream_t stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone;
}
}  // namespace

void tmc_timing_begin(cudaStream_t stream) {
  tl_start = nullptr;
  if (!g_timing_on.load(std::memory_order_relaxed) || capturing(stream)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  tl_start = e;
}

void tmc_timing_end(const char* kernel, cudaStream_t stream) {
  if (tl_start == nullptr) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  std::lock_guard<std::mutex> lock(g_timing_mutex);
  g_timed.push_back({kernel, tl_start, e});
  tl_start = nullptr;
}

// enable != 0: start recording (drops earlier records); 0: stop recording (records are kept for the report)
TMC_API int tmc_kernel_timing(int enable) {
  std::lock_guard<std::mutex> lock(g_timing_mutex);
  if (enable) {
    for (auto& t : g_timed) {
      cudaEventDestroy(t.start);
      cudaEventDestroy(t.stop);
    }
    g_timed.clear();
  }
  g_timing_on.store(enable ? 1 : 0);
  return TMC_OK;
}

// "kernel,launches,total_ms\n" per timed kernel into buf (NUL-terminated, truncated to size); waits for the recorded
// events; returns the number of bytes the full report needs
TMC_API long tmc_kernel_timing_report(char* buf, long size) {
  std::lock_guard<std::mutex> lock(g_timing_mutex);
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& t : g_timed) {
    float ms = 0.f;
    if (cudaEventSynchronize(t.stop) == cudaSuccess && cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) {
      auto& a = agg[t.kernel];
      a.first += 1;
      a.second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s,%ld,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && size > 0) {
    const long n = (long)out.size() < size - 1 ? (long)out.size() : size - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (long)out.size() + 1;
}

TMC_API int tmc_version(void) { return 100; }  // 0.1.0

// number of SMs of the current device (for persistent-grid sizing on the host side)
TMC_API int tmc_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

// ---- small host -> device uploads that do not queue behind a large H2D copy ------------------------------------
// A cudaMemcpyAsync of a few KB shares the host-to-device copy engine with the (2.7 GB) upload of the NEXT movie of a
// pipelined run and would wait for it; a kernel reading the pinned host buffer through its device mapping does not.
namespace {
__global__ void upload_words_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

// host_pinned: page-locked host memory (cudaHostAlloc / torch pin_memory); nbytes a multiple of 4
TMC_API int tmc_upload_pinned(const void* host_pinned, void* dst, long nbytes, cudaStream_t stream) {
  TMC_CHECK_ARG(host_pinned && dst && nbytes >= 0 && nbytes % 4 == 0, "upload_pinned: bad arguments");
  if (nbytes == 0) return TMC_OK;
  void* mapped = nullptr;
  if (cudaHostGetDevicePointer(&mapped, const_cast<void*>(host_pinned), 0) != cudaSuccess || mapped == nullptr) {
    cudaGetLastError();  // not mapped: plain copy
    TMC_CUDA(cudaMemcpyAsync(dst, host_pinned, (size_t)nbytes, cudaMemcpyHostToDevice, stream));
    return TMC_OK;
  }
  const long n = nbytes / 4;
  const int blocks = (int)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64);
  upload_words_kernel<<<blocks, 256, 0, stream>>>((const unsigned*)mapped, (unsigned*)dst, n);
  tmc_count_launch();
  TMC_CHECK_LAUNCH("tmc_upload_pinned");
  return TMC_OK;
}
