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
