// Stack statistics for normalize_image (reference utils.py:49-84): scalar mean and unbiased
// std over the central box [y0,y1) x [x0,x1) of ALL frames.  The affine (x - mean) / std is
// NOT materialised: consumers (patch extraction, warp) apply it while loading.
#include "common.cuh"

namespace {

constexpr int kStatsThreads = 256;

__global__ void __launch_bounds__(kStatsThreads)
stats_partial_kernel(const float* __restrict__ image, int t, int h, int w, int y0, int y1, int x0, int x1,
                     double* __restrict__ partial) {
  const int bh = y1 - y0, bw = x1 - x0;
  const long rows = (long)t * bh;
  double s = 0.0, ss = 0.0;
  const bool vec = ((x0 & 3) == 0) && ((bw & 3) == 0) && ((w & 3) == 0) && ((reinterpret_cast<uintptr_t>(image) & 15) == 0);
  for (long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int f = r / bh, y = y0 + (int)(r % bh);
    const float* row = image + ((long)f * h + y) * w + x0;
    // fp32 partial per row chunk, folded into double: keeps the inner loop cheap and the error ~1e-7
    float fs = 0.f, fss = 0.f;
    if (vec) {
      const float4* row4 = reinterpret_cast<const float4*>(row);
      for (int i = threadIdx.x; i < bw / 4; i += kStatsThreads) {
        float4 v = __ldg(row4 + i);
        fs += (v.x + v.y) + (v.z + v.w);
        fss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    } else {
      for (int i = threadIdx.x; i < bw; i += kStatsThreads) {
        float v = __ldg(row + i);
        fs += v;
        fss += v * v;
      }
    }
    s += (double)fs;
    ss += (double)fss;
  }
  __shared__ double sh[2][kStatsThreads / 32];
  s = warp_sum(s);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s;
    sh[1][threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < kStatsThreads / 32; ++i) {
      a += sh[0][i];
      b += sh[1][i];
    }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
}

// moments[3] = {sum, sum of squares, count}: the exchange format for frame-split movies
__global__ void stats_moments_kernel(const double* __restrict__ partial, int nblocks, double count, double* __restrict__ moments) {
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    s += partial[2 * i];
    ss += partial[2 * i + 1];
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (threadIdx.x == 0) {
    moments[0] = s;
    moments[1] = ss;
    moments[2] = count;
  }
}

__global__ void moments_to_mean_std_kernel(const double* __restrict__ moments, float* __restrict__ mean_std) {
  const double s = moments[0], ss = moments[1], n = moments[2];
  const double mean = s / n;
  mean_std[0] = (float)mean;
  mean_std[1] = (float)sqrt((ss - s * mean) / (n - 1.0));
}

__global__ void stats_final_kernel(const double* __restrict__ partial, int nblocks, double count, float* __restrict__ mean_std) {
  // one warp, fixed order => deterministic
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    s += partial[2 * i];
    ss += partial[2 * i + 1];
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (threadIdx.x == 0) {
    double mean = s / count;
    double var = (ss - s * mean) / (count - 1.0);  // unbiased; NaN for count == 1 like torch
    mean_std[0] = (float)mean;
    mean_std[1] = (float)sqrt(var);
  }
}

}  // namespace

TMC_API int tmc_stack_stats_workspace_doubles(void) { return 2 * 148 * 8; }

// mean_std: device float[2] = {mean, unbiased std}; workspace: tmc_stack_stats_workspace_doubles() doubles
TMC_API int tmc_stack_stats(const float* image, int t, int h, int w, int y0, int y1, int x0, int x1, float* mean_std,
                            double* workspace, cudaStream_t stream) {
  TMC_CHECK_ARG(image && mean_std && workspace, "stack_stats: null pointer");
  TMC_CHECK_ARG(t >= 1 && h >= 1 && w >= 1, "stack_stats: bad shape (%d,%d,%d)", t, h, w);
  TMC_CHECK_ARG(0 <= y0 && y0 < y1 && y1 <= h && 0 <= x0 && x0 < x1 && x1 <= w, "stack_stats: empty or out-of-range box");
  const long rows = (long)t * (y1 - y0);
  int nblocks = (int)(rows < 148 * 8 ? rows : 148 * 8);
  TMC_TIMED("stats_partial_kernel", stream, stats_partial_kernel<<<nblocks, kStatsThreads, 0, stream>>>(image, t, h, w, y0, y1, x0, x1, workspace));
  TMC_TIMED("stats_final_kernel", stream, stats_final_kernel<<<1, 32, 0, stream>>>(workspace, nblocks, (double)rows * (x1 - x0), mean_std));
  TMC_CHECK_LAUNCH("tmc_stack_stats");
  return TMC_OK;
}

// Raw moments {sum, sum of squares, count} (device double[3]) of the central box of the local frames: ranks
// holding frame blocks of one movie all-reduce these (SUM) and call tmc_moments_to_mean_std.
TMC_API int tmc_stack_moments(const float* image, int t, int h, int w, int y0, int y1, int x0, int x1, double* moments,
                              double* workspace, cudaStream_t stream) {
  TMC_CHECK_ARG(image && moments && workspace, "stack_moments: null pointer");
  TMC_CHECK_ARG(t >= 1 && h >= 1 && w >= 1, "stack_moments: bad shape (%d,%d,%d)", t, h, w);
  TMC_CHECK_ARG(0 <= y0 && y0 < y1 && y1 <= h && 0 <= x0 && x0 < x1 && x1 <= w, "stack_moments: empty or out-of-range box");
  const long rows = (long)t * (y1 - y0);
  int nblocks = (int)(rows < 148 * 8 ? rows : 148 * 8);
  TMC_TIMED("stats_partial_kernel", stream, stats_partial_kernel<<<nblocks, kStatsThreads, 0, stream>>>(image, t, h, w, y0, y1, x0, x1, workspace));
  TMC_TIMED("stats_moments_kernel", stream, stats_moments_kernel<<<1, 32, 0, stream>>>(workspace, nblocks, (double)rows * (x1 - x0), moments));
  TMC_CHECK_LAUNCH("tmc_stack_moments");
  return TMC_OK;
}

TMC_API int tmc_moments_to_mean_std(const double* moments, float* mean_std, cudaStream_t stream) {
  TMC_CHECK_ARG(moments && mean_std, "moments_to_mean_std: null pointer");
  TMC_TIMED("moments_to_mean_std_kernel", stream, moments_to_mean_std_kernel<<<1, 1, 0, stream>>>(moments, mean_std));
  TMC_CHECK_LAUNCH("tmc_moments_to_mean_std");
  return TMC_OK;
}
