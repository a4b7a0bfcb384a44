// Deformation-field warp of a movie stack: per-pixel bicubic lookup of the per-frame shift
// lattice (reflection padding) followed by a bicubic gather of the frame (border-clamped
// taps, zero outside), either written as a (t,h,w) stack or accumulated over frames into one
// (h,w) sum without materialising the warped stack.
//
// Replaces the reference's correct_motion.py:81-185 (`_correct_frame`, `get_pixel_shifts`) and
// torch_image_interpolation.sample_image_2d / F.grid_sample(bicubic, align_corners=True)
// underneath it.  The fp32 coordinate round trip
// pixel -> [-1,1] -> pixel of grid_sample is reproduced operation by operation (no FMA
// contraction on that chain) because it alone costs ~3.5e-5 relative L2 otherwise.
#include "common.cuh"

namespace {

constexpr float kA = -0.75f;  // Keys cubic convolution parameter used by ATen's bicubic

// ATen get_cubic_upsample_coefficients
__device__ __forceinline__ void cubic_weights(float t, float (&w)[4]) {
  // ((A x - 5A) x + 8A) x - 4A   and   ((A + 2) x - (A + 3)) x x + 1, three FMAs each
  float x = t + 1.0f;
  w[0] = fmaf(fmaf(fmaf(kA, x, -5.0f * kA), x, 8.0f * kA), x, -4.0f * kA);
  x = t;
  w[1] = fmaf(fmaf(kA + 2.0f, x, -(kA + 3.0f)) * x, x, 1.0f);
  x = 1.0f - t;
  w[2] = fmaf(fmaf(kA + 2.0f, x, -(kA + 3.0f)) * x, x, 1.0f);
  x = 2.0f - t;
  w[3] = fmaf(fmaf(fmaf(kA, x, -5.0f * kA), x, 8.0f * kA), x, -4.0f * kA);
}

// c / d for a divisor known per launch: q0 = c * (1/d), one Newton step on the residual.  Correctly
// rounded for all but vanishingly rare operands -- a third of the instructions of __fdiv_rn.
struct Divisor {
  float d, rcp;
};
__device__ __forceinline__ Divisor make_divisor(float d) {
  Divisor v;
  v.d = d;
  v.rcp = __frcp_rn(d);
  return v;
}
__device__ __forceinline__ float div_by(float c, const Divisor& v) {
  const float q0 = __fmul_rn(c, v.rcp);
  const float r = fmaf(-q0, v.d, c);
  return fmaf(r, v.rcp, q0);
}

// grid_round_trip with the divisor prepared once per thread
struct RoundTrip {
  Divisor denom;  // 0.5 n - 0.5
  float scale;    // n - 1
};
__device__ __forceinline__ RoundTrip make_round_trip(int n) {
  RoundTrip r;
  r.denom = make_divisor(__fsub_rn(__fmul_rn(0.5f, (float)n), 0.5f));
  r.scale = (float)(n - 1);
  return r;
}
__device__ __forceinline__ float round_trip(float c, const RoundTrip& r) {
  const float g = __fsub_rn(div_by(c, r.denom), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), r.scale);
}

// array coordinate -> grid_sample [-1,1] -> unnormalised source coordinate, as fp32 ops:
//   g = c / (0.5*n - 0.5) - 1          (array_to_grid_sample, align_corners=True)
//   u = ((g + 1) / 2) * (n - 1)        (grid_sampler_unnormalize)
__device__ __forceinline__ float grid_round_trip(float c, int n) {
  float denom = __fsub_rn(__fmul_rn(0.5f, (float)n), 0.5f);
  float g = __fsub_rn(__fdiv_rn(c, denom), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(n - 1));
}

// ATen reflect_coordinates(align_corners=True) + clip, on an integer index
__device__ __forceinline__ int reflect_index(int i, int n) {
  if (n <= 1) return 0;
  int span = n - 1;
  int a = i < 0 ? -i : i;
  int flips = a / span;
  int extra = a - flips * span;
  int r = (flips & 1) ? span - extra : extra;
  return min(max(r, 0), n - 1);
}

struct LatticeAxis {
  int j[4];    // reflected tap indices
  float w[4];  // cubic weights
  int i0;      // floor index (for "same taps" tests)
};

// lattice lookup coordinate of pixel index p along an axis of image length n, lattice length L
// (correct_motion.py:162-172): q = (p / (n-1)) * (L-1), then the grid_sample round trip.
__device__ __forceinline__ LatticeAxis lattice_axis(int p, int n, int L) {
  LatticeAxis a;
  float q = __fmul_rn(__fdiv_rn((float)p, (float)(n - 1)), (float)(L - 1));
  float u = grid_round_trip(q, L);
  float f = floorf(u);
  cubic_weights(__fsub_rn(u, f), a.w);
  a.i0 = (int)f;
#pragma unroll
  for (int k = 0; k < 4; ++k) a.j[k] = reflect_index(a.i0 - 1 + k, L);
  return a;
}

struct ImageAxis {
  int j[4];
  float w[4];
  bool inside;
};

// sampling coordinate c = p + shift along an axis of length n (sample_image_2d semantics)
__device__ __forceinline__ ImageAxis image_axis(float c, int n) {
  ImageAxis a;
  a.inside = (c >= 0.0f) && (c <= (float)(n - 1));
  float u = grid_round_trip(c, n);
  float f = floorf(u);
  cubic_weights(__fsub_rn(u, f), a.w);
  // clamp in float first: a wild shift must not overflow the int conversion
  int i0 = (int)fminf(fmaxf(f, -4.0f), (float)(n + 4));
#pragma unroll
  for (int k = 0; k < 4; ++k) a.j[k] = min(max(i0 - 1 + k, 0), n - 1);
  return a;
}

__device__ __forceinline__ float gather_bicubic(const float* __restrict__ frame, int w, const ImageAxis& ay, const ImageAxis& ax) {
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* row = frame + (long)ay.j[r] * w;
    float v = ax.w[0] * __ldg(row + ax.j[0]) + ax.w[1] * __ldg(row + ax.j[1]) + ax.w[2] * __ldg(row + ax.j[2]) +
              ax.w[3] * __ldg(row + ax.j[3]);
    acc += ay.w[r] * v;
  }
  return acc;
}

constexpr int kTileX = 128;  // threads along x
constexpr int kTileYGroups = 8;
constexpr int kRowsPerThread = 1;

// Stage 1: interpolate every lattice row along x once per (frame, channel, lattice row, image
// column): RX[f][ch][a][x] = sum_b wx_b(x) * L[f][ch][a][jx_b(x)].  Same x-then-y order as ATen.
__global__ void lattice_xinterp_kernel(const float* __restrict__ lattice, int T, int lh, int lw, int W, float* __restrict__ rx) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const LatticeAxis lx = lattice_axis(x, W, lw);
  const long rows = (long)T * 2 * lh;
  for (long r = blockIdx.y; r < rows; r += gridDim.y) {
    const float* p = lattice + r * lw;
    rx[r * W + x] = lx.w[0] * __ldg(p + lx.j[0]) + lx.w[1] * __ldg(p + lx.j[1]) + lx.w[2] * __ldg(p + lx.j[2]) +
                    lx.w[3] * __ldg(p + lx.j[3]);
  }
}

// Stage 2: one thread = kRowsPerThread vertically adjacent output pixels of one column.
template <bool WRITE_STACK, bool WRITE_SUM, bool NORMALISE>
__global__ void __launch_bounds__(kTileX* kTileYGroups, 1)
warp_lattice_kernel(const float* __restrict__ image, int T, int H, int W, const float* __restrict__ rx, int lh,
                    float pixel_spacing, const float* __restrict__ mean_std, float* __restrict__ out_stack,
                    float* __restrict__ out_sum, int accumulate_sum) {
  const int x = blockIdx.x * kTileX + threadIdx.x;
  const int y_base = (blockIdx.y * kTileYGroups + threadIdx.y) * kRowsPerThread;
  if (x >= W || y_base >= H) return;

  int jy[kRowsPerThread][4];
  float wy[kRowsPerThread][4];
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r) {
    LatticeAxis a = lattice_axis(min(y_base + r, H - 1), H, lh);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      jy[r][k] = a.j[k] * W;
      wy[r][k] = a.w[k];
    }
  }
  float mean = 0.f, inv_std = 1.f;
  if (NORMALISE) {
    mean = __ldg(mean_std);
    inv_std = 1.0f / __ldg(mean_std + 1);
  }
  float acc[kRowsPerThread];
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r) acc[r] = 0.f;

  const RoundTrip rty = make_round_trip(H), rtx = make_round_trip(W);
  const float inv_px = 1.0f / pixel_spacing;
  const float hmax = (float)(H - 1), wmax = (float)(W - 1);
  const unsigned rx_plane = (unsigned)lh * (unsigned)W;
  const float* rx_frame = rx + x;
  const float* frame = image;
  for (int f = 0; f < T; ++f, rx_frame += 2 * (size_t)rx_plane, frame += (size_t)H * W) {
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
      const int y = y_base + r;
      if (y < H) {
        const float* Ry = rx_frame;
        const float* Rx = rx_frame + rx_plane;
        float sy = wy[r][0] * __ldg(Ry + jy[r][0]);
        sy = fmaf(wy[r][1], __ldg(Ry + jy[r][1]), sy);
        sy = fmaf(wy[r][2], __ldg(Ry + jy[r][2]), sy);
        sy = fmaf(wy[r][3], __ldg(Ry + jy[r][3]), sy);
        float sx = wy[r][0] * __ldg(Rx + jy[r][0]);
        sx = fmaf(wy[r][1], __ldg(Rx + jy[r][1]), sx);
        sx = fmaf(wy[r][2], __ldg(Rx + jy[r][2]), sx);
        sx = fmaf(wy[r][3], __ldg(Rx + jy[r][3]), sx);
        // Angstrom -> px, then pixel_grid + pixel_shifts (two roundings, like the reference)
        const float cy = __fadd_rn((float)y, __fmul_rn(sy, inv_px));
        const float cx = __fadd_rn((float)x, __fmul_rn(sx, inv_px));
        float v = 0.f;
        if (cy >= 0.0f && cy <= hmax && cx >= 0.0f && cx <= wmax) {
          const float uy = round_trip(cy, rty), ux = round_trip(cx, rtx);
          const float fy = floorf(uy), fx = floorf(ux);
          float ay[4], ax[4];
          cubic_weights(__fsub_rn(uy, fy), ay);
          cubic_weights(__fsub_rn(ux, fx), ax);
          const int iy = (int)fy, ix = (int)fx;
          if (iy >= 1 && iy <= H - 3 && ix >= 1 && ix <= W - 3) {
            // interior: 4 row pointers, taps at immediate offsets
            const float* p = frame + (unsigned)((iy - 1) * W + (ix - 1));
#pragma unroll
            for (int a = 0; a < 4; ++a, p += W) {
              float row = ax[0] * __ldg(p);
              row = fmaf(ax[1], __ldg(p + 1), row);
              row = fmaf(ax[2], __ldg(p + 2), row);
              row = fmaf(ax[3], __ldg(p + 3), row);
              v = fmaf(ay[a], row, v);
            }
          } else {
            // border: every tap index clamped individually (grid_sample padding_mode="border")
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const float* p = frame + (size_t)min(max(iy - 1 + a, 0), H - 1) * W;
              float row = ax[0] * __ldg(p + min(max(ix - 1, 0), W - 1));
              row = fmaf(ax[1], __ldg(p + min(max(ix, 0), W - 1)), row);
              row = fmaf(ax[2], __ldg(p + min(max(ix + 1, 0), W - 1)), row);
              row = fmaf(ax[3], __ldg(p + min(max(ix + 2, 0), W - 1)), row);
              v = fmaf(ay[a], row, v);
            }
          }
          if (NORMALISE) v = (v - mean) * inv_std;
        }
        if (WRITE_STACK) out_stack[((size_t)f * H + y) * W + x] = v;
        if (WRITE_SUM) acc[r] += v;
      }
    }
  }
  if (WRITE_SUM) {
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
      const int y = y_base + r;
      if (y < H) {
        float* o = out_sum + (long)y * W + x;
        *o = accumulate_sum ? (*o + acc[r]) : acc[r];
      }
    }
  }
}

// get_pixel_shifts (correct_motion.py:132-185): (h, w, 2) px shifts of one lattice (2, lh, lw)
__global__ void pixel_shifts_kernel(const float* __restrict__ lattice, int lh, int lw, int H, int W, float pixel_spacing,
                                    float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const LatticeAxis lx = lattice_axis(x, W, lw);
  const LatticeAxis ly = lattice_axis(y, H, lh);
  const long plane = (long)lh * lw;
  float s[2];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float* p = lattice + ch * plane + (long)ly.j[a] * lw;
      float v = lx.w[0] * __ldg(p + lx.j[0]) + lx.w[1] * __ldg(p + lx.j[1]) + lx.w[2] * __ldg(p + lx.j[2]) +
                lx.w[3] * __ldg(p + lx.j[3]);
      acc += ly.w[a] * v;
    }
    s[ch] = __fdiv_rn(acc, pixel_spacing);
  }
  float2* o = reinterpret_cast<float2*>(out) + (long)y * W + x;
  *o = make_float2(s[0], s[1]);
}

// correct_motion_slow (correct_motion.py:371-427) given per-pixel shifts computed elsewhere:
// out[f](y,x) = bicubic(frame_f; y + shift_y, x + shift_x) with shifts (t, h, w, 2) in px
__global__ void warp_dense_shifts_kernel(const float* __restrict__ image, int T, int H, int W, const float* __restrict__ shifts,
                                         float* __restrict__ out_stack) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  if (x >= W) return;
  const long i = ((long)f * H + y) * W + x;
  const float2 s = reinterpret_cast<const float2*>(shifts)[i];
  ImageAxis ay = image_axis(__fadd_rn((float)y, s.x), H);
  ImageAxis ax = image_axis(__fadd_rn((float)x, s.y), W);
  float v = 0.f;
  if (ay.inside && ax.inside) v = gather_bicubic(image + (long)f * H * W, W, ay, ax);
  out_stack[i] = v;
}

// normalised (t, y, x) coordinates of every pixel of frame f (correct_motion.py:401-409)
__global__ void pixel_tyx_kernel(int H, int W, int T, int frame_offset, int total_frames, float* __restrict__ tyx) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  if (x >= W) return;
  float* o = tyx + (((long)f * H + y) * W + x) * 3;
  o[0] = linspace01(f + frame_offset, total_frames);
  o[1] = __fdiv_rn((float)y, (float)(H - 1));
  o[2] = __fdiv_rn((float)x, (float)(W - 1));
}

}  // namespace

TMC_API long tmc_warp_workspace_floats(int t, int w, int lh) { return (long)t * 2 * lh * w; }

// image (t,h,w); lattice (t,2,lh,lw) in Angstrom; mean_std nullable device float[2] (input affine
// (v - mean) / std applied to the warped value, zero outside); out_stack (t,h,w) nullable;
// out_sum (h,w) nullable; accumulate_sum != 0 adds to out_sum; workspace:
// tmc_warp_workspace_floats(t, w, lh) floats.
TMC_API int tmc_warp_lattice(const float* image, int t, int h, int w, const float* lattice, int lh, int lw,
                             float pixel_spacing, const float* mean_std, float* out_stack, float* out_sum,
                             int accumulate_sum, float* workspace, cudaStream_t stream) {
  TMC_CHECK_ARG(image && lattice && workspace, "warp_lattice: null pointer");
  TMC_CHECK_ARG(out_stack || out_sum, "warp_lattice: need out_stack and/or out_sum");
  TMC_CHECK_ARG(t >= 1 && h >= 2 && w >= 2 && lh >= 1 && lw >= 1, "warp_lattice: bad shape t=%d h=%d w=%d lattice=%dx%d", t,
                h, w, lh, lw);
  TMC_CHECK_ARG((long)lh * w < (1l << 31), "warp_lattice: lattice rows x width overflows int");
  TMC_CHECK_ARG(pixel_spacing > 0.f, "warp_lattice: pixel_spacing must be > 0");
  {
    long rows = (long)t * 2 * lh;
    dim3 g1(tmc_div_up(w, 128), (unsigned)(rows < 4096 ? rows : 4096));
    lattice_xinterp_kernel<<<g1, 128, 0, stream>>>(lattice, t, lh, lw, w, workspace); tmc_count_launch();
  }
  dim3 block(kTileX, kTileYGroups);
  dim3 grid(tmc_div_up(w, kTileX), tmc_div_up(h, kTileYGroups * kRowsPerThread));
#define LAUNCH(S, A, N)                                                                                          \
  warp_lattice_kernel<S, A, N><<<grid, block, 0, stream>>>(image, t, h, w, workspace, lh, pixel_spacing, mean_std, \
                                                           out_stack, out_sum, accumulate_sum)
  const bool s = out_stack != nullptr, a = out_sum != nullptr, n = mean_std != nullptr;
  if (s && a && n) LAUNCH(true, true, true);
  else if (s && a) LAUNCH(true, true, false);
  else if (s && n) LAUNCH(true, false, true);
  else if (s) LAUNCH(true, false, false);
  else if (n) LAUNCH(false, true, true);
  else LAUNCH(false, true, false);
#undef LAUNCH
  tmc_count_launch();
  TMC_CHECK_LAUNCH("tmc_warp_lattice");
  return TMC_OK;
}

TMC_API int tmc_pixel_shifts(const float* lattice, int lh, int lw, int h, int w, float pixel_spacing, float* out,
                             cudaStream_t stream) {
  TMC_CHECK_ARG(lattice && out && lh >= 1 && lw >= 1 && h >= 2 && w >= 2 && pixel_spacing > 0.f, "pixel_shifts: bad arguments");
  dim3 grid(tmc_div_up(w, 128), h);
  pixel_shifts_kernel<<<grid, 128, 0, stream>>>(lattice, lh, lw, h, w, pixel_spacing, out); tmc_count_launch();
  TMC_CHECK_LAUNCH("tmc_pixel_shifts");
  return TMC_OK;
}

TMC_API int tmc_warp_dense_shifts(const float* image, int t, int h, int w, const float* shifts, float* out_stack,
                                  cudaStream_t stream) {
  TMC_CHECK_ARG(image && shifts && out_stack && t >= 1 && h >= 2 && w >= 2, "warp_dense_shifts: bad arguments");
  dim3 grid(tmc_div_up(w, 128), h, t);
  warp_dense_shifts_kernel<<<grid, 128, 0, stream>>>(image, t, h, w, shifts, out_stack); tmc_count_launch();
  TMC_CHECK_LAUNCH("tmc_warp_dense_shifts");
  return TMC_OK;
}

TMC_API int tmc_pixel_tyx(int h, int w, int t, int frame_offset, int total_frames, float* tyx, cudaStream_t stream) {
  TMC_CHECK_ARG(tyx && t >= 1 && h >= 2 && w >= 2 && total_frames >= t + frame_offset, "pixel_tyx: bad arguments");
  dim3 grid(tmc_div_up(w, 128), h, t);
  pixel_tyx_kernel<<<grid, 128, 0, stream>>>(h, w, t, frame_offset, total_frames, tyx); tmc_count_launch();
  TMC_CHECK_LAUNCH("tmc_pixel_tyx");
  return TMC_OK;
}
