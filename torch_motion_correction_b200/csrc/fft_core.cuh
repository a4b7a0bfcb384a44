// Shared-memory complex FFT building block (power-of-two lengths 16..8192, fp32).
//
// Stockham autosort with register radix-16/8/4/2 butterflies: every pass reads N complex
// values from one shared-memory buffer, does radix-R DFTs in registers and writes them,
// already in order, to the other buffer.  A CTA transforms B sequences at once.  Shared
// memory is padded by one element per 16 so the strided butterfly stores are conflict-free.
// Twiddles come from a per-length table W_N^m = exp(-2 pi i m / N) (tmc_fft_twiddles) that
// stays L1/L2 resident.  The inverse transform is the forward one with re/im swapped at the
// load and store boundaries of the calling kernel.
#pragma once
#include "common.cuh"

namespace tmcfft {

__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 4); }
__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 4); }

// d * W_16^k, k in [0, 8)
__device__ __forceinline__ float2 mul_w16(float2 d, int k) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  switch (k) {
    case 0: return d;
    case 1: return make_float2(d.x * c1 + d.y * s1, d.y * c1 - d.x * s1);
    case 2: return make_float2((d.x + d.y) * h, (d.y - d.x) * h);
    case 3: return make_float2(d.x * s1 + d.y * c1, d.y * s1 - d.x * c1);
    case 4: return make_float2(d.y, -d.x);
    case 5: return make_float2(d.y * c1 - d.x * s1, -d.y * s1 - d.x * c1);
    case 6: return make_float2((d.y - d.x) * h, -(d.x + d.y) * h);
    default: return make_float2(d.y * s1 - d.x * c1, -d.y * c1 - d.x * s1);
  }
}

template <int R>
__device__ __forceinline__ constexpr int bitrev(int r) {
  int out = 0;
  for (int b = 1; b < R; b <<= 1) {
    out = (out << 1) | (r & 1);
    r >>= 1;
  }
  return out;
}

// in-register decimation-in-frequency DFT of R points; X[r] ends up in v[bitrev<R>(r)]
template <int R>
__device__ __forceinline__ void fft_reg(float2 (&v)[R]) {
#pragma unroll
  for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
      for (int k = 0; k < half; ++k) {
        float2 a = v[base + k], b = v[base + k + half];
        v[base + k] = cadd(a, b);
        v[base + k + half] = mul_w16(csub(a, b), k * (8 / half));
      }
    }
  }
}

// same, on a pointer into a register array (all indices are compile-time after unrolling)
template <int R>
__device__ __forceinline__ void fft_reg_ptr(float2* v) {
#pragma unroll
  for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
      for (int k = 0; k < half; ++k) {
        float2 a = v[base + k], b = v[base + k + half];
        v[base + k] = cadd(a, b);
        v[base + k + half] = mul_w16(csub(a, b), k * (8 / half));
      }
    }
  }
}

// One Stockham pass over B sequences of length N (padded stride), radix R, sub-transform NS.
template <int N, int R, int NS, int B, int THREADS>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out,
                                              const float2* __restrict__ tw) {
  constexpr int NB = N / R;  // butterflies per sequence
  constexpr int STRIDE = padded_len(N);
#pragma unroll 1
  for (int idx = threadIdx.x; idx < B * NB; idx += THREADS) {
    const int seq = idx / NB, j = idx % NB;
    const float2* src = in + seq * STRIDE;
    float2* dst = out + seq * STRIDE;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src[pad_idx(j + r * NB)];
    if (NS > 1) {
      const int k = j % NS;
      constexpr int STEP = N / (NS * R);  // W_{NS*R}^{k r} = W_N^{k r STEP}
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(tw + k * r * STEP));
    }
    fft_reg<R>(v);
    const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
    for (int r = 0; r < R; ++r) dst[pad_idx(j0 + r * NS)] = v[bitrev<R>(r)];
  }
}

template <int N, int NS>
struct NextRadix {
  static constexpr int rem = N / NS;
  static constexpr int value = rem >= 16 ? 16 : rem;
};

template <int N, int NS, int B, int THREADS>
struct Passes {
  __device__ static __forceinline__ float2* run(float2* a, float2* b, const float2* __restrict__ tw) {
    if constexpr (NS >= N) {
      return a;
    } else {
      constexpr int R = NextRadix<N, NS>::value;
      stockham_pass<N, R, NS, B, THREADS>(a, b, tw);
      __syncthreads();
      return Passes<N, NS * R, B, THREADS>::run(b, a, tw);
    }
  }
};

// Forward DFT of B sequences stored (padded) in `a`; `b` is scratch of the same size.  The
// caller must __syncthreads() after filling `a`.  Returns the buffer holding the result
// (natural order, padded); a __syncthreads() has been issued after the last pass.
template <int N, int B, int THREADS>
__device__ __forceinline__ float2* fft_forward(float2* a, float2* b, const float2* __restrict__ tw) {
  return Passes<N, 1, B, THREADS>::run(a, b, tw);
}

template <int N>
constexpr int num_passes() {
  int n = 0;
  for (int ns = 1; ns < N; ns *= ((N / ns) >= 16 ? 16 : (N / ns))) ++n;
  return n;
}

}  // namespace tmcfft
