// Issue-rate probe of the fp32 pipe on sm_100a (diagnostic, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_probe tools/fma_probe.cu && tools/fma_probe
// Every variant runs `iters` loop trips of NS scalar FFMA chains and NP packed FFMA2 chains per thread (all independent,
// operands in registers) with 16 warps per SM and reports SM cycles per warp-instruction per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <int NS, int NP, bool SCALAR_OPERAND>
__global__ void __launch_bounds__(512, 1) probe(float* out, int iters, long long* cycles) {
  float s[NS > 0 ? NS : 1];
  float2 p[NP > 0 ? NP : 1];
  const float a = 1.0001f + threadIdx.x * 1e-7f, b = 0.5f;
  const float2 a2 = make_float2(a, a + 1e-6f), b2 = make_float2(b, b + 1e-6f);
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = make_float2(threadIdx.x + i, i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < NS; ++i) s[i] = __fmaf_rn(s[i], a, b);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        if (SCALAR_OPERAND)
          p[i] = __ffma2_rn(make_float2(a, a), p[i], b2);  // scalar broadcast multiplier
        else
          p[i] = __ffma2_rn(p[i], a2, b2);
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) acc += s[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) acc += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// all three operands are fresh register pairs (no loop-invariant operand the reuse cache could hold): register-file
// bandwidth of the packed form.  MODE 0: d = a * b + d; 1: d = scalar * b + d; 2: scalar FFMA with 3 fresh operands
template <int NP, int MODE>
__global__ void __launch_bounds__(512, 1) probe_rf(float* out, int iters, long long* cycles) {
  float2 p[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = make_float2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f - 1e-3f * i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 x = p[(i + 3) % NP], y = p[(i + 5) % NP];
        if (MODE == 0) p[i] = __ffma2_rn(x, y, p[i]);
        if (MODE == 1) p[i] = __ffma2_rn(make_float2(x.x, x.x), y, p[i]);
        if (MODE == 2) {
          p[i].x = __fmaf_rn(x.x, y.x, p[i].x);
          p[i].y = __fmaf_rn(x.y, y.y, p[i].y);
        }
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) acc += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NP, int MODE>
void run_rf(const char* name, float* out, long long* cyc, int sms) {
  const int iters = 2000;
  probe_rf<NP, MODE><<<sms, 512>>>(out, iters, cyc);
  probe_rf<NP, MODE><<<sms, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
  const double lanes = 4.0 * iters * 8 * NP * 2.0 * 32;
  printf("%-44s %2d chains: %.1f fma lanes / cycle / scheduler\n", name, NP, lanes / mx);
}

template <int NS, int NP, bool SO>
void run(const char* name, float* out, long long* cyc, int sms) {
  const int iters = 2000;
  probe<NS, NP, SO><<<sms, 512>>>(out, iters, cyc);
  probe<NS, NP, SO><<<sms, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
  // per scheduler: 4 warps x iters x 8 x (NS + NP) warp-instructions
  const double inst = 4.0 * iters * 8 * (NS + NP);
  const double flops_lane = 4.0 * iters * 8 * (NS + 2.0 * NP) * 32;  // fma lane-ops per scheduler
  printf("%-34s %2d scalar + %2d packed chains: %.3f cycles / warp-instruction, %.1f fma lanes / cycle / scheduler\n", name, NS, NP,
         mx / inst, flops_lane / mx);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * 512 * sizeof(float));
  cudaMalloc(&cyc, 256 * sizeof(long long));
  run<8, 0, false>("scalar only", out, cyc, sms);
  run<16, 0, false>("scalar only", out, cyc, sms);
  run<0, 8, false>("packed only", out, cyc, sms);
  run<0, 16, false>("packed only", out, cyc, sms);
  run<0, 8, true>("packed, scalar operand", out, cyc, sms);
  run<4, 8, false>("mixed", out, cyc, sms);
  run<8, 8, false>("mixed", out, cyc, sms);
  run<8, 4, false>("mixed", out, cyc, sms);
  run<16, 8, false>("mixed", out, cyc, sms);
  run<8, 8, true>("mixed, scalar operand", out, cyc, sms);
  run_rf<12, 0>("packed, 3 fresh pair operands", out, cyc, sms);
  run_rf<12, 1>("packed, scalar + 2 fresh pair operands", out, cyc, sms);
  run_rf<12, 2>("scalar, 3 fresh operands", out, cyc, sms);
  run_rf<16, 0>("packed, 3 fresh pair operands", out, cyc, sms);
  run_rf<16, 2>("scalar, 3 fresh operands", out, cyc, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
