// Stand-alone probe of one 3-D tiled TMA box load (debugging aid for csrc/warp_tma.cuh):
//   tma_probe W H T boxW boxH boxZ cx cy cz
// prints the sum of the box as read through TMA and as computed on the host (zero fill outside the tensor).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int cx, int cy, int cz, int bytes, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  unsigned char* dst = smem + ((128u - (smem_u32(smem) & 127u)) & 127u);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(&map), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(&bar))
        : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(
          smem_u32(&bar)),
      "r"(0)
      : "memory");
  const float* f = reinterpret_cast<const float*>(dst);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = f[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  if (argc < 10) return 1;
  const int W = atoi(argv[1]), H = atoi(argv[2]), T = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]), bz = atoi(argv[6]),
            cx = atoi(argv[7]), cy = atoi(argv[8]), cz = atoi(argv[9]);
  std::vector<float> host((size_t)W * H * T);
  for (size_t i = 0; i < host.size(); ++i) host[i] = (float)(i % 1000) * 0.001f + 1.0f;
  float *dev, *out;
  cudaMalloc(&dev, host.size() * 4);
  cudaMemcpy(dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
  const int bytes = bw * bh * bz * 4;
  cudaMalloc(&out, bytes);
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult st;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &st);
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bz};
  const cuuint32_t elem[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dev, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d; ", (int)r);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 256);
  probe<<<1, 128, bytes + 256>>>(map, cx, cy, cz, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("tensor %dx%dx%d box %dx%dx%d at (%d,%d,%d): %s; ", W, H, T, bw, bh, bz, cx, cy, cz, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> got(bytes / 4);
    cudaMemcpy(got.data(), out, bytes, cudaMemcpyDeviceToHost);
    double sg = 0, sw = 0;
    for (int z = 0; z < bz; ++z)
      for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x) {
          sg += got[(z * bh + y) * bw + x];
          const int gx = cx + x, gy = cy + y, gz = cz + z;
          if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < T) sw += host[((size_t)gz * H + gy) * W + gx];
        }
    printf("sum %.4f want %.4f", sg, sw);
  }
  printf("\n");
  return 0;
}
