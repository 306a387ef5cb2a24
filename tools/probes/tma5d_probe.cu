// probe: 5-D tiled TMA store with a box wider than the tensor dimension and with negative start coordinates (sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma5d_probe tma5d_probe.cu -lcuda ; ./tma5d_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int q0, int nh, int dx, int dy) {
  __shared__ __align__(1024) uint16_t buf[32 * 32];
  // row r, chunk g (8 elements) at chunk g ^ ((r >> 1) & 3): value = r * 100 + channel
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    int r = i / 32, c = i % 32, g = c / 8;
    __nv_bfloat16 v = __float2bfloat16((float)(r * 4 + (c & 3)));
    buf[r * 32 + ((g ^ ((r >> 1) & 3)) * 8) + (c % 8)] = *reinterpret_cast<uint16_t*>(&v);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(buf)), "r"(0), "r"(dx), "r"(q0), "r"(dy), "r"(nh) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int C = 32, Wo = 6, NH = 4, ld = 32;
  size_t n = (size_t)NH * 2 * Wo * 2 * ld;
  uint16_t* d; cudaMalloc(&d, n * 2); cudaMemset(d, 0xFF, n * 2);
  cuuint64_t ldb = ld * 2;
  cuuint64_t dims[5] = {C, 2, Wo, 2, NH};
  cuuint64_t strides[4] = {ldb, 2 * ldb, 2 * Wo * ldb, 4 * Wo * ldb};
  cuuint32_t box[5] = {32, 1, (cuuint32_t)(variant == 2 ? 4 : 32), 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUtensorMap tm;
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode rc=%d\n", variant, (int)r);
  int q0 = (variant == 1) ? -3 : 0;
  k<<<1, 128>>>(tm, q0, 1, 1, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    uint16_t* h = (uint16_t*)malloc(n * 2); cudaMemcpy(h, d, n * 2, cudaMemcpyDeviceToHost);
    int written = 0;
    for (size_t i = 0; i < n; ++i) if (h[i] != 0xFFFF) ++written;
    printf("elements written: %d (expect %d)\n", written, (variant == 1 ? 3 : (variant == 2 ? 4 : 6)) * 32);
    // print first element of each written pixel at (nh=1, dy=0, q, dx=1)
    for (int q = 0; q < Wo; ++q) {
      size_t off = ((((size_t)1 * 2 + 0) * Wo + q) * 2 + 1) * ld;
      uint32_t u = (uint32_t)h[off] << 16; float f; memcpy(&f, &u, 4);
      printf(" q=%d first=%g", q, h[off] == 0xFFFF ? -1.0f : f);
    }
    printf("\n");
  }
  return 0;
}
