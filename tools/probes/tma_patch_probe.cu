// probe: 3-D tiled TMA LOAD of a raw image patch (uint8 and float32 elements) with negative / odd start coordinates and
// rows past the image (sm_100a).  Checks zero fill and byte-exact placement: what the fused stem kernel relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_patch_probe tma_patch_probe.cu -lcuda ; ./tma_patch_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <typename T, int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, T* out) {
  __shared__ __align__(128) T buf[BW * BH];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(BW * BH * sizeof(T))) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = buf[i];
}
template <typename T, CUtensorMapDataType DT>
int run(const char* name, int only) {
  const int N = 2, H = 8, WB = 48;   // WB elements per row (W*3)
  const int BW = 32, BH = 6;
  T* h = (T*)malloc(sizeof(T) * N * H * WB);
  for (int i = 0; i < N * H * WB; ++i) h[i] = (T)(1 + i % 200);
  T *d, *o; cudaMalloc(&d, sizeof(T) * N * H * WB); cudaMalloc(&o, sizeof(T) * BW * BH);
  cudaMemcpy(d, h, sizeof(T) * N * H * WB, cudaMemcpyHostToDevice);
  cuuint64_t dims[3] = {WB, H, N};
  cuuint64_t strides[2] = {WB * sizeof(T), (cuuint64_t)WB * H * sizeof(T)};
  cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
  CUtensorMap tm;
  CUresult r = cuTensorMapEncodeTiled(&tm, DT, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("%s encode rc=%d\n", name, (int)r);
  int bad_total = 0;
  const int starts[7][3] = {{3, 0, 1}, {21, 5, 0}, {-4, 0, 1}, {0, -2, 1}, {-4, -2, 1}, {-5, -2, 1}, {17, 3, 0}};
  for (int v = 0; v < 7; ++v) {
    if (only >= 0 && v != only) continue;
    cudaMemset(o, 0xEE, sizeof(T) * BW * BH);
    k<T, BW, BH><<<1, 64>>>(tm, starts[v][0], starts[v][1], starts[v][2], o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  start (%d,%d,%d): kernel error %s\n", starts[v][0], starts[v][1], starts[v][2], cudaGetErrorString(e)); return 1; }
    T got[BW * BH]; cudaMemcpy(got, o, sizeof(got), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
      const int gx = starts[v][0] + x, gy = starts[v][1] + y, n = starts[v][2];
      const T want = (gx < 0 || gx >= WB || gy < 0 || gy >= H) ? (T)0 : h[(n * H + gy) * WB + gx];
      if (memcmp(&want, &got[y * BW + x], sizeof(T)) != 0) ++bad;
    }
    printf("  start (%d,%d,%d): %d mismatches of %d\n", starts[v][0], starts[v][1], starts[v][2], bad, BW * BH);
    bad_total += bad;
  }
  return bad_total;
}
int main(int argc, char** argv) {
  cuInit(0);
  cudaFree(0);
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  int b = 0;
  if (argc < 2 || argv[1][0] == 'u') b += run<uint8_t, CU_TENSOR_MAP_DATA_TYPE_UINT8>("uint8", only);
  if (argc < 2 || argv[1][0] == 'f') b += run<float, CU_TENSOR_MAP_DATA_TYPE_FLOAT32>("float32", only);
  printf(b == 0 ? "ALL OK\n" : "MISMATCHES\n");
  return 0;
}
