// hmma_rate_probe.cu -- issue rate of the legacy warp-level mma.sync.m16n8k16 (bf16 -> fp32) on sm_100a: cycles per HMMA
// per SM sub-partition with 1, 2, 4 warps per sub-partition, 4 independent accumulators per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probes/hmma_rate_probe tools/probes/hmma_rate_probe.cu
#include <cstdio>
#include <cstdint>
__global__ void k(float* out, long long* cyc, int iters) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
  float acc[4][4] = {};
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  float s = 0;
  for (int j = 0; j < 4; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 4 * 1024 * 148); cudaMalloc(&c, 8);
  const int iters = 4096;
  for (int warps : {1, 4, 8, 16}) {
    k<<<1, warps * 32>>>(o, c, iters);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = (double)warps / 4.0 < 1 ? 1 : warps / 4.0;   // warps per sub-partition
    printf("%2d warps in one CTA: %lld cycles for %d HMMA per warp -> %.2f cycles per HMMA per sub-partition\n", warps, h, iters * 4,
           (double)h / (iters * 4 * per_smsp));
  }
  return 0;
}
