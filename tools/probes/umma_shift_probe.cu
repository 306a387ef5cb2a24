// umma_shift_probe.cu -- hardware question behind a planned conv redesign ("next"):
// can a K-major SWIZZLE_128B UMMA operand start at an arbitrary ROW of a larger swizzled shared-memory patch?
// If yes, a 3x3 convolution can load each input patch once and feed all nine taps from it through shifted
// descriptors instead of nine im2col loads.
//
// The kernel fills a patch of 192 rows x 64 bf16 (128-byte rows, TMA's SWIZZLE_128B layout: 16-byte chunk c of row r
// lives at chunk c ^ (r & 7)) and a 64 x 64 B tile, then for each shift s issues MMAs (M=128, N=64, K=64) whose A
// descriptor starts at row s, once with base_offset = 0 and once with base_offset = (start >> 7) & 7, and compares
// D with the CPU product of rows s..s+127.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probes/umma_shift_probe tools/probes/umma_shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../tensorflow_yolo_b200/csrc/conv_tc.cuh"

using namespace yb;

constexpr int PATCH_ROWS = 192;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                       float* __restrict__ D, int shift, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // 192 x 128 B
  uint8_t* sB = smem + PATCH_ROWS * 128;                // 64 x 128 B (24576 is a multiple of 1024)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < PATCH_ROWS * 8; i += 128) {     // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<64>(&tmem_ptr);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1 && elect_one()) {
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)shift * 128u;
    uint64_t da = make_kmajor_desc<64>(a_addr);
    if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;       // matrix-descriptor base offset, bits [49,52)
    const uint64_t db = make_kmajor_desc<64>(smem_u32(sB));
    constexpr uint32_t idesc = make_idesc<64>();
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int chunk = 0; chunk < 2; ++chunk) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(chunk * 32), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)tid * 64 + chunk * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  std::vector<uint16_t> hA(PATCH_ROWS * 64), hB(64 * 64);
  srand(1);
  for (auto& x : hA) x = f2bf((float)(rand() % 17 - 8) / 8.0f);
  for (auto& x : hB) x = f2bf((float)(rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = PATCH_ROWS * 128 + 64 * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hD(128 * 64);
  int rc = 0;
  for (int ubo = 0; ubo < 2; ++ubo) {
    for (int shift : {0, 1, 2, 3, 5, 7, 8, 9, 15, 16, 27, 54, 55, 63}) {
      cudaMemset(dD, 0, hD.size() * 4);
      probe_kernel<<<1, 128, smem>>>(dA, dB, dD, shift, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", shift, ubo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0; double maxerr = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < 64; ++n) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += bf2f(hA[(shift + i) * 64 + k]) * bf2f(hB[n * 64 + k]);
          const double err = fabs((double)ref - hD[i * 64 + n]);
          maxerr = err > maxerr ? err : maxerr;
          if (err > 1e-3) ++bad;
        }
      printf("base_offset %s  shift %2d rows: %s (mismatches %d / 8192, max err %.3g)\n", ubo ? "set " : "zero", shift,
             bad ? "WRONG" : "ok", bad, maxerr);
      if (bad && ubo) rc = 1;
    }
  }
  return rc;
}
