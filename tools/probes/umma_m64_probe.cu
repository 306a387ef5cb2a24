// umma_m64_probe.cu -- swapped operand roles for the narrow (Cout = 64) fused convs: A = weights (M = 64 rows),
// B = pixels (N = 256).  Questions: (1) where does D[m][n] of an M = 64, cta_group::1 tcgen05.mma land in TMEM,
// (2) cycles per MMA (K = 16) for (M, N) = (64, 256), (64, 128), (128, 64), (128, 128), (128, 256), SS mode, SWIZZLE_64B
// K-major operands (64-byte rows), back-to-back issue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probes/umma_m64_probe tools/probes/umma_m64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../tensorflow_yolo_b200/csrc/conv_tc.cuh"

using namespace yb;

template <int M, int N>
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                       float* __restrict__ D, int iters, long long* cycles, int distinct) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // M x 64 B
  uint8_t* sB = smem + 4 * 128 * 64;                    // N x 64 B (x 4 tiles in the distinct-operand timing mode)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * 4; i += 128) {              // 16-byte chunks, SWIZZLE_64B: chunk ^= (row >> 1) & 3
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(sA + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 32 + c * 8);
  }
  for (int i = tid; i < N * 4; i += 128) {
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(sB + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 32 + c * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_ptr);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1 && elect_one()) {
    const uint64_t da = make_kmajor_desc<32>(smem_u32(sA));
    const uint64_t db = make_kmajor_desc<32>(smem_u32(sB));
    constexpr uint32_t idesc = make_idesc_m<N, M>();
    const long long t0 = clock64();
    if (distinct) {
      // successive MMAs read different operand tiles: A tile t at +t*M*64 bytes, B tile t at +t*N*64 bytes (garbage data: timing only)
      for (int it = 0; it < iters; ++it) {
        const int ta = it % distinct;
        const uint64_t da2 = make_kmajor_desc<32>(smem_u32(sA) + (uint32_t)(ta * M * 64));
        const uint64_t db2 = make_kmajor_desc<32>(smem_u32(sB) + (uint32_t)(ta * N * 64));
        for (int k = 0; k < 2; ++k) umma_bf16(tmem, da2 + 2 * k, db2 + 2 * k, idesc, (it | k) ? 1u : 0u);
      }
    } else
    for (int it = 0; it < iters; ++it)
      for (int k = 0; k < 2; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, (it | k) ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    if (cycles) *cycles = clock64() - t0;
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (D) {
    uint32_t v[32];
    for (int chunk = 0; chunk < N / 32; ++chunk) {
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(chunk * 32), v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(size_t)tid * N + chunk * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

template <int M, int N>
void run(const std::vector<uint16_t>& hA, const std::vector<uint16_t>& hB, __nv_bfloat16* dA, __nv_bfloat16* dB, float* dD, long long* dC) {
  const int smem = 4 * 128 * 64 + 4 * 256 * 64 + 1024;
  cudaFuncSetAttribute(probe_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(dD, 0, 128 * 256 * 4);
  probe_kernel<M, N><<<1, 128, smem>>>(dA, dB, dD, 1, nullptr, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e)); exit(1); }
  std::vector<float> hD(128 * N);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  // hypotheses for the lane of row m
  int bad_id = 0, bad_h16 = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < 32; ++k) ref += bf2f(hA[m * 32 + k]) * bf2f(hB[n * 32 + k]);
      const int lane_id = m, lane_h16 = (m / 16) * 32 + (m % 16);
      if (fabsf(hD[(size_t)lane_id * N + n] - ref) > 1e-3f) ++bad_id;
      if (fabsf(hD[(size_t)lane_h16 * N + n] - ref) > 1e-3f) ++bad_h16;
    }
  printf("M=%3d N=%3d  layout: lane=m mismatches %d, lane=(m/16)*32+m%%16 mismatches %d", M, N, bad_id, bad_h16);
  if (M == 64 && bad_id && bad_h16) {          // find row 0..63 empirically: which lane holds row m (column 0..7 signature)
    printf("\n  lanes of rows:");
    for (int m = 0; m < M; ++m) {
      float ref[8];
      for (int n = 0; n < 8; ++n) { ref[n] = 0; for (int k = 0; k < 32; ++k) ref[n] += bf2f(hA[m * 32 + k]) * bf2f(hB[n * 32 + k]); }
      int found = -1;
      for (int l = 0; l < 128 && found < 0; ++l) {
        bool ok = true;
        for (int n = 0; n < 8; ++n) ok = ok && fabsf(hD[(size_t)l * N + n] - ref[n]) < 1e-3f;
        if (ok) found = l;
      }
      printf(" %d", found);
    }
  }
  for (int distinct : {0, 4}) {
    probe_kernel<M, N><<<1, 128, smem>>>(dA, dB, nullptr, 512, dC, distinct);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
    printf("  | %s operands: %.1f clk/MMA", distinct ? "4 distinct" : "same", (double)c / 1024);
  }
  printf("\n");
}

int main() {
  std::vector<uint16_t> hA(128 * 32), hB(256 * 32);
  srand(1);
  for (auto& x : hA) x = f2bf((float)(rand() % 17 - 8) / 8.0f);
  for (auto& x : hB) x = f2bf((float)(rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  run<64, 256>(hA, hB, dA, dB, dD, dC);
  run<64, 128>(hA, hB, dA, dB, dD, dC);
  run<128, 64>(hA, hB, dA, dB, dD, dC);
  run<128, 128>(hA, hB, dA, dB, dD, dC);
  run<128, 256>(hA, hB, dA, dB, dD, dC);
  return 0;
}
