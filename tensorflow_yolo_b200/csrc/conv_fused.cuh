// conv_fused.cuh -- two-layer fused convolutions for the HBM-bound head of Darknet-53 (sm_100a only).
//
// The first layers of the reference's YOLOv3 graph (net/v3.py:25-27 -> net/layers.py:17-67) move far more bytes than
// they compute on: conv 3->32 @416^2 writes 1.4 GB per 128 images only for the stride-2 conv behind it to read them
// back nine times through im2col TMA loads.  Here ONE kernel computes both layers of such a pair:
//
//   producer conv (small K, on mma.sync)  ->  bf16 activations written straight into the K-major, 64-byte-swizzled
//   im2col operand tiles of the 3x3 consumer conv  ->  tcgen05.mma (M=128, N=64, K=9*32) with the accumulator in TMEM
//   ->  BN / leaky (/ residual) epilogue  ->  TMA store.
//
// The intermediate activation never exists in global memory.  An output tile is 8 rows x 16 columns of the consumer
// conv (128 GEMM rows, row m = r*16 + c), so the producer recomputes only the one-pixel halo between tiles
// (17x33 producer pixels per 8x16 stride-2 outputs: x1.10; 10x18 per 8x16 stride-1 outputs: x1.41).
//
//   STEM  (net/v3.py:25-26): conv 3x3 s1 3->32 + BN + leaky  ->  conv 3x3 s2 32->64 + BN + leaky
//   BLOCK (net/v3.py:16-19,27): conv 1x1 64->32 + BN + leaky  ->  conv 3x3 s1 32->64 + BN + leaky  ->  + block input
//
// Arithmetic is the same as the unfused kernels': the producer's mma.sync fragments, k order and bf16 rounding are
// those of conv_first_mma_kernel / the tcgen05 1x1 conv up to fp32 summation order, the consumer's K walk (tap-major,
// 32 channels per tap, two K=16 MMAs per tap) is that of conv_tc_persist_kernel<64,32,false>.
//
// Schedule (one persistent CTA per SM, 12 warps in lock step, all of them producers):
//   iteration i:  registers -> smem input patch (prefetched during iteration i-1)   | __syncthreads
//                 producer conv of tile i on mma.sync, scatter into A[i&1]           | fence.proxy.async, __syncthreads
//                 warp 8: 18 x tcgen05.mma into TMEM accumulator i&1, commit         |
//                 warps 0-7: epilogue of tile i-1 (overlaps the MMAs of tile i)      |
#pragma once
#include "aux_kernels.cuh"
#include "conv_tc.cuh"

namespace yb {

constexpr int FUSE_THREADS = 384;
constexpr int FUSE_WARPS = FUSE_THREADS / 32;
constexpr int FUSE_TH = 8, FUSE_TW = 16;                  // output tile of the consumer conv
constexpr int FUSE_CMID = 32, FUSE_COUT = 64;
constexpr int FUSE_TAP_BYTES = 128 * FUSE_CMID * 2;       // one tap of the A operand: 128 rows x 64 bytes
constexpr int FUSE_A_BYTES = 9 * FUSE_TAP_BYTES;          // 73,728
constexpr int FUSE_BTAP_BYTES = FUSE_COUT * FUSE_CMID * 2;
constexpr int FUSE_B_BYTES = 9 * FUSE_BTAP_BYTES;         // 36,864
constexpr int FUSE_EPI_BYTES = 8 * 2048;                  // 8 epilogue warps x (32 rows x 64 B)
// STEM input patch: 19 rows x 112 bf16 (pixels 2*q0-4 .. 2*q0+33, 3 channels, 16-byte aligned row starts in global)
constexpr int STEM_PATCH_ROWS = 19, STEM_PATCH_PITCH = 112, STEM_PATCH_WORDS = STEM_PATCH_ROWS * 28;
constexpr int STEM_PATCH_BYTES = ((STEM_PATCH_ROWS * STEM_PATCH_PITCH * 2 + 16) + 127) / 128 * 128;
// BLOCK input patch: 10 x 18 pixels x 64 channels bf16, pixel pitch 144 bytes (conflict-free ldmatrix rows)
constexpr int BLOCK_PATCH_PIX = 10 * 18, BLOCK_PATCH_PITCH = 144;
constexpr int BLOCK_PATCH_BYTES = BLOCK_PATCH_PIX * BLOCK_PATCH_PITCH;     // 25,920
constexpr int FUSE_HEADER = 1024;                         // barriers, TMEM pointer, consumer scale/shift
constexpr int FUSE_SMEM_STEM = 1024 + FUSE_HEADER + FUSE_B_BYTES + 2 * FUSE_A_BYTES + FUSE_EPI_BYTES + STEM_PATCH_BYTES;
constexpr int FUSE_SMEM_BLOCK = 1024 + FUSE_HEADER + FUSE_B_BYTES + 2 * FUSE_A_BYTES + FUSE_EPI_BYTES + BLOCK_PATCH_BYTES;

struct FuseArgs {
  int n_img, Ho, Wo;             // consumer output size; Ho % 8 == 0, Wo % 16 == 0
  int H, W;                      // producer input size (STEM: 2*Ho x 2*Wo image; BLOCK: Ho x Wo)
  const void* in;                // STEM: [N,H,W,3] float32 or uint8; BLOCK: [N,H,W,in_ld] bf16 (64 channels used)
  int in_ld;
  const void* w1;                // STEM: fp32 [27][32] (tap*3+ci major); BLOCK: bf16 [32][64] (cout major, K contiguous)
  const float *scale1, *shift1;  // producer BN (32 channels)
  const float *scale2, *shift2;  // consumer BN (64 channels)
  int leaky1, leaky2;
  int tiles_h, tiles_w, n_tiles; // Ho/8, Wo/16, n_img * tiles_h * tiles_w
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// Scatter of one producer pixel (8 channels = 16-byte chunk `t` of its 64-byte row) into the im2col tiles of a
// stride-S 3x3 consumer: the pixel at tile-relative consumer-input position (yy, xx) is tap (kh, kw) of output
// (r, c) whenever S*r + kh == yy and S*c + kw == xx.  Row m = r*16 + c of tap kh*3+kw, 64-byte rows, SWIZZLE_64B:
// 16-byte chunk index ^= (m >> 1) & 3 -- the layout TMA writes and the UMMA descriptor (make_kmajor_desc<32>) reads.
template <int S>
__device__ __forceinline__ void scatter_pixel(uint32_t a_base, int yy, int xx, int t, const uint32_t (&pk)[4]) {
  if (S == 2) {
    // stride 2: the tap parity equals the coordinate parity, so at most two kh and two kw qualify
    const int kh0 = yy & 1, kw0 = xx & 1;
    const int r0 = (yy - kh0) >> 1, c0 = (xx - kw0) >> 1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kh = kh0 + 2 * i, r = r0 - i;
      if (kh > 2 || r < 0 || r >= FUSE_TH) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kw = kw0 + 2 * j, c = c0 - j;
        if (kw > 2 || c < 0 || c >= FUSE_TW) continue;
        const int m = r * FUSE_TW + c;
        sts128(a_base + (uint32_t)((kh * 3 + kw) * FUSE_TAP_BYTES + m * 64 + ((t ^ ((m >> 1) & 3)) << 4)), pk[0], pk[1], pk[2], pk[3]);
      }
    }
  } else {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int r = yy - kh;
      if (r < 0 || r >= FUSE_TH) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int c = xx - kw;
        if (c < 0 || c >= FUSE_TW) continue;
        const int m = r * FUSE_TW + c;
        sts128(a_base + (uint32_t)((kh * 3 + kw) * FUSE_TAP_BYTES + m * 64 + ((t ^ ((m >> 1) & 3)) << 4)), pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

// Epilogue of one tile by warps 0-7: TMEM lane quarter = warp & 3 (tile rows 2q, 2q+1), 32-column chunk = warp >> 2.
// BN + leaky (+ residual from global memory) in registers, bf16 through 64-byte-swizzled staging, one 4-D TMA store
// of a [2 rows][16 cols][32 channels] box.
__device__ __forceinline__ void fused_epilogue(const CUtensorMap* tmOut, uint64_t* mma_done, uint32_t tmem_base, int acc, uint32_t parity,
                                               uint8_t* stage_buf, const float* s_scale2, const float* s_shift2, int leaky,
                                               int img, int p0, int q0, const __nv_bfloat16* res, int res_ld, int Ho, int Wo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, chunk = warp >> 2;
  uint4 rv[4];
  if (res != nullptr) {           // requested before the accumulator wait: the latency hides behind the MMAs
    const int p = p0 + 2 * quarter + (lane >> 4), q = q0 + (lane & 15);
    const __nv_bfloat16* rp = res + (((long long)img * Ho + p) * Wo + q) * res_ld + chunk * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) rv[g] = __ldg(reinterpret_cast<const uint4*>(rp + g * 8));
  }
  mbar_wait(&mma_done[acc], parity);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * FUSE_COUT + chunk * 32), v);
  if (lane == 0) tma_store_wait_read<0>();          // the previous tile's store has read this staging buffer
  __syncwarp();
  tmem_ld_wait();
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float y = __uint_as_float(v[j]) * s_scale2[chunk * 32 + j] + s_shift2[chunk * 32 + j];
    if (leaky) y = fmaxf(y, 0.1f * y);
    f[j] = y;
  }
  if (res != nullptr) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t rw[4] = {rv[g].x, rv[g].y, rv[g].z, rv[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[g * 8 + 2 * j] += __uint_as_float(rw[j] << 16);
        f[g * 8 + 2 * j + 1] += __uint_as_float(rw[j] & 0xFFFF0000u);
      }
    }
  }
  const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
  const uint32_t row = smem_u32(stage_buf) + (uint32_t)lane * 64u;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    sts128(row + (((uint32_t)g ^ sw) << 4), pack_bf16(f[g * 8 + 0], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
           pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(tmOut, stage_buf, chunk * 32, q0, p0 + 2 * quarter, img);
    tma_store_commit();
  }
  tc_fence_before();
}

// Issues the consumer conv of one tile: 9 taps x 2 (K = 16) tcgen05.mma, M = 128, N = 64, then commits to `done`.
__device__ __forceinline__ void fused_issue_mma(uint32_t a_addr, uint32_t b_addr, uint32_t tmem_d, uint64_t* done) {
  constexpr uint32_t idesc = make_idesc<FUSE_COUT>();
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint64_t da = make_kmajor_desc<FUSE_CMID>(a_addr + (uint32_t)(tap * FUSE_TAP_BYTES));
    const uint64_t db = make_kmajor_desc<FUSE_CMID>(b_addr + (uint32_t)(tap * FUSE_BTAP_BYTES));
#pragma unroll
    for (int k = 0; k < FUSE_CMID / 16; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
  }
  umma_commit(done);
}

struct FuseSmem {
  uint64_t* mma_done;     // [2]
  uint64_t* b_full;       // [1]
  uint32_t* tmem_ptr;
  float *s_scale2, *s_shift2;
  uint8_t *bs, *a0, *epi, *patch;
};
__device__ __forceinline__ FuseSmem fuse_carve(uint8_t* smem) {
  FuseSmem s;
  s.mma_done = reinterpret_cast<uint64_t*>(smem);
  s.b_full = s.mma_done + 2;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(smem + 64);
  s.s_scale2 = reinterpret_cast<float*>(smem + 256);
  s.s_shift2 = s.s_scale2 + FUSE_COUT;
  s.bs = smem + FUSE_HEADER;
  s.a0 = s.bs + FUSE_B_BYTES;
  s.epi = s.a0 + 2 * FUSE_A_BYTES;
  s.patch = s.epi + FUSE_EPI_BYTES;
  return s;
}

// Common prologue: barriers, TMEM (128 columns = two 64-column accumulators), the consumer's weights (9 taps x
// [64 x 32] through the conv's own weight map), its scale/shift.
__device__ __forceinline__ uint32_t fuse_prologue(const FuseSmem& s, const CUtensorMap* tmB, const CUtensorMap* tmOut, const FuseArgs& a) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(tmB);
    tma_prefetch_desc(tmOut);
    mbar_init(&s.mma_done[0], 1);
    mbar_init(&s.mma_done[1], 1);
    mbar_init(s.b_full, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc<2 * FUSE_COUT>(s.tmem_ptr);
  if (threadIdx.x == 32) {
    mbar_expect_tx(s.b_full, (uint32_t)FUSE_B_BYTES);
    for (int tap = 0; tap < 9; ++tap) tma_load_2d(tmB, s.b_full, s.bs + tap * FUSE_BTAP_BYTES, tap * FUSE_CMID, 0);
  }
  for (int i = threadIdx.x; i < FUSE_COUT; i += FUSE_THREADS) {
    s.s_scale2[i] = a.scale2[i];
    s.s_shift2[i] = a.shift2[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *s.tmem_ptr;
}

// ================================================================================================
// STEM: conv 3x3 s1 3->32 (mma.sync m16n8k16, K = 27 padded to 32) feeding conv 3x3 s2 32->64 (tcgen05)
// ================================================================================================
// Producer pixels of a tile, relative to (2*p0 - 1, 2*q0 - 1): yy in [0,17), xx in [0,33).  They are walked by parity
// class so that the 16 pixels of an mma tile land in consecutive rows of one im2col tap (conflict-free stores):
//   mt  0.. 7  yy odd,  xx odd   (row r = mt, 16 columns)             -> centre tap only
//   mt  8..16  yy even, xx odd   (9 rows x 16)                        -> kh in {0,2}
//   mt 17..25  yy odd,  xx even  (8 rows x 17, flattened, 136 pixels) -> kw in {0,2}
//   mt 26..35  yy even, xx even  (9 rows x 17, flattened, 153 pixels) -> four taps
template <bool U8>
__global__ void __launch_bounds__(FUSE_THREADS, 1)
stem_fused_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut, const FuseArgs a) {
  extern __shared__ __align__(1024) uint8_t fuse_smem_raw[];
  uint8_t* smem = fuse_smem_raw + ((1024u - (smem_u32(fuse_smem_raw) & 1023u)) & 1023u);
  const FuseSmem s = fuse_carve(smem);
  uint16_t* s_patch = reinterpret_cast<uint16_t*>(s.patch);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  griddep_launch_dependents();
  const uint32_t tmem_base = fuse_prologue(s, &tmB, &tmOut, a);

  // ---- producer constants: B fragments of the first conv (permuted columns: column g of n-tile nt holds channel
  // (g/2)*8 + 2*nt + (g&1), so a thread's accumulators are the 8 consecutive channels 8t..8t+7), scale/shift ----
  constexpr int K1 = 27;
  const float* w1 = reinterpret_cast<const float*>(a.w1);
  uint32_t bf[2][4][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = ks * 16 + h * 8 + 2 * t;
        const int n = (g >> 1) * 8 + 2 * nt + (g & 1);
        const float w0 = (k < K1) ? w1[k * FUSE_CMID + n] : 0.0f;
        const float w1v = (k + 1 < K1) ? w1[(k + 1) * FUSE_CMID + n] : 0.0f;
        bf[ks][nt][h] = pack_bf16(w0, w1v);
      }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = a.scale1[8 * t + j]; sh[j] = a.shift1[8 * t + j]; }
  // patch element offsets of this thread's k columns, relative to (yy*PITCH + xx*3): k = tap*3 + ci, tap = dy*3 + dx,
  // input pixel (yy + dy, xx + 2 + dx) of the patch (patch column 0 is image column 2*q0 - 4)
  int soff[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = (j >> 2) * 16 + ((j >> 1) & 1) * 8 + 2 * t + (j & 1);
    const int kk = k < K1 ? k : 0;                  // padded columns read a valid element; their weights are zero
    const int tap = kk / 3, ci = kk - tap * 3;
    soff[j] = (tap / 3) * STEM_PATCH_PITCH + 6 + (tap % 3) * 3 + ci;
  }

  const int tiles_per_img = a.tiles_h * a.tiles_w;
  const int row_bytes = a.W * 3;                    // elements per image row
  // ---- input patch prefetch: 19 rows x 28 words; a word is 4 uint8 or 4 floats (16-byte aligned either way) ----
  uint4 pre[2];
  auto load_patch = [&](int tile) {
    const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
    const int p0 = (rem / a.tiles_w) * FUSE_TH, q0 = (rem % a.tiles_w) * FUSE_TW;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = threadIdx.x + i * FUSE_THREADS;
      pre[i] = make_uint4(0u, 0u, 0u, 0u);
      if (e < STEM_PATCH_WORDS) {
        const int rr = e / 28, wq = e - rr * 28;
        const int y = 2 * p0 - 2 + rr;
        const int bx = (2 * q0 - 4) * 3 + 4 * wq;
        if ((unsigned)y < (unsigned)a.H && (unsigned)bx < (unsigned)row_bytes) {
          const long long off = ((long long)img * a.H + y) * row_bytes + bx;
          if (U8) pre[i].x = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(a.in) + off));
          else pre[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.in) + off));
        }
      }
    }
  };
  auto store_patch = [&]() {
    constexpr float k255 = 0.003921568859368562698f;     // same conversion as conv_first_mma_kernel (bit-identical bf16)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = threadIdx.x + i * FUSE_THREADS;
      if (e < STEM_PATCH_WORDS) {
        uint2 o;
        if (U8) {
          const uint32_t w4 = pre[i].x;
          o = make_uint2(pack_bf16((float)(w4 & 0xFFu) * k255, (float)((w4 >> 8) & 0xFFu) * k255),
                         pack_bf16((float)((w4 >> 16) & 0xFFu) * k255, (float)(w4 >> 24) * k255));
        } else {
          o = make_uint2(pack_bf16(__uint_as_float(pre[i].x), __uint_as_float(pre[i].y)),
                         pack_bf16(__uint_as_float(pre[i].z), __uint_as_float(pre[i].w)));
        }
        *reinterpret_cast<uint2*>(s_patch + 4 * e) = o;            // row rr, word wq: element rr*112 + 4*wq = 4*e
      }
    }
  };

  int it = 0;
  int prev_img = 0, prev_p0 = 0, prev_q0 = 0;
  if ((int)blockIdx.x < a.n_tiles) load_patch(blockIdx.x);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
    const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
    const int p0 = (rem / a.tiles_w) * FUSE_TH, q0 = (rem % a.tiles_w) * FUSE_TW;
    const int buf = it & 1;
    store_patch();
    if (tile + (int)gridDim.x < a.n_tiles) load_patch(tile + gridDim.x);      // in flight during the producer conv
    __syncthreads();
    // ---- producer conv: three mma tiles of 16 pixels per warp ----
    const uint32_t a_base = smem_u32(s.a0) + (uint32_t)(buf * FUSE_A_BYTES);
#pragma unroll 1
    for (int k3 = 0; k3 < 3; ++k3) {
      const int mt = warp + FUSE_WARPS * k3;
      int yy[2], xx[2];
      bool ok[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int gi = g + 8 * r;
        if (mt < 8) { yy[r] = 2 * mt + 1; xx[r] = 2 * gi + 1; ok[r] = true; }
        else if (mt < 17) { yy[r] = 2 * (mt - 8); xx[r] = 2 * gi + 1; ok[r] = true; }
        else if (mt < 26) {
          int f = (mt - 17) * 16 + gi;
          ok[r] = f < 136; f = ok[r] ? f : 135;
          const int rr = f / 17;
          yy[r] = 2 * rr + 1; xx[r] = 2 * (f - rr * 17);
        } else {
          int f = (mt - 26) * 16 + gi;
          ok[r] = f < 153; f = ok[r] ? f : 152;
          const int rr = f / 17;
          yy[r] = 2 * rr; xx[r] = 2 * (f - rr * 17);
        }
      }
      uint32_t afrag[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint16_t* p = s_patch + yy[r] * STEM_PATCH_PITCH + xx[r] * 3;
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = p[soff[j]];
        afrag[0][r] = v[0] | (v[1] << 16);
        afrag[0][r + 2] = v[2] | (v[3] << 16);
        afrag[1][r] = v[4] | (v[5] << 16);
        afrag[1][r + 2] = v[6] | (v[7] << 16);
      }
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f; }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], afrag[ks], bf[ks][nt][0], bf[ks][nt][1]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        uint32_t pk[4];
        // the consumer's zero padding: producer pixels outside the image are zeros, not conv(0)
        const bool inside = (2 * p0 - 1 + yy[r] >= 0) && (2 * q0 - 1 + xx[r] >= 0);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float y0v = acc[nt][2 * r] * sc[2 * nt] + sh[2 * nt];
          float y1v = acc[nt][2 * r + 1] * sc[2 * nt + 1] + sh[2 * nt + 1];
          if (a.leaky1) { y0v = fmaxf(y0v, 0.1f * y0v); y1v = fmaxf(y1v, 0.1f * y1v); }
          pk[nt] = inside ? pack_bf16(y0v, y1v) : 0u;
        }
        if (ok[r]) scatter_pixel<2>(a_base, yy[r], xx[r], t, pk);
      }
    }
    fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
      if (elect_one()) {
        tc_fence_after();
        if (it == 0) mbar_wait(s.b_full, 0);
        fused_issue_mma(a_base, smem_u32(s.bs), tmem_base + (uint32_t)(buf * FUSE_COUT), &s.mma_done[buf]);
      }
      __syncwarp();
    } else if (warp < 8 && it > 0) {
      fused_epilogue(&tmOut, s.mma_done, tmem_base, buf ^ 1, (uint32_t)((it - 1) >> 1) & 1u, s.epi + warp * 2048, s.s_scale2,
                     s.s_shift2, a.leaky2, prev_img, prev_p0, prev_q0, nullptr, 0, a.Ho, a.Wo);
    }
    prev_img = img; prev_p0 = p0; prev_q0 = q0;
  }
  if (warp < 8 && it > 0) {
    fused_epilogue(&tmOut, s.mma_done, tmem_base, (it - 1) & 1, (uint32_t)((it - 1) >> 1) & 1u, s.epi + warp * 2048, s.s_scale2,
                   s.s_shift2, a.leaky2, prev_img, prev_p0, prev_q0, nullptr, 0, a.Ho, a.Wo);
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<2 * FUSE_COUT>(tmem_base);
  }
}

// ================================================================================================
// BLOCK: conv 1x1 64->32 (mma.sync m16n8k16, A fragments by ldmatrix) feeding conv 3x3 s1 32->64 (tcgen05) + residual
// ================================================================================================
// The residual block of net/v3.py:16-19 at 64 channels: X -> conv1x1(32) -> conv3x3(64) -> + X.  The producer needs the
// 10 x 18 input pixels around an 8 x 16 output tile; each 16-pixel mma tile is 16 consecutive pixels of that patch
// (180 pixels = 12 tiles, one per warp; the last holds 4 pixels).  Producer pixels outside the image are the 3x3
// conv's zero padding.  The residual is re-read from global memory by the epilogue (the patch was just fetched: L2 hits).
__global__ void __launch_bounds__(FUSE_THREADS, 1)
block_fused_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut, const FuseArgs a) {
  extern __shared__ __align__(1024) uint8_t fuse_smem_raw[];
  uint8_t* smem = fuse_smem_raw + ((1024u - (smem_u32(fuse_smem_raw) & 1023u)) & 1023u);
  const FuseSmem s = fuse_carve(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  griddep_launch_dependents();
  const uint32_t tmem_base = fuse_prologue(s, &tmB, &tmOut, a);

  // B fragments of the 1x1 conv: w1 is [32 cout][64 cin] bf16; b0 = (k = 2t, 2t+1; n), b1 = (k = 2t+8, 2t+9; n) with the
  // permuted column n = (g/2)*8 + 2*nt + (g&1), so a thread's accumulators are channels 8t..8t+7
  const uint32_t* w1 = reinterpret_cast<const uint32_t*>(a.w1);      // pairs of bf16 along K
  uint32_t bf[4][4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int n = (g >> 1) * 8 + 2 * nt + (g & 1);
      bf[ks][nt][0] = __ldg(w1 + (n * 64 + ks * 16 + 2 * t) / 2);
      bf[ks][nt][1] = __ldg(w1 + (n * 64 + ks * 16 + 8 + 2 * t) / 2);
    }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = a.scale1[8 * t + j]; sh[j] = a.shift1[8 * t + j]; }

  const int tiles_per_img = a.tiles_h * a.tiles_w;
  const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(a.in);
  // ---- input patch prefetch: 180 pixels x 8 chunks of 16 bytes, four chunks per thread ----
  uint4 pre[4];
  auto load_patch = [&](int tile) {
    const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
    const int p0 = (rem / a.tiles_w) * FUSE_TH, q0 = (rem % a.tiles_w) * FUSE_TW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * FUSE_THREADS;
      pre[i] = make_uint4(0u, 0u, 0u, 0u);
      if (e < BLOCK_PATCH_PIX * 8) {
        const int px = e >> 3, ck = e & 7;
        const int py = px / 18, pxx = px - py * 18;
        const int y = p0 - 1 + py, x = q0 - 1 + pxx;
        if ((unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W)
          pre[i] = __ldg(reinterpret_cast<const uint4*>(X + (((long long)img * a.H + y) * a.W + x) * a.in_ld + ck * 8));
      }
    }
  };
  auto store_patch = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * FUSE_THREADS;
      if (e < BLOCK_PATCH_PIX * 8)
        sts128(smem_u32(s.patch) + (uint32_t)((e >> 3) * BLOCK_PATCH_PITCH + (e & 7) * 16), pre[i].x, pre[i].y, pre[i].z, pre[i].w);
    }
  };
  // ldmatrix.x4 row address of this lane: matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
  const int lm_row = (lane & 7) + 8 * ((lane >> 3) & 1);
  const int lm_kofs = 8 * (lane >> 4);

  int it = 0;
  int prev_img = 0, prev_p0 = 0, prev_q0 = 0;
  if ((int)blockIdx.x < a.n_tiles) load_patch(blockIdx.x);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
    const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img;
    const int p0 = (rem / a.tiles_w) * FUSE_TH, q0 = (rem % a.tiles_w) * FUSE_TW;
    const int buf = it & 1;
    store_patch();
    if (tile + (int)gridDim.x < a.n_tiles) load_patch(tile + gridDim.x);
    __syncthreads();
    const uint32_t a_base = smem_u32(s.a0) + (uint32_t)(buf * FUSE_A_BYTES);
    {
      const int f0 = warp * 16;                            // 12 warps x 16 pixels cover the 180-pixel patch
      const int fr = min(f0 + lm_row, BLOCK_PATCH_PIX - 1);
      const uint32_t lm_addr = smem_u32(s.patch) + (uint32_t)(fr * BLOCK_PATCH_PITCH + lm_kofs * 2);
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t af[4];
        ldmatrix_x4(lm_addr + (uint32_t)(ks * 32), af);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], af, bf[ks][nt][0], bf[ks][nt][1]);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int f = f0 + g + 8 * r;
        const int py = f / 18, pxx = f - py * 18;
        const int y = p0 - 1 + py, x = q0 - 1 + pxx;
        const bool inside = (unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W;
        uint32_t pk[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float y0v = acc[nt][2 * r] * sc[2 * nt] + sh[2 * nt];
          float y1v = acc[nt][2 * r + 1] * sc[2 * nt + 1] + sh[2 * nt + 1];
          if (a.leaky1) { y0v = fmaxf(y0v, 0.1f * y0v); y1v = fmaxf(y1v, 0.1f * y1v); }
          pk[nt] = inside ? pack_bf16(y0v, y1v) : 0u;
        }
        if (f < BLOCK_PATCH_PIX) scatter_pixel<1>(a_base, py, pxx, t, pk);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
      if (elect_one()) {
        tc_fence_after();
        if (it == 0) mbar_wait(s.b_full, 0);
        fused_issue_mma(a_base, smem_u32(s.bs), tmem_base + (uint32_t)(buf * FUSE_COUT), &s.mma_done[buf]);
      }
      __syncwarp();
    } else if (warp < 8 && it > 0) {
      fused_epilogue(&tmOut, s.mma_done, tmem_base, buf ^ 1, (uint32_t)((it - 1) >> 1) & 1u, s.epi + warp * 2048, s.s_scale2,
                     s.s_shift2, a.leaky2, prev_img, prev_p0, prev_q0, X, a.in_ld, a.Ho, a.Wo);
    }
    prev_img = img; prev_p0 = p0; prev_q0 = q0;
  }
  if (warp < 8 && it > 0) {
    fused_epilogue(&tmOut, s.mma_done, tmem_base, (it - 1) & 1, (uint32_t)((it - 1) >> 1) & 1u, s.epi + warp * 2048, s.s_scale2,
                   s.s_shift2, a.leaky2, prev_img, prev_p0, prev_q0, X, a.in_ld, a.Ho, a.Wo);
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<2 * FUSE_COUT>(tmem_base);
  }
}

}  // namespace yb
