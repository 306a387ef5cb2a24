// conv_fused.cuh -- two-layer fused convolutions for the HBM-bound head of Darknet-53 (sm_100a only).
//
// The first layers of the reference's YOLOv3 graph (net/v3.py:25-27 -> net/layers.py:17-67) move far more bytes than
// they compute on: conv 3->32 @416^2 writes 1.4 GB per 128 images only for the stride-2 conv behind it to read them
// back nine times through im2col TMA loads.  Here ONE kernel computes both layers of such a pair:
//
//   producer conv (small K, on mma.sync)  ->  bf16 activations written straight into K-major, 64-byte-swizzled
//   operand tiles of the 3x3 consumer conv in shared memory  ->  tcgen05.mma (M=128, N=64, K=9*32), accumulator in
//   TMEM  ->  BN / leaky (/ residual) epilogue  ->  TMA store.
//
// The intermediate activation never exists in global memory.  An output tile is 8 rows x 16 columns of the consumer
// conv (128 GEMM rows, row m = r*16 + c), so the producer recomputes only the one-pixel halo between tiles
// (17x33 producer pixels per 8x16 stride-2 outputs: x1.10; 10x18 per 8x16 stride-1 outputs: x1.41).
//
//   STEM  (net/v3.py:25-26): conv 3x3 s1 3->32 + BN + leaky  ->  conv 3x3 s2 32->64 + BN + leaky
//   BLOCK (net/v3.py:16-19,27): conv 1x1 64->32 + BN + leaky  ->  conv 3x3 s1 32->64 + BN + leaky  ->  + block input
//
// Operand layout.  A full im2col tile would hold every producer pixel nine times.  Instead the producer patch is kept
// as one copy per horizontal tap kw (and, for stride 2, per row parity): copy_kw[row][c] = P[row][S*c + kw], 16
// columns = 1024 bytes per row.  Tap (kh, kw) of the whole 8x16 tile is then the CONTIGUOUS 128-row block of copy_kw
// that starts kh (stride 1) or kh/2 (stride 2, parity kh&1) rows down -- a plain K-major SWIZZLE_64B operand whose
// start address is a multiple of 1024 bytes.  A producer pixel is stored 3 times (stride 1) or at most twice (stride 2).
//
// Arithmetic is that of the unfused kernels: the STEM producer runs the mma.sync sequence of conv_first_mma_kernel
// (bit-identical results), the BLOCK producer sums its K=64 in four mma.sync steps (fp32 summation order differs from
// the tcgen05 1x1 conv), the consumer's K walk (tap-major, two K=16 MMAs per tap) is conv_tc_persist_kernel<64,32>'s.
//
// Warp roles (one persistent CTA per SM, 384 threads):
//   warps 0-7   producers: stage the input patch (prefetched into registers one tile ahead), producer conv, operand
//               stores into A[tile&1]; named barrier 1 among themselves, mbarrier a_full[] towards the MMA issuer
//   warps 8-11  epilogue, TMEM lane quarter = warp & 3, both 32-column chunks: waits mma_done, BN/leaky(/residual),
//               swizzled staging, 4-D TMA store; arrives on acc_empty[].  Warp 8 also owns TMEM and, ahead of its
//               epilogue of tile i-1, issues the MMAs of tile i through one elected lane (waits a_full / acc_empty,
//               18 MMAs, commit -> mma_done[]).
// mma_done[b] doubles as "A[b] may be overwritten" for the producers.
#pragma once
#include "aux_kernels.cuh"
#include "conv_tc.cuh"

namespace yb {

constexpr int FUSE_PRODUCER_WARPS = 8;
constexpr int FUSE_PRODUCER_THREADS = FUSE_PRODUCER_WARPS * 32;
constexpr int FUSE_MMA_WARP = 8;                          // the first epilogue warp also owns TMEM and issues the MMAs
constexpr int FUSE_EPI_WARP0 = 8, FUSE_EPI_WARPS = 4;
constexpr int FUSE_THREADS = (FUSE_EPI_WARP0 + FUSE_EPI_WARPS) * 32;     // 384 (12 warps: 168 registers per thread)
constexpr int FUSE_TH = 8, FUSE_TW = 16;                  // output tile of the consumer conv
constexpr int FUSE_CMID = 32, FUSE_COUT = 64;
constexpr int FUSE_ROW_BYTES = FUSE_TW * FUSE_CMID * 2;   // one operand row of 16 pixels: 1024 bytes
constexpr int FUSE_BTAP_BYTES = FUSE_COUT * FUSE_CMID * 2;
constexpr int FUSE_B_BYTES = 9 * FUSE_BTAP_BYTES;         // 36,864
constexpr int FUSE_EPI_BYTES = FUSE_EPI_WARPS * 2 * 4096; // per epilogue warp two staging buffers of 32 rows x 128 B
constexpr int FUSE_HEADER = 1024;                         // barriers, TMEM pointer, consumer scale/shift
// STEM: operand = 3 kw copies x (9 even + 8 odd producer rows); input patch 19 rows x 112 bf16 (pixels 2*q0-4 ..
// 2*q0+33, 3 channels; 16-byte aligned row starts in global memory for float and uint8 images alike)
constexpr int STEM_COPY_ROWS = 17, STEM_ODD_ROW0 = 9;
constexpr int STEM_A_BYTES = 3 * STEM_COPY_ROWS * FUSE_ROW_BYTES;          // 52,224
constexpr int STEM_PATCH_ROWS = 19, STEM_PATCH_PITCH = 112, STEM_PATCH_WORDS = STEM_PATCH_ROWS * 28;
constexpr int STEM_PATCH_BYTES = ((STEM_PATCH_ROWS * STEM_PATCH_PITCH * 2 + 16) + 127) / 128 * 128;
// BLOCK: operand = 3 kw copies x 10 producer rows; input patch 10 x 18 pixels x 64 channels bf16, pixel pitch 144 bytes
// (conflict-free ldmatrix rows)
constexpr int BLOCK_COPY_ROWS = 10;
constexpr int BLOCK_A_BYTES = 3 * BLOCK_COPY_ROWS * FUSE_ROW_BYTES;        // 30,720
constexpr int BLOCK_PATCH_PIX = 10 * 18, BLOCK_PATCH_PITCH = 144;
constexpr int BLOCK_PATCH_BYTES = (BLOCK_PATCH_PIX * BLOCK_PATCH_PITCH + 127) / 128 * 128;    // 25,984
constexpr int FUSE_SMEM_STEM = 1024 + FUSE_HEADER + FUSE_B_BYTES + 2 * STEM_A_BYTES + FUSE_EPI_BYTES + 2 * STEM_PATCH_BYTES;
constexpr int FUSE_SMEM_BLOCK = 1024 + FUSE_HEADER + FUSE_B_BYTES + 2 * BLOCK_A_BYTES + FUSE_EPI_BYTES + 2 * BLOCK_PATCH_BYTES;

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
  unsigned long long* dbg;       // optional [16] cycle counters summed over CTAs (FUSE_DBG_*); nullptr = off
};
// cycle counters (engine option "cycles"), per role: lane 0 of producer warp 0, the MMA-issuing lane of warp 8, lane 0 of
// epilogue warp 9
enum FuseDbg {
  FUSE_DBG_PROD_STAGE = 0,    // producer: registers -> patch, named barrier, next tile's loads issued
  FUSE_DBG_PROD_WAIT_A,       // producer: waiting for the MMAs that still read this operand buffer
  FUSE_DBG_PROD_CONV,         // producer: mma.sync conv + operand stores + fence + arrive
  FUSE_DBG_PROD_TOTAL,
  FUSE_DBG_MMA_WAIT_A,        // MMA warp: waiting for the operand
  FUSE_DBG_MMA_WAIT_ACC,      // MMA warp: waiting for the epilogue to drain the accumulator
  FUSE_DBG_MMA_TOTAL,
  FUSE_DBG_EPI_WAIT_MMA,      // epilogue: waiting for the accumulator
  FUSE_DBG_EPI_WAIT_RES,      // epilogue: waiting for the residual tile
  FUSE_DBG_EPI_TOTAL,
  FUSE_DBG_TILES,
  FUSE_DBG_COUNT
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producer_bar() {          // named barrier 1: the 256 producer threads only
  asm volatile("bar.sync 1, %0;" ::"n"(FUSE_PRODUCER_THREADS) : "memory");
}
// byte offset of 16-byte chunk t of operand row `row` (64-byte rows, SWIZZLE_64B: chunk ^= (row >> 1) & 3) -- the
// layout TMA writes and make_kmajor_desc<32> describes; the operand buffers are 1024-byte aligned
__device__ __forceinline__ uint32_t operand_ofs(int row, int t) { return (uint32_t)(row * 64 + ((t ^ ((row >> 1) & 3)) << 4)); }

struct FuseSmem {
  uint64_t *a_full, *mma_done, *acc_empty, *b_full, *res_bar;   // [2] [2] [2] [1] [4 warps][2]
  uint32_t* tmem_ptr;
  float *s_scale2, *s_shift2;
  uint8_t *bs, *a0, *epi, *patch;
};
template <int A_BYTES>
__device__ __forceinline__ FuseSmem fuse_carve(uint8_t* smem) {
  FuseSmem s;
  s.a_full = reinterpret_cast<uint64_t*>(smem);
  s.mma_done = s.a_full + 2;
  s.acc_empty = s.mma_done + 2;
  s.b_full = s.acc_empty + 2;
  s.res_bar = s.b_full + 1;                                      // 8 barriers: bytes [56, 120)
  s.tmem_ptr = reinterpret_cast<uint32_t*>(smem + 128);
  s.s_scale2 = reinterpret_cast<float*>(smem + 256);
  s.s_shift2 = s.s_scale2 + FUSE_COUT;
  s.bs = smem + FUSE_HEADER;
  s.a0 = s.bs + FUSE_B_BYTES;
  s.epi = s.a0 + 2 * A_BYTES;
  s.patch = s.epi + FUSE_EPI_BYTES;
  return s;
}

// Common prologue: barriers, TMEM (128 columns = two 64-column accumulators), the consumer's weights (9 taps x
// [64 x 32] through the conv's own weight map), its scale/shift.
__device__ __forceinline__ uint32_t fuse_prologue(const FuseSmem& s, const CUtensorMap* tmB, const CUtensorMap* tmOut,
                                                  const CUtensorMap* tmRes, const FuseArgs& a) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(tmB);
    tma_prefetch_desc(tmOut);
    if (tmRes) tma_prefetch_desc(tmRes);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.a_full[i], FUSE_PRODUCER_WARPS);
      mbar_init(&s.mma_done[i], 1);
      mbar_init(&s.acc_empty[i], FUSE_EPI_WARPS);
    }
    mbar_init(s.b_full, 1);
    for (int i = 0; i < 2 * FUSE_EPI_WARPS; ++i) mbar_init(&s.res_bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == FUSE_MMA_WARP) {
    tmem_alloc<2 * FUSE_COUT>(s.tmem_ptr);
    if (elect_one()) {
      mbar_expect_tx(s.b_full, (uint32_t)FUSE_B_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(tmB, s.b_full, s.bs + tap * FUSE_BTAP_BYTES, tap * FUSE_CMID, 0);
    }
    __syncwarp();
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

struct TileWalk {
  int tiles_per_img, tiles_w;
  __device__ __forceinline__ void decode(int tile, int& img, int& p0, int& q0) const {
    img = tile / tiles_per_img;
    const int rem = tile - img * tiles_per_img;
    const int th = rem / tiles_w;
    p0 = th * FUSE_TH; q0 = (rem - th * tiles_w) * FUSE_TW;
  }
};

// ---- MMA issue of one tile (one elected lane of warp 8): 9 taps x 2 (K = 16) tcgen05.mma, M = 128, N = 64 ----
// STRIDE2: tap (kh, kw) starts at copy kw, parity block kh & 1, row kh >> 1; else at copy kw, row kh.
template <bool STRIDE2, int A_BYTES>
__device__ __forceinline__ void fuse_issue_tile(const FuseSmem& s, uint32_t tmem_base, int it, long long* t_dbg) {
  constexpr uint32_t idesc = make_idesc<FUSE_COUT>();
  constexpr int COPY_BYTES = (STRIDE2 ? STEM_COPY_ROWS : BLOCK_COPY_ROWS) * FUSE_ROW_BYTES;
  const int buf = it & 1;
  const uint32_t use = (uint32_t)(it >> 1);
  long long t0 = t_dbg ? clk() : 0;
  if (it == 0) mbar_wait(s.b_full, 0);
  if (it >= 2) mbar_wait(&s.acc_empty[buf], (use - 1u) & 1u);          // the epilogue has drained this accumulator
  if (t_dbg) { const long long t1 = clk(); t_dbg[1] += t1 - t0; t0 = t1; }
  mbar_wait(&s.a_full[buf], use & 1u);
  if (t_dbg) t_dbg[0] += clk() - t0;
  tc_fence_after();
  const uint32_t a_addr = smem_u32(s.a0) + (uint32_t)(buf * A_BYTES);
  const uint32_t b_addr = smem_u32(s.bs);
  const uint32_t tmem_d = tmem_base + (uint32_t)(buf * FUSE_COUT);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int kh = tap / 3, kw = tap - kh * 3;
    const int row0 = STRIDE2 ? ((kh & 1) * STEM_ODD_ROW0 + (kh >> 1)) : kh;
    const uint64_t da = make_kmajor_desc<FUSE_CMID>(a_addr + (uint32_t)(kw * COPY_BYTES + row0 * FUSE_ROW_BYTES));
    const uint64_t db = make_kmajor_desc<FUSE_CMID>(b_addr + (uint32_t)(tap * FUSE_BTAP_BYTES));
#pragma unroll
    for (int k = 0; k < FUSE_CMID / 16; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
  }
  umma_commit(&s.mma_done[buf]);             // accumulator ready; the operand buffer may be overwritten
}

// ---- epilogue warps: TMEM lane quarter = warp & 3 (tile rows 2q, 2q+1), all 64 columns of the tile ----
// BN + leaky (+ residual, fetched by TMA into the staging buffer the result is then written to) in registers, bf16
// through 128-byte-swizzled staging, one 4-D TMA store of a [2 rows][16 cols][64 channels] box per tile and warp.
// Warp 8 runs one tile behind: it issues the MMAs of tile i before it drains tile i-1.
template <bool RES, bool STRIDE2, int A_BYTES>
__device__ __forceinline__ void fuse_consumer_role(const FuseSmem& s, const CUtensorMap* tmOut, const CUtensorMap* tmRes,
                                                   uint32_t tmem_base, const FuseArgs& a, const TileWalk& tw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, ew = warp - FUSE_EPI_WARP0;
  const bool issuer = warp == FUSE_MMA_WARP;
  uint64_t* rbar = s.res_bar + ew * 2;
  const bool dbg = a.dbg != nullptr && ew == 1 && lane == 0;            // a pure epilogue warp
  const bool dbg_mma = a.dbg != nullptr && issuer;
  long long t_mma = 0, t_res = 0;
  long long t_issue[2] = {0, 0};
  const long long t_begin = (dbg || dbg_mma) ? clk() : 0;

  auto drain = [&](int it, int tile) {
    const int acc = it & 1;
    int img, p0, q0;
    tw.decode(tile, img, p0, q0);
    const int prow = p0 + 2 * quarter;
    // one staging buffer per warp: 32 rows (pixels) x 128 bytes (64 channels), SWIZZLE_128B: 16-byte chunk j of row r
    // sits at chunk j ^ (r & 7)
    uint8_t* buf = s.epi + (ew * 2 + (it & 1)) * 4096;
    const uint32_t row = smem_u32(buf) + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)lane & 7u;
    if (lane == 0) {
      tma_store_wait_read<1>();                    // the store of two tiles ago has read this staging buffer
      if (RES) {
        mbar_expect_tx(&rbar[it & 1], 4096u);
        tma_load_4d(tmRes, &rbar[it & 1], buf, 0, q0, prow, img);
      }
    }
    __syncwarp();
    long long t0 = dbg ? clk() : 0;
    mbar_wait(&s.mma_done[acc], (uint32_t)(it >> 1) & 1u);
    if (dbg) t_mma += clk() - t0;
    tc_fence_after();
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * FUSE_COUT);
    tmem_ld_32x32(taddr, v0);
    tmem_ld_32x32(taddr + 32u, v1);
    if (RES) {
      t0 = dbg ? clk() : 0;
      mbar_wait(&rbar[it & 1], (uint32_t)(it >> 1) & 1u);
      if (dbg) t_res += clk() - t0;
    }
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 8; ++g) {                  // 8 channels = one 16-byte chunk per step
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        float y = __uint_as_float(c < 32 ? v0[c] : v1[c - 32]) * s.s_scale2[c] + s.s_shift2[c];
        if (a.leaky2) y = fmaxf(y, 0.1f * y);
        f[j] = y;
      }
      const uint32_t addr = row + (((uint32_t)g ^ sw) << 4);
      if (RES) {
        const uint4 rv = lds128(addr);
        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[2 * j] += __uint_as_float(rw[j] << 16);
          f[2 * j + 1] += __uint_as_float(rw[j] & 0xFFFF0000u);
        }
      }
      sts128(addr, pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&s.acc_empty[acc]);              // the accumulator is in registers / staging: hand it back early
      tma_store_4d(tmOut, buf, 0, q0, prow, img);
      tma_store_commit();
    }
  };

  int it = 0;
  if (issuer) {
    int prev_tile = -1;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      if (elect_one()) fuse_issue_tile<STRIDE2, A_BYTES>(s, tmem_base, it, dbg_mma ? t_issue : nullptr);
      __syncwarp();
      if (prev_tile >= 0) drain(it - 1, prev_tile);
      prev_tile = tile;
    }
    if (prev_tile >= 0) drain(it - 1, prev_tile);
  } else {
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) drain(it, tile);
  }
  if (lane == 0) tma_store_wait_all();
  if (dbg_mma && lane == 0) {
    atomicAdd(&a.dbg[FUSE_DBG_MMA_WAIT_A], (unsigned long long)t_issue[0]);
    atomicAdd(&a.dbg[FUSE_DBG_MMA_WAIT_ACC], (unsigned long long)t_issue[1]);
    atomicAdd(&a.dbg[FUSE_DBG_MMA_TOTAL], (unsigned long long)(clk() - t_begin));
  }
  if (dbg) {
    atomicAdd(&a.dbg[FUSE_DBG_EPI_WAIT_MMA], (unsigned long long)t_mma);
    atomicAdd(&a.dbg[FUSE_DBG_EPI_WAIT_RES], (unsigned long long)t_res);
    atomicAdd(&a.dbg[FUSE_DBG_EPI_TOTAL], (unsigned long long)(clk() - t_begin));
    atomicAdd(&a.dbg[FUSE_DBG_TILES], (unsigned long long)it);
  }
}

// ================================================================================================
// STEM: conv 3x3 s1 3->32 (mma.sync m16n8k16, K = 27 padded to 32) feeding conv 3x3 s2 32->64 (tcgen05)
// ================================================================================================
// Producer pixels of a tile, relative to (2*p0 - 1, 2*q0 - 1): yy in [0,17), xx in [0,33).  They are walked by parity
// class so that the 16 pixels of an mma tile land in consecutive operand rows (conflict-free stores):
//   mt  0.. 7  yy odd,  xx odd   (row j = mt, 16 columns)             -> copy kw=1
//   mt  8..16  yy even, xx odd   (9 rows x 16)                        -> copy kw=1
//   mt 17..25  yy odd,  xx even  (8 rows x 17, flattened, 136 pixels) -> copies kw=0 (c = xx/2) and kw=2 (c = xx/2 - 1)
//   mt 26..35  yy even, xx even  (9 rows x 17, flattened, 153 pixels) -> copies kw=0 and kw=2
// Producer warp w takes mma tiles w, w+8, ... (< 36): five for warps 0-3, four for warps 4-7.
template <bool U8>
__global__ void __launch_bounds__(FUSE_THREADS, 1)
stem_fused_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut, const FuseArgs a) {
  extern __shared__ __align__(1024) uint8_t fuse_smem_raw[];
  uint8_t* smem = fuse_smem_raw + ((1024u - (smem_u32(fuse_smem_raw) & 1023u)) & 1023u);
  const FuseSmem s = fuse_carve<STEM_A_BYTES>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  const uint32_t tmem_base = fuse_prologue(s, &tmB, &tmOut, nullptr, a);
  const TileWalk tw{a.tiles_h * a.tiles_w, a.tiles_w};

  if (warp < FUSE_PRODUCER_WARPS) {
    const int g = lane >> 2, t = lane & 3;
    // ---- constants: B fragments of the first conv (permuted columns: column g of n-tile nt holds channel
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
      const int kk = k < K1 ? k : 0;                // padded columns read a valid element; their weights are zero
      const int tap = kk / 3, ci = kk - tap * 3;
      soff[j] = (tap / 3) * STEM_PATCH_PITCH + 6 + (tap % 3) * 3 + ci;
    }
    // ---- tile-invariant tables of this thread's producer pixels b = 2*k + r (mma tile warp + 8k, rows g + 8r) ----
    // gofs: patch element offset; dofs[.][j]: operand byte offset of its copy for column option j (kw = kw0 + 2j,
    // c = c0 - j), 0xFFFFFFFF = none; bmask bit b / 16 + b: pixel on the patch's first row / column (outside the
    // image when the tile touches the top / left border: the consumer's zero padding, not conv(0)).
    constexpr int NMT = 5;
    int gofs[2 * NMT];
    uint32_t dofs[2 * NMT][2];
    uint32_t bmask = 0u;
#pragma unroll
    for (int k = 0; k < NMT; ++k)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * k + r;
        const int mt = warp + FUSE_PRODUCER_WARPS * k, gi = g + 8 * r;
        int yy, xx;
        bool ok = mt < 36;
        if (mt < 8) { yy = 2 * mt + 1; xx = 2 * gi + 1; }
        else if (mt < 17) { yy = 2 * (mt - 8); xx = 2 * gi + 1; }
        else if (mt < 26) {
          int f = (mt - 17) * 16 + gi;
          ok = f < 136; f = ok ? f : 135;
          const int rr = f / 17;
          yy = 2 * rr + 1; xx = 2 * (f - rr * 17);
        } else {
          int f = (mt - 26) * 16 + gi;
          ok = ok && f < 153; f = f < 153 ? f : 152;
          const int rr = f / 17;
          yy = 2 * rr; xx = 2 * (f - rr * 17);
        }
        gofs[b] = yy * STEM_PATCH_PITCH + xx * 3;
        const int row_blk = ((yy & 1) ? STEM_ODD_ROW0 : 0) + (yy >> 1);
        const int kw0 = xx & 1, c0 = (xx - kw0) >> 1;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int kw = kw0 + 2 * j, c = c0 - j;
          const bool v = ok && kw <= 2 && c >= 0 && c < FUSE_TW;
          dofs[b][j] = v ? (uint32_t)(kw * STEM_COPY_ROWS * FUSE_ROW_BYTES) + operand_ofs(row_blk * FUSE_TW + c, t) : 0xFFFFFFFFu;
        }
        if (yy == 0) bmask |= 1u << b;
        if (xx == 0) bmask |= 1u << (16 + b);
      }
    const int n_mt = warp < 4 ? 5 : 4;                                // 36 = 4*5 + 4*4
    // ---- input patch prefetch: 19 rows x 28 words (a word is 4 uint8 or 4 floats), 256 threads x 3 ----
    const int row_bytes = a.W * 3;                  // elements per image row
    int pf_rr[3], pf_wq[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int e = threadIdx.x + i * FUSE_PRODUCER_THREADS;
      pf_rr[i] = e < STEM_PATCH_WORDS ? e / 28 : -1000000;             // out-of-range words never pass the row test
      pf_wq[i] = e - (e / 28) * 28;
    }
    uint4 pre[3];
    auto load_patch = [&](int img, int p0, int q0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        pre[i] = make_uint4(0u, 0u, 0u, 0u);
        const int y = 2 * p0 - 2 + pf_rr[i];
        const int bx = (2 * q0 - 4) * 3 + 4 * pf_wq[i];
        if ((unsigned)y < (unsigned)a.H && (unsigned)bx < (unsigned)row_bytes) {
          const long long off = ((long long)img * a.H + y) * row_bytes + bx;
          if (U8) pre[i].x = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(a.in) + off));
          else pre[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.in) + off));
        }
      }
    };
    auto store_patch = [&](uint16_t* s_patch) {
      constexpr float k255 = 0.003921568859368562698f;   // same conversion as conv_first_mma_kernel (bit-identical bf16)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int e = threadIdx.x + i * FUSE_PRODUCER_THREADS;
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
          *reinterpret_cast<uint2*>(s_patch + 4 * e) = o;          // row rr, word wq: element rr*112 + 4*wq = 4*e
        }
      }
    };
    const bool dbg = a.dbg != nullptr && threadIdx.x == 0;
    long long t_stage = 0, t_wait = 0, t_conv = 0;
    const long long t_begin = dbg ? clk() : 0;
    int it = 0;
    int img = 0, p0 = 0, q0 = 0;
    if ((int)blockIdx.x < a.n_tiles) { tw.decode(blockIdx.x, img, p0, q0); load_patch(img, p0, q0); }
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      long long t0 = dbg ? clk() : 0;
      uint16_t* s_patch = reinterpret_cast<uint16_t*>(s.patch + buf * STEM_PATCH_BYTES);
      store_patch(s_patch);
      producer_bar();       // patch[buf] complete; also: every producer is past its reads of patch[buf] two tiles ago
      int n_img = 0, n_p0 = 0, n_q0 = 0;
      if (tile + (int)gridDim.x < a.n_tiles) {           // next tile's patch: in flight during the producer conv
        tw.decode(tile + gridDim.x, n_img, n_p0, n_q0);
        load_patch(n_img, n_p0, n_q0);
      }
      if (dbg) { const long long t1 = clk(); t_stage += t1 - t0; t0 = t1; }
      if (it >= 2) mbar_wait(&s.mma_done[buf], (uint32_t)((it >> 1) - 1) & 1u);     // the MMAs of tile it-2 have read A[buf]
      if (dbg) { const long long t1 = clk(); t_wait += t1 - t0; t0 = t1; }
      const uint32_t a_base = smem_u32(s.a0) + (uint32_t)(buf * STEM_A_BYTES);
      const uint32_t kill = bmask & ((p0 == 0 ? 0x3FFu : 0u) | (q0 == 0 ? 0x3FF0000u : 0u));
#pragma unroll
      for (int k = 0; k < NMT; ++k) {
        if (k < NMT - 1 || n_mt == NMT) {              // only the last mma tile is conditional (warps 0-3)
          uint32_t afrag[2][4];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint16_t* p = s_patch + gofs[2 * k + r];
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
            const int b = 2 * k + r;
            uint32_t pk[4];
            const bool inside = ((kill >> b) & 0x10001u) == 0u;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              float y0v = acc[nt][2 * r] * sc[2 * nt] + sh[2 * nt];
              float y1v = acc[nt][2 * r + 1] * sc[2 * nt + 1] + sh[2 * nt + 1];
              if (a.leaky1) { y0v = fmaxf(y0v, 0.1f * y0v); y1v = fmaxf(y1v, 0.1f * y1v); }
              pk[nt] = inside ? pack_bf16(y0v, y1v) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (dofs[b][j] != 0xFFFFFFFFu) sts128(a_base + dofs[b][j], pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      fence_proxy_async();             // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_full[buf]);
      if (dbg) t_conv += clk() - t0;
      img = n_img; p0 = n_p0; q0 = n_q0;
    }
    if (dbg) {
      atomicAdd(&a.dbg[FUSE_DBG_PROD_STAGE], (unsigned long long)t_stage);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_WAIT_A], (unsigned long long)t_wait);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_CONV], (unsigned long long)t_conv);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_TOTAL], (unsigned long long)(clk() - t_begin));
    }
  } else {
    fuse_consumer_role<false, true, STEM_A_BYTES>(s, &tmOut, nullptr, tmem_base, a, tw);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FUSE_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<2 * FUSE_COUT>(tmem_base);
  }
}

// ================================================================================================
// BLOCK: conv 1x1 64->32 (mma.sync m16n8k16, A fragments by ldmatrix) feeding conv 3x3 s1 32->64 (tcgen05) + residual
// ================================================================================================
// The residual block of net/v3.py:16-19 at 64 channels: X -> conv1x1(32) -> conv3x3(64) -> + X.  The producer needs the
// 10 x 18 input pixels around an 8 x 16 output tile; each 16-pixel mma tile is 16 consecutive pixels of that patch
// (180 pixels = 12 tiles: two for producer warps 0-3, one for warps 4-7; the last tile holds 4 pixels).  Producer
// pixels outside the image are the 3x3 conv's zero padding.  The residual tile is fetched by TMA (tmRes, the same
// tensor as the producer input) into the epilogue's staging buffers.
__global__ void __launch_bounds__(FUSE_THREADS, 1)
block_fused_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ CUtensorMap tmRes, const FuseArgs a) {
  extern __shared__ __align__(1024) uint8_t fuse_smem_raw[];
  uint8_t* smem = fuse_smem_raw + ((1024u - (smem_u32(fuse_smem_raw) & 1023u)) & 1023u);
  const FuseSmem s = fuse_carve<BLOCK_A_BYTES>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  const uint32_t tmem_base = fuse_prologue(s, &tmB, &tmOut, &tmRes, a);
  const TileWalk tw{a.tiles_h * a.tiles_w, a.tiles_w};

  if (warp < FUSE_PRODUCER_WARPS) {
    const int g = lane >> 2, t = lane & 3;
    // B fragments of the 1x1 conv: w1 is [32 cout][64 cin] bf16; b0 = (k = 2t, 2t+1; n), b1 = (k = 2t+8, 2t+9; n) with
    // the permuted column n = (g/2)*8 + 2*nt + (g&1), so a thread's accumulators are channels 8t..8t+7
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
    const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(a.in);
    // ---- tile-invariant tables of this thread's producer pixels b = 2*k + r: patch pixel f = (warp + 8k)*16 + g + 8r ----
    // dofs[b][kw]: operand byte offset in copy kw (column c = px - kw), 0xFFFFFFFF = none
    constexpr int NMT = 2;
    int ppy[2 * NMT], ppx[2 * NMT];
    uint32_t dofs[2 * NMT][3];
    uint32_t lm_ofs[NMT];
    // ldmatrix.x4 row address of this lane: matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
    const int lm_row = (lane & 7) + 8 * ((lane >> 3) & 1);
    const int lm_kofs = 8 * (lane >> 4);
#pragma unroll
    for (int k = 0; k < NMT; ++k) {
      const int f0 = (warp + FUSE_PRODUCER_WARPS * k) * 16;
      lm_ofs[k] = (uint32_t)(min(f0 + lm_row, BLOCK_PATCH_PIX - 1) * BLOCK_PATCH_PITCH + lm_kofs * 2);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * k + r;
        const int f = f0 + g + 8 * r;
        const int py = f / 18, px = f - py * 18;
        ppy[b] = py; ppx[b] = px;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int c = px - kw;
          const bool v = f < BLOCK_PATCH_PIX && c >= 0 && c < FUSE_TW;
          dofs[b][kw] = v ? (uint32_t)(kw * BLOCK_COPY_ROWS * FUSE_ROW_BYTES) + operand_ofs(py * FUSE_TW + c, t) : 0xFFFFFFFFu;
        }
      }
    }
    const int n_mt = warp < 4 ? 2 : 1;                                // 12 = 4*2 + 4*1
    // ---- input patch prefetch: 180 pixels x 8 chunks of 16 bytes = 1440 chunks, 256 threads x 6 ----
    int pf_py[6], pf_px[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int e = threadIdx.x + i * FUSE_PRODUCER_THREADS;
      const int px = e >> 3;
      pf_py[i] = e < BLOCK_PATCH_PIX * 8 ? px / 18 : -1000000;        // out-of-range chunks never pass the row test
      pf_px[i] = px - (px / 18) * 18;
    }
    uint4 pre[6];
    auto load_patch = [&](int img, int p0, int q0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        pre[i] = make_uint4(0u, 0u, 0u, 0u);
        const int y = p0 - 1 + pf_py[i], x = q0 - 1 + pf_px[i];
        if ((unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W)
          pre[i] = __ldg(reinterpret_cast<const uint4*>(X + (((long long)img * a.H + y) * a.W + x) * a.in_ld + (threadIdx.x & 7) * 8));
      }
    };
    auto store_patch = [&](uint32_t patch_addr) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int e = threadIdx.x + i * FUSE_PRODUCER_THREADS;
        if (e < BLOCK_PATCH_PIX * 8) sts128(patch_addr + (uint32_t)((e >> 3) * BLOCK_PATCH_PITCH + (e & 7) * 16), pre[i].x, pre[i].y, pre[i].z, pre[i].w);
      }
    };
    const bool dbg = a.dbg != nullptr && threadIdx.x == 0;
    long long t_stage = 0, t_wait = 0, t_conv = 0;
    const long long t_begin = dbg ? clk() : 0;
    int it = 0;
    int img = 0, p0 = 0, q0 = 0;
    if ((int)blockIdx.x < a.n_tiles) { tw.decode(blockIdx.x, img, p0, q0); load_patch(img, p0, q0); }
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      long long t0 = dbg ? clk() : 0;
      const uint32_t patch_addr = smem_u32(s.patch) + (uint32_t)(buf * BLOCK_PATCH_BYTES);
      store_patch(patch_addr);
      producer_bar();
      int n_img = 0, n_p0 = 0, n_q0 = 0;
      if (tile + (int)gridDim.x < a.n_tiles) {
        tw.decode(tile + gridDim.x, n_img, n_p0, n_q0);
        load_patch(n_img, n_p0, n_q0);
      }
      if (dbg) { const long long t1 = clk(); t_stage += t1 - t0; t0 = t1; }
      if (it >= 2) mbar_wait(&s.mma_done[buf], (uint32_t)((it >> 1) - 1) & 1u);
      if (dbg) { const long long t1 = clk(); t_wait += t1 - t0; t0 = t1; }
      const uint32_t a_base = smem_u32(s.a0) + (uint32_t)(buf * BLOCK_A_BYTES);
#pragma unroll
      for (int k = 0; k < NMT; ++k) {
        if (k < NMT - 1 || n_mt == NMT) {              // only the last mma tile is conditional (warps 0-3)
          float acc[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f; }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t af[4];
            ldmatrix_x4(patch_addr + lm_ofs[k] + (uint32_t)(ks * 32), af);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], af, bf[ks][nt][0], bf[ks][nt][1]);
          }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int b = 2 * k + r;
            const int y = p0 - 1 + ppy[b], x = q0 - 1 + ppx[b];
            const bool inside = (unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W;   // else: the 3x3 conv's zero padding
            uint32_t pk[4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              float y0v = acc[nt][2 * r] * sc[2 * nt] + sh[2 * nt];
              float y1v = acc[nt][2 * r + 1] * sc[2 * nt + 1] + sh[2 * nt + 1];
              if (a.leaky1) { y0v = fmaxf(y0v, 0.1f * y0v); y1v = fmaxf(y1v, 0.1f * y1v); }
              pk[nt] = inside ? pack_bf16(y0v, y1v) : 0u;
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              if (dofs[b][kw] != 0xFFFFFFFFu) sts128(a_base + dofs[b][kw], pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_full[buf]);
      if (dbg) t_conv += clk() - t0;
      img = n_img; p0 = n_p0; q0 = n_q0;
    }
    if (dbg) {
      atomicAdd(&a.dbg[FUSE_DBG_PROD_STAGE], (unsigned long long)t_stage);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_WAIT_A], (unsigned long long)t_wait);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_CONV], (unsigned long long)t_conv);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_TOTAL], (unsigned long long)(clk() - t_begin));
    }
  } else {
    fuse_consumer_role<true, false, BLOCK_A_BYTES>(s, &tmOut, &tmRes, tmem_base, a, tw);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FUSE_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<2 * FUSE_COUT>(tmem_base);
  }
}

}  // namespace yb
