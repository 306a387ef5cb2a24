// conv_fused.cuh -- two-layer fused convolutions for the HBM-bound head of Darknet-53 (sm_100a only).
//
// The first layers of the reference's YOLOv3 graph (net/v3.py:25-27 -> net/layers.py:17-67) move far more bytes than
// they compute on: conv 3->32 @416^2 writes 1.4 GB per 128 images only for the stride-2 conv behind it to read them
// back nine times through im2col TMA loads.  Here ONE kernel computes both layers of such a pair:
//
//   input patch by TMA (tile mode, zero-filled outside the image)  ->  producer conv (small K)  ->  bf16
//   activations written straight into K-major, 64-byte-swizzled operand tiles of the 3x3 consumer conv in shared
//   memory  ->  tcgen05.mma (M=128, N=64, K=9*32), accumulator in TMEM  ->  BN / leaky (/ residual) epilogue  ->  TMA store.
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
// (same K order, bit-identical results), the BLOCK producer is a tcgen05 1x1 conv like the stand-alone one (K = 64 in four
// K = 16 steps), the consumer's K walk (tap-major, two K=16 MMAs per tap) is conv_tc_persist_kernel<64,32>'s.  BN and the
// leaky multiply run on packed fp32 pairs (fma.rn.f32x2 / mul.rn.f32x2: the same IEEE operations, two per instruction).
//
// What bounds them: the first version was instruction-issue bound (~2 warp instructions per cycle per
// SM, half of them address arithmetic), so everything per tile costs few instructions: no integer division in the tile
// walk, every shared-memory address of the producer loops is one per-thread register plus a compile-time offset, bf16
// pairs of the first conv's im2col rows are fetched with one aligned 32-bit load from one of two patch copies (the
// second shifted by one element), input patches arrive by TMA (requested into L2 eight tiles ahead).  Then: warp-level
// mma.sync shares the tensor pipe with tcgen05.mma and stalls its warps while the consumer's MMAs run, and issuing 18
// MMAs blocks the issuing thread for most of their execution -- hence a warp of its own for MMA issue, and the BLOCK
// producer on tcgen05.  What is left is shared-memory bandwidth (operand reads of the MMAs + operand stores + staging).
//
// STEM warp roles (one persistent CTA per SM, 544 threads; BLOCK: see block_fused_kernel):
//   warps 0-11  producers: convert the raw uint8 / float patch to the two bf16 copies; producer conv on three mma tiles
//               of 16 pixels each, operand stores into A[tile&1]; named barrier 1 among themselves, mbarrier a_full[]
//               towards the MMA issuer
//   warps 12-15 epilogue, TMEM lane quarter = warp & 3, both 32-column chunks: waits mma_done, BN/leaky, swizzled
//               staging, 4-D TMA store; arrives on acc_empty[]
//   warp 16     owns TMEM; one elected lane requests the input patches (L2 prefetch 8 tiles ahead, TMA load 2 tiles ahead)
//               and issues the MMAs of every tile (waits a_full / acc_empty, 18 MMAs, commit -> mma_done[])
// mma_done[b] doubles as "A[b] may be overwritten" for the producers.
#pragma once
#include "aux_kernels.cuh"
#include "conv_tc.cuh"

namespace yb {

constexpr int FUSE_PRODUCER_WARPS = 12;
constexpr int FUSE_PRODUCER_THREADS = FUSE_PRODUCER_WARPS * 32;
constexpr int FUSE_EPI_WARP0 = 12, FUSE_EPI_WARPS = 4;
constexpr int FUSE_MMA_WARP = 16;                         // owns TMEM; one elected lane requests input patches and issues the MMAs
constexpr int FUSE_THREADS = (FUSE_MMA_WARP + 1) * 32;    // 544 (17 warps, register file allocated as for 20: 96 per thread)
constexpr int FUSE_TH = 8, FUSE_TW = 16;                  // output tile of the consumer conv
constexpr int FUSE_CMID = 32, FUSE_COUT = 64;
constexpr int FUSE_ROW_BYTES = FUSE_TW * FUSE_CMID * 2;   // one operand row of 16 pixels: 1024 bytes
constexpr int FUSE_BTAP_BYTES = FUSE_COUT * FUSE_CMID * 2;
constexpr int FUSE_B_BYTES = 9 * FUSE_BTAP_BYTES;         // 36,864
constexpr int FUSE_EPI_BYTES = FUSE_EPI_WARPS * 2 * 4096; // per epilogue warp two staging buffers of 32 rows x 128 B
constexpr int FUSE_HEADER = 1024;                         // barriers, TMEM pointer, consumer scale/shift
constexpr int FUSE_MAX_IN_STAGES = 4;
// STEM: operand = 3 kw copies x (9 even + 8 odd producer rows).  Input patch: 19 rows x 112 elements (pixels 2*q0-4 ..
// 2*q0+33, 3 channels), fetched raw by TMA (uint8: 128-byte rows starting 4 elements earlier so the box starts on a
// 16-byte boundary; float: 448-byte rows), converted to two bf16 copies (copy 1 = copy 0 shifted by one element)
constexpr int STEM_COPY_ROWS = 17, STEM_ODD_ROW0 = 9;
constexpr int STEM_COPY_BYTES = STEM_COPY_ROWS * FUSE_ROW_BYTES;           // 17,408
constexpr int STEM_A_BYTES = 3 * STEM_COPY_BYTES;                          // 52,224
constexpr int STEM_PATCH_ROWS = 19, STEM_PATCH_PITCH = 112, STEM_PATCH_WORDS = STEM_PATCH_ROWS * 28;
constexpr int STEM_PCOPY_BYTES = 4352;                                     // 19 * 112 * 2 = 4256, padded
constexpr int STEM_PBUF_BYTES = 2 * STEM_PCOPY_BYTES;
constexpr int STEM_IN_STAGES = 3;
constexpr int STEM_RAW_ROW_U8 = 128, STEM_RAW_ROW_F32 = STEM_PATCH_PITCH * 4;
constexpr int STEM_RAW_TX_U8 = STEM_PATCH_ROWS * STEM_RAW_ROW_U8, STEM_RAW_TX_F32 = STEM_PATCH_ROWS * STEM_RAW_ROW_F32;
constexpr int STEM_RAW_STAGE_U8 = 2560, STEM_RAW_STAGE_F32 = 8576;         // 2432 / 8512 + slack, multiples of 128
// BLOCK: operand = 3 kw copies x 10 producer rows; input patch 10 x 18 pixels x 64 channels bf16 by one 4-D TMA box,
// SWIZZLE_128B (16-byte chunk j of pixel f at chunk j ^ (f & 7)): conflict-free ldmatrix rows and residual reads
constexpr int BLOCK_COPY_ROWS = 10;
constexpr int BLOCK_COPY_BYTES = BLOCK_COPY_ROWS * FUSE_ROW_BYTES;
constexpr int BLOCK_A_BYTES = 3 * BLOCK_COPY_BYTES;                        // 30,720
constexpr int BLOCK_PATCH_W = 18, BLOCK_PATCH_H = 10, BLOCK_PATCH_PIX = BLOCK_PATCH_W * BLOCK_PATCH_H;
constexpr int BLOCK_PATCH_TX = BLOCK_PATCH_PIX * 128;                      // 23,040
constexpr int BLOCK_PATCH_STAGE = 23552;                                   // multiple of 1024
constexpr int BLOCK_IN_STAGES = 4;                                         // a patch lives until the epilogue has read the residual
template <bool U8>
constexpr int fuse_smem_stem() {
  return 1024 + FUSE_HEADER + FUSE_B_BYTES + 2 * STEM_A_BYTES + FUSE_EPI_BYTES + 2 * STEM_PBUF_BYTES +
         STEM_IN_STAGES * (U8 ? STEM_RAW_STAGE_U8 : STEM_RAW_STAGE_F32);
}

struct FuseArgs {
  int n_img, Ho, Wo;             // consumer output size; Ho % 8 == 0, Wo % 16 == 0
  int H, W;                      // producer input size (STEM: 2*Ho x 2*Wo image; BLOCK: Ho x Wo)
  const void* w1;                // STEM: fp32 [27][32] (tap*3+ci major); BLOCK: bf16 [32][64] (cout major, K contiguous)
  const float *scale1, *shift1;  // producer BN (32 channels)
  float scale2[FUSE_COUT], shift2[FUSE_COUT];   // consumer BN, by value: warp-uniform constant-bank operands in the epilogue
  int leaky1, leaky2;
  int tiles_h, tiles_w, n_tiles; // Ho/8, Wo/16, n_img * tiles_h * tiles_w
  unsigned long long* dbg;       // optional [16] cycle counters summed over CTAs (FUSE_DBG_*); nullptr = off
};
// cycle counters (engine option "cycles"), per role: lane 0 of producer warp 0, the MMA-issuing lane of warp 12, lane 0 of
// epilogue warp 13
enum FuseDbg {
  FUSE_DBG_PROD_STAGE = 0,    // producer: waiting for the input patch, conversion (STEM), named barrier
  FUSE_DBG_PROD_WAIT_A,       // producer: waiting for the MMAs that still read this operand buffer
  FUSE_DBG_PROD_CONV,         // producer: mma.sync conv + operand stores + fence + arrive
  FUSE_DBG_PROD_TOTAL,
  FUSE_DBG_MMA_WAIT_A,        // MMA warp: waiting for the operand
  FUSE_DBG_MMA_WAIT_ACC,      // MMA warp: waiting for the epilogue to drain the accumulator
  FUSE_DBG_MMA_TOTAL,
  FUSE_DBG_EPI_WAIT_MMA,      // epilogue: waiting for the accumulator
  FUSE_DBG_EPI_WAIT_RES,      // (unused since the residual comes from the input patch)
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
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// tile-mode loads need a 16-byte aligned inner start (coordinate * element size); other coordinates may be negative or
// past the tensor: those elements are zero-filled (tools/probes/tma_patch_probe.cu)
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// request a box into L2 only: hides the DRAM latency (~2 us under this kernel's write traffic) without holding shared memory
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return (uint32_t)v;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producer_bar() {          // named barrier 1: the producer threads only
  asm volatile("bar.sync 1, %0;" ::"n"(FUSE_PRODUCER_THREADS) : "memory");
}
// byte offset of 16-byte chunk t of operand row `row` (64-byte rows, SWIZZLE_64B: chunk ^= (row >> 1) & 3) -- the
// layout TMA writes and make_kmajor_desc<32> describes; the operand buffers are 1024-byte aligned
__device__ __forceinline__ uint32_t operand_ofs(int row, int t) { return (uint32_t)(row * 64 + ((t ^ ((row >> 1) & 3)) << 4)); }

struct FuseSmem {
  uint64_t *a_full, *mma_done, *acc_empty, *b_full, *in_full, *in_empty;   // [2] [2] [2] [1] [4] [4]
  uint32_t* tmem_ptr;
  uint8_t *bs, *a0, *epi, *in;   // `in`: STEM bf16 patch buffers then raw stages; BLOCK: patch stages
};
template <int A_BYTES>
__device__ __forceinline__ FuseSmem fuse_carve(uint8_t* smem) {
  FuseSmem s;
  s.a_full = reinterpret_cast<uint64_t*>(smem);
  s.mma_done = s.a_full + 2;
  s.acc_empty = s.mma_done + 2;
  s.b_full = s.acc_empty + 2;
  s.in_full = s.b_full + 1;
  s.in_empty = s.in_full + FUSE_MAX_IN_STAGES;                   // 15 barriers: bytes [0, 120)
  s.tmem_ptr = reinterpret_cast<uint32_t*>(smem + 128);
  s.bs = smem + FUSE_HEADER;
  s.a0 = s.bs + FUSE_B_BYTES;
  s.epi = s.a0 + 2 * A_BYTES;
  s.in = s.epi + FUSE_EPI_BYTES;
  return s;
}

// Common prologue: barriers, TMEM (128 columns = two 64-column accumulators), the consumer's weights (9 taps x
// [64 x 32] through the conv's own weight map).
__device__ __forceinline__ uint32_t fuse_prologue(const FuseSmem& s, const CUtensorMap* tmIn, const CUtensorMap* tmB,
                                                  const CUtensorMap* tmOut, const FuseArgs& a, int in_stages, int in_empty_count) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(tmIn);
    tma_prefetch_desc(tmB);
    tma_prefetch_desc(tmOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.a_full[i], FUSE_PRODUCER_WARPS);
      mbar_init(&s.mma_done[i], 1);
      mbar_init(&s.acc_empty[i], FUSE_EPI_WARPS);
    }
    mbar_init(s.b_full, 1);
    for (int i = 0; i < in_stages; ++i) {
      mbar_init(&s.in_full[i], 1);
      mbar_init(&s.in_empty[i], in_empty_count);
    }
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *s.tmem_ptr;
}

// Tile walk without divisions in the loop: tile = first + j * step, as (image, tile row, tile column)
struct TileWalk {
  int tiles_w, tiles_h, sw, sh, si;
  int img, th, tw;
  __device__ __forceinline__ void init(int tiles_w_, int tiles_h_, int first, int step) {
    tiles_w = tiles_w_; tiles_h = tiles_h_;
    const int per_img = tiles_w * tiles_h;
    si = step / per_img;
    int rem = step - si * per_img;
    sh = rem / tiles_w; sw = rem - sh * tiles_w;
    img = first / per_img;
    rem = first - img * per_img;
    th = rem / tiles_w; tw = rem - th * tiles_w;
  }
  __device__ __forceinline__ void advance() {
    tw += sw;
    if (tw >= tiles_w) { tw -= tiles_w; ++th; }
    th += sh;
    if (th >= tiles_h) { th -= tiles_h; ++img; }
    img += si;
  }
  __device__ __forceinline__ int p0() const { return th * FUSE_TH; }
  __device__ __forceinline__ int q0() const { return tw * FUSE_TW; }
};

// ---- MMA issue of one tile (one elected lane of warp 12): 9 taps x 2 (K = 16) tcgen05.mma, M = 128, N = 64 ----
// STRIDE2: tap (kh, kw) starts at copy kw, parity block kh & 1, row kh >> 1; else at copy kw, row kh.
template <bool STRIDE2, int A_BYTES>
__device__ __forceinline__ void fuse_issue_tile(const FuseSmem& s, uint32_t tmem_base, int it, long long* t_dbg) {
  constexpr uint32_t idesc = make_idesc<FUSE_COUT>();
  constexpr int COPY_BYTES = STRIDE2 ? STEM_COPY_BYTES : BLOCK_COPY_BYTES;
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
// BN + leaky (+ residual: the centre of the producer's input patch, still in shared memory) in registers, bf16 through
// 128-byte-swizzled staging, one 4-D TMA store of a [2 rows][16 cols][64 channels] box per tile and warp.
template <bool RES, int IN_STAGES, int EPI_WARP0>
__device__ __forceinline__ void fuse_epilogue_role(const FuseSmem& s, const CUtensorMap* tmOut, uint32_t tmem_base, const FuseArgs& a,
                                                   uint32_t patch0, int patch_stage_bytes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, ew = warp - EPI_WARP0;
  const bool dbg = a.dbg != nullptr && ew == 1 && lane == 0;
  long long t_mma = 0;
  const long long t_begin = dbg ? clk() : 0;
  // residual position of this lane in the input patch: output pixel (2*quarter + lane/16, lane%16) -> patch pixel (+1, +1)
  const int res_f = (2 * quarter + (lane >> 4) + 1) * BLOCK_PATCH_W + (lane & 15) + 1;
  const uint32_t res_ofs = (uint32_t)res_f * 128u, res_sw = (uint32_t)res_f & 7u;
  int res_stage = 0;
  uint32_t res_phase = 0;
  const bool leaky2 = a.leaky2 != 0;
  TileWalk tw;
  tw.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
  int it = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
    const int acc = it & 1;
    const int prow = tw.p0() + 2 * quarter;
    // one staging buffer per warp: 32 rows (pixels) x 128 bytes (64 channels), SWIZZLE_128B: 16-byte chunk j of row r
    // sits at chunk j ^ (r & 7)
    uint8_t* buf = s.epi + (ew * 2 + (it & 1)) * 4096;
    const uint32_t row = smem_u32(buf) + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)lane & 7u;
    if (lane == 0) tma_store_wait_read<1>();       // the store of two tiles ago has read this staging buffer
    __syncwarp();
    const long long t0 = dbg ? clk() : 0;
    mbar_wait(&s.mma_done[acc], (uint32_t)(it >> 1) & 1u);
    if (dbg) t_mma += clk() - t0;
    tc_fence_after();
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * FUSE_COUT);
    tmem_ld_32x32(taddr, v0);
    tmem_ld_32x32(taddr + 32u, v1);
    const uint32_t res_row = patch0 + (uint32_t)(res_stage * patch_stage_bytes) + res_ofs;
    uint4 rv[4];
    if (RES) {
      mbar_wait(&s.in_full[res_stage], res_phase);   // long complete (the producers consumed the patch): orders the reads below
#pragma unroll
      for (int g = 0; g < 4; ++g) rv[g] = lds128(res_row + (((uint32_t)g ^ res_sw) << 4));
    }
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 8; ++g) {                  // 8 channels = one 16-byte chunk per step
      uint32_t pk[4];
      if (RES && g == 4) {
#pragma unroll
        for (int h = 0; h < 4; ++h) rv[h] = lds128(res_row + (((uint32_t)(4 + h) ^ res_sw) << 4));
      }
      const uint32_t rw[4] = {RES ? rv[g & 3].x : 0u, RES ? rv[g & 3].y : 0u, RES ? rv[g & 3].z : 0u, RES ? rv[g & 3].w : 0u};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = g * 8 + 2 * j;
        float y0, y1;
        bn_leaky2(__uint_as_float(c < 32 ? v0[c] : v1[c - 32]), __uint_as_float(c < 32 ? v0[c + 1] : v1[c - 31]),
                  f2_pack(a.scale2[c], a.scale2[c + 1]), f2_pack(a.shift2[c], a.shift2[c + 1]), leaky2, y0, y1);
        if (RES) { y0 += __uint_as_float(rw[j] << 16); y1 += __uint_as_float(rw[j] & 0xFFFF0000u); }
        pk[j] = pack_bf16(y0, y1);
      }
      sts128(row + (((uint32_t)g ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&s.acc_empty[acc]);              // the accumulator is in registers / staging: hand it back early
      if (RES) mbar_arrive(&s.in_empty[res_stage]);
      tma_store_4d(tmOut, buf, 0, tw.q0(), prow, tw.img);
      tma_store_commit();
    }
    if (RES && ++res_stage == IN_STAGES) { res_stage = 0; res_phase ^= 1u; }
    tw.advance();
  }
  if (lane == 0) tma_store_wait_all();
  if (dbg) {
    atomicAdd(&a.dbg[FUSE_DBG_EPI_WAIT_MMA], (unsigned long long)t_mma);
    atomicAdd(&a.dbg[FUSE_DBG_EPI_TOTAL], (unsigned long long)(clk() - t_begin));
    atomicAdd(&a.dbg[FUSE_DBG_TILES], (unsigned long long)it);
  }
}

// ---- issuer (one elected lane of warp 16): input patches + MMAs ----
// LoadIn(stage address, barrier, image, p0, q0) issues the TMA load of one input patch, PrefL2(image, p0, q0) requests a
// patch into L2 only.  The TMA loads run IN_AHEAD tiles ahead of the MMAs (IN_STAGES - 1 when only the producers release a
// stage; less when the epilogue holds it for the residual), the L2 requests FUSE_L2_AHEAD tiles.
constexpr int FUSE_L2_AHEAD = 8;
template <bool STRIDE2, int A_BYTES, int IN_STAGES, int IN_AHEAD, int IN_TX, class LoadIn, class PrefL2>
__device__ __forceinline__ void fuse_issuer_role(const FuseSmem& s, uint32_t tmem_base, const FuseArgs& a, uint32_t patch0,
                                                 int patch_stage_bytes, LoadIn load_in, PrefL2 pref_l2) {
  if (!elect_one()) return;
  const bool dbg_mma = a.dbg != nullptr;
  long long t_issue[2] = {0, 0};
  const long long t_begin = dbg_mma ? clk() : 0;
  TileWalk pf, l2;
  pf.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
  l2 = pf;
  int pf_tile = blockIdx.x, pf_stage = 0, l2_tile = blockIdx.x;
  uint32_t pf_phase = 0;                          // parity of the in_empty completion to wait for (from the second round on)
  bool pf_wrapped = false;
  auto prefetch_one = [&]() {
    if (pf_tile < a.n_tiles) {
      if (pf_wrapped) mbar_wait(&s.in_empty[pf_stage], pf_phase);
      mbar_expect_tx(&s.in_full[pf_stage], (uint32_t)IN_TX);
      load_in(patch0 + (uint32_t)(pf_stage * patch_stage_bytes), &s.in_full[pf_stage], pf.img, pf.p0(), pf.q0());
      pf.advance();
      pf_tile += gridDim.x;
      if (++pf_stage == IN_STAGES) { pf_stage = 0; if (pf_wrapped) pf_phase ^= 1u; pf_wrapped = true; }
    }
  };
  auto l2_one = [&]() {
    if (l2_tile < a.n_tiles) {
      pref_l2(l2.img, l2.p0(), l2.q0());
      l2.advance();
      l2_tile += gridDim.x;
    }
  };
  for (int i = 0; i < FUSE_L2_AHEAD; ++i) l2_one();
  for (int i = 0; i < IN_AHEAD; ++i) prefetch_one();
  int it = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
    l2_one();
    prefetch_one();
    fuse_issue_tile<STRIDE2, A_BYTES>(s, tmem_base, it, dbg_mma ? t_issue : nullptr);
  }
  if (dbg_mma) {
    atomicAdd(&a.dbg[FUSE_DBG_MMA_WAIT_A], (unsigned long long)t_issue[0]);
    atomicAdd(&a.dbg[FUSE_DBG_MMA_WAIT_ACC], (unsigned long long)t_issue[1]);
    atomicAdd(&a.dbg[FUSE_DBG_MMA_TOTAL], (unsigned long long)(clk() - t_begin));
  }
}

// ================================================================================================
// STEM: conv 3x3 s1 3->32 (mma.sync m16n8k16, K = 27 padded to 32) feeding conv 3x3 s2 32->64 (tcgen05)
// ================================================================================================
// Producer pixels of a tile, relative to (2*p0 - 1, 2*q0 - 1): yy in [0,17), xx in [0,33).  An mma tile is 16 pixels of
// one row with the same column parity, xx = par + 2i (i = 0..15): all of them land in consecutive operand rows
// (conflict-free stores) and read their im2col pairs with the same alignment.  34 such tiles + the column xx = 32
// (rows 0..15: tile C; row 16: the single pixel L) = 36 tiles = 3 per producer warp:
//   warp w: par = w & 1, rows yy = (w >> 1) + 6k, k = 0..2; the slot yy = 17 (w = 10, 11; k = 2) holds C (w = 10) and L (w = 11)
// im2col row of a pixel: K index k = 9*dy + e, e = 3*dx + ci in [0,9) -- nine consecutive patch elements of row yy + dy
// starting at element 3*(xx + 2) (the patch starts at image column 2*q0 - 4).  A thread's mma fragment slots are the
// pairs k0 = 2t, 8+2t, 16+2t, 24+2t: a pair (k0, k0+1) inside one patch row is ONE 32-bit load from patch copy
// (xx + e) & 1 (copy 1 holds the patch shifted by one element, so every pair is 4-byte aligned in one of the two);
// the slot 8+2t crosses rows for t = 0 and is fetched as two 16-bit loads; K indices >= 27 meet zero weights.
__device__ __forceinline__ int stem_elem(int k) { return (k / 9) * STEM_PATCH_PITCH + 6 + (k % 9); }

template <bool U8>
__global__ void __launch_bounds__(FUSE_THREADS, 1)
stem_fused_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const FuseArgs a) {
  extern __shared__ __align__(1024) uint8_t fuse_smem_raw[];
  uint8_t* smem = fuse_smem_raw + ((1024u - (smem_u32(fuse_smem_raw) & 1023u)) & 1023u);
  const FuseSmem s = fuse_carve<STEM_A_BYTES>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int RAW_STAGE = U8 ? STEM_RAW_STAGE_U8 : STEM_RAW_STAGE_F32;
  constexpr int RAW_TX = U8 ? STEM_RAW_TX_U8 : STEM_RAW_TX_F32;
  griddep_launch_dependents();
  const uint32_t tmem_base = fuse_prologue(s, &tmIn, &tmB, &tmOut, a, STEM_IN_STAGES, FUSE_PRODUCER_WARPS);
  const uint32_t patch_u32 = smem_u32(s.in);                                   // 2 x (copy 0, copy 1)
  const uint32_t raw_u32 = patch_u32 + 2u * STEM_PBUF_BYTES;                   // STEM_IN_STAGES raw stages
  if (!U8) {                                         // the element one past a float stage is read (and never used): keep it finite
    for (int i = threadIdx.x; i < STEM_IN_STAGES * 16; i += FUSE_THREADS)
      *reinterpret_cast<uint32_t*>(s.in + 2 * STEM_PBUF_BYTES + (i >> 4) * RAW_STAGE + RAW_TX + (i & 15) * 4) = 0u;
    __syncthreads();
  }

  if (warp < FUSE_PRODUCER_WARPS) {
    const int g = lane >> 2, t = lane & 3;
    const int par = warp & 1, e6 = warp >> 1;
    const bool special = e6 == 5;                    // third tile: C (par 0) / L (par 1)
    // ---- constants: B fragments of the first conv (permuted columns: column g of n-tile nt holds channel
    // (g/2)*8 + 2*nt + (g&1), so a thread's accumulators are the 8 consecutive channels 8t..8t+7), scale/shift pairs ----
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
    uint64_t scp[4], shp[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      scp[nt] = f2_pack(a.scale1[8 * t + 2 * nt], a.scale1[8 * t + 2 * nt + 1]);
      shp[nt] = f2_pack(a.shift1[8 * t + 2 * nt], a.shift1[8 * t + 2 * nt + 1]);
    }
    const bool leaky1 = a.leaky1 != 0;
    // ---- fragment slot addresses (bytes inside one patch buffer) of pixel (yy = e6, xx = par + 2g) ----
    const int pix = 2 * (e6 * STEM_PATCH_PITCH + 3 * (par + 2 * g));
    auto pair_ofs = [&](int k0) {                    // aligned 32-bit load of elements (k0, k0 + 1)
      const int E = stem_elem(k0);
      return ((par + k0 % 9) & 1) ? STEM_PCOPY_BYTES + 2 * (E - 1) : 2 * E;
    };
    const int s0 = pix + pair_ofs(2 * t);
    const int s2 = pix + pair_ofs(16 + 2 * t);
    const int s3 = pix + pair_ofs(t < 2 ? 24 + 2 * t : 0);          // K 28..31: zero weights, any finite data
    const int s1lo = pix + 2 * stem_elem(8 + 2 * t), s1hi = pix + 2 * stem_elem(9 + 2 * t);
    // ---- operand store offsets (bytes inside one operand buffer) ----
    const int rb0 = ((e6 & 1) ? STEM_ODD_ROW0 : 0) + (e6 >> 1);
    const int d1 = (par ? STEM_COPY_BYTES : 0) + rb0 * FUSE_ROW_BYTES + g * 64 + ((t ^ ((g >> 1) & 3)) << 4);         // kw = par, c = i
    const int d2 = 2 * STEM_COPY_BYTES + rb0 * FUSE_ROW_BYTES + (g - 1) * 64 + ((t ^ (((g - 1) >> 1) & 3)) << 4);    // kw = 2, c = i - 1
    const bool st2_r0 = par == 0 && g != 0, st2_r1 = par == 0;
    // special tiles (column xx = 32 -> copy kw = 2, c = 15): C: pixel (yy = g + 8r, 32); L: pixel (16, 32)
    const int dC_gather = (2 * STEM_PATCH_PITCH - 12) * g + 2 * 96 - 2 * 5 * STEM_PATCH_PITCH;
    const int dC_store = 2 * STEM_COPY_BYTES + (((g & 1) ? STEM_ODD_ROW0 : 0) + (g >> 1)) * FUSE_ROW_BYTES + 15 * 64 + ((t ^ 3) << 4);
    const int dL_store = 2 * STEM_COPY_BYTES + 8 * FUSE_ROW_BYTES + 15 * 64 + ((t ^ 3) << 4);
    // ---- raw -> bf16 conversion: patch words e = threadIdx.x, + 384 (19 rows x 28 words of 4 elements) ----
    int cv_raw[2], cv_dst[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = threadIdx.x + i * FUSE_PRODUCER_THREADS;
      const int rr = e / 28, wq = e - rr * 28;
      cv_raw[i] = e < STEM_PATCH_WORDS ? (U8 ? rr * STEM_RAW_ROW_U8 + 4 + 4 * wq : rr * STEM_RAW_ROW_F32 + 16 * wq) : -1;
      cv_dst[i] = 8 * e;
    }
    auto convert = [&](uint32_t raw, uint32_t dst) {
      constexpr float k255 = 0.003921568859368562698f;   // same conversion as conv_first_mma_kernel (bit-identical bf16)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (cv_raw[i] >= 0) {
          float f[5];
          if (U8) {
            const uint32_t w0 = lds32(raw + cv_raw[i]), w1 = lds32(raw + cv_raw[i] + 4);
            f[0] = (float)(w0 & 0xFFu) * k255; f[1] = (float)((w0 >> 8) & 0xFFu) * k255;
            f[2] = (float)((w0 >> 16) & 0xFFu) * k255; f[3] = (float)(w0 >> 24) * k255;
            f[4] = (float)(w1 & 0xFFu) * k255;
          } else {
            const uint4 v = lds128(raw + cv_raw[i]);
            f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
            f[4] = __uint_as_float(lds32(raw + cv_raw[i] + 16));
          }
          sts64(dst + cv_dst[i], pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
          sts64(dst + STEM_PCOPY_BYTES + cv_dst[i], pack_bf16(f[1], f[2]), pack_bf16(f[3], f[4]));
        }
      }
    };
    // mma.sync + BN + leaky of one 16-pixel tile: pk[r] = channels 8t..8t+7 of pixel row g + 8r
    auto tile_math = [&](const uint32_t (&afrag)[2][4], uint32_t (&pk)[2][4]) {
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f; }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], afrag[ks], bf[ks][nt][0], bf[ks][nt][1]);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float y0, y1;
          bn_leaky2(acc[nt][2 * r], acc[nt][2 * r + 1], scp[nt], shp[nt], leaky1, y0, y1);
          pk[r][nt] = pack_bf16(y0, y1);
        }
    };
    const bool dbg = a.dbg != nullptr && threadIdx.x == 0;
    long long t_stage = 0, t_wait = 0, t_conv = 0;
    const long long t_begin = dbg ? clk() : 0;
    TileWalk tw;
    tw.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
    int it = 0, in_stage = 0;
    uint32_t in_phase = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int p0 = tw.p0(), q0 = tw.q0();
      long long t0 = dbg ? clk() : 0;
      const uint32_t pbuf = patch_u32 + (uint32_t)(buf * STEM_PBUF_BYTES);
      mbar_wait(&s.in_full[in_stage], in_phase);
      convert(raw_u32 + (uint32_t)(in_stage * RAW_STAGE), pbuf);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.in_empty[in_stage]);
      if (++in_stage == STEM_IN_STAGES) { in_stage = 0; in_phase ^= 1u; }
      producer_bar();       // patch[buf] complete; also: every producer is past its reads of patch[buf] two tiles ago
      if (dbg) { const long long t1 = clk(); t_stage += t1 - t0; t0 = t1; }
      const uint32_t a_base = smem_u32(s.a0) + (uint32_t)(buf * STEM_A_BYTES);
      const uint32_t g0 = pbuf + (uint32_t)s0, g2 = pbuf + (uint32_t)s2, g3 = pbuf + (uint32_t)s3;
      const uint32_t g1lo = pbuf + (uint32_t)s1lo, g1hi = pbuf + (uint32_t)s1hi;
      const bool top = p0 == 0, left = q0 == 0;
      bool a_free = it < 2;                          // the MMAs of tile it-2 have read A[buf]: awaited before the first store
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uint32_t afrag[2][4], pk[2][4];
        if (k < 2 || !special) {
          constexpr int ROW6 = 6 * STEM_PATCH_PITCH * 2;                // six patch rows down: bytes
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t o = (uint32_t)(k * ROW6 + r * 96);           // pixel i + 8: 16 columns = 48 elements further
            afrag[0][r] = lds32(g0 + o);
            afrag[0][r + 2] = lds16(g1lo + o) | (lds16(g1hi + o) << 16);
            afrag[1][r] = lds32(g2 + o);
            afrag[1][r + 2] = lds32(g3 + o);
          }
          tile_math(afrag, pk);
          if (k == 0 && top && e6 == 0) {            // producer row -1 of the image: the consumer's zero padding
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) pk[0][nt] = pk[1][nt] = 0u;
          }
          if (left && par == 0 && g == 0) {          // producer column -1
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) pk[0][nt] = 0u;
          }
          if (!a_free) {
            const long long tw0 = dbg ? clk() : 0;
            mbar_wait(&s.mma_done[buf], (uint32_t)((it >> 1) - 1) & 1u);
            if (dbg) t_wait += clk() - tw0;
            a_free = true;
          }
          const uint32_t o1 = a_base + (uint32_t)(d1 + k * 3 * FUSE_ROW_BYTES), o2 = a_base + (uint32_t)(d2 + k * 3 * FUSE_ROW_BYTES);
          sts128(o1, pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
          sts128(o1 + 512u, pk[1][0], pk[1][1], pk[1][2], pk[1][3]);
          if (st2_r0) sts128(o2, pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
          if (st2_r1) sts128(o2 + 512u, pk[1][0], pk[1][1], pk[1][2], pk[1][3]);
        } else if (par == 0) {                       // tile C: pixels (yy = g + 8r, xx = 32)
          constexpr int ROW8 = 8 * STEM_PATCH_PITCH * 2;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t o = (uint32_t)(dC_gather + r * ROW8);
            afrag[0][r] = lds32(g0 + o);
            afrag[0][r + 2] = lds16(g1lo + o) | (lds16(g1hi + o) << 16);
            afrag[1][r] = lds32(g2 + o);
            afrag[1][r + 2] = lds32(g3 + o);
          }
          tile_math(afrag, pk);
          if (top && g == 0) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) pk[0][nt] = 0u;
          }
          sts128(a_base + (uint32_t)dC_store, pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
          sts128(a_base + (uint32_t)(dC_store + 4 * FUSE_ROW_BYTES), pk[1][0], pk[1][1], pk[1][2], pk[1][3]);
        } else {                                     // tile L: the single pixel (16, 32) in fragment row 0 of every lane group
          // stem_elem() of this thread's K indices 2t, 8+2t, 9+2t, 16+2t, 24+2t (28.. -> 0), written out: no division per tile
          const int e0 = 6 + 2 * t, e1lo = t == 0 ? 14 : 117 + 2 * t, e1hi = 118 + 2 * t, e2 = t == 0 ? 125 : 228 + 2 * t, e3 = t < 2 ? 236 + 2 * t : 6;
          const uint32_t lp = pbuf + (uint32_t)(2 * (16 * STEM_PATCH_PITCH + 96));
          uint32_t v[8];
          v[0] = lds16(lp + 2 * e0); v[1] = lds16(lp + 2 * e0 + 2);
          v[2] = lds16(lp + 2 * e1lo); v[3] = lds16(lp + 2 * e1hi);
          v[4] = lds16(lp + 2 * e2); v[5] = lds16(lp + 2 * e2 + 2);
          v[6] = lds16(lp + 2 * e3); v[7] = lds16(lp + 2 * e3 + 2);
          afrag[0][0] = v[0] | (v[1] << 16); afrag[0][2] = v[2] | (v[3] << 16);
          afrag[1][0] = v[4] | (v[5] << 16); afrag[1][2] = v[6] | (v[7] << 16);
          afrag[0][1] = afrag[0][3] = afrag[1][1] = afrag[1][3] = 0u;
          tile_math(afrag, pk);
          if (g == 0) sts128(a_base + (uint32_t)dL_store, pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
        }
      }
      fence_proxy_async();             // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_full[buf]);
      if (dbg) t_conv += clk() - t0;
      tw.advance();
    }
    if (dbg) {
      atomicAdd(&a.dbg[FUSE_DBG_PROD_STAGE], (unsigned long long)t_stage);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_WAIT_A], (unsigned long long)t_wait);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_CONV], (unsigned long long)(t_conv - t_wait));
      atomicAdd(&a.dbg[FUSE_DBG_PROD_TOTAL], (unsigned long long)(clk() - t_begin));
    }
  } else if (warp < FUSE_MMA_WARP) {
    fuse_epilogue_role<false, STEM_IN_STAGES, FUSE_EPI_WARP0>(s, &tmOut, tmem_base, a, 0u, 0);
  } else {
    // input patch of a tile: 19 rows from image row 2*p0 - 2, 112 elements from element (2*q0 - 4) * 3 (uint8: 4 more on
    // the left, so the box starts on a 16-byte boundary)
    auto load_in = [&](uint32_t dst, uint64_t* bar, int img, int p0, int q0) {
      tma_load_3d(&tmIn, bar, dst, (2 * q0 - 4) * 3 - (U8 ? 4 : 0), 2 * p0 - 2, img);
    };
    auto pref_l2 = [&](int img, int p0, int q0) { tma_prefetch_l2_3d(&tmIn, (2 * q0 - 4) * 3 - (U8 ? 4 : 0), 2 * p0 - 2, img); };
    fuse_issuer_role<true, STEM_A_BYTES, STEM_IN_STAGES, STEM_IN_STAGES - 1, RAW_TX>(s, tmem_base, a, raw_u32, RAW_STAGE, load_in, pref_l2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FUSE_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<2 * FUSE_COUT>(tmem_base);
  }
}

// ================================================================================================
// BLOCK: conv 1x1 64->32 feeding conv 3x3 s1 32->64 + residual -- both on tcgen05
// ================================================================================================
// The residual block of net/v3.py:16-19 at 64 channels: X -> conv1x1(32) -> conv3x3(64) -> + X.  The producer needs the
// 10 x 18 input pixels around an 8 x 16 output tile: one TMA box (zero-filled outside the image), 128-byte-swizzled --
// which IS a K-major SWIZZLE_128B UMMA operand of 180 rows (pixel f = py*18 + px) x K = 64.  The 1x1 conv is therefore
// two M = 128, N = 32 accumulations of four K = 16 MMAs each (rows 180..255 of the second read whatever follows the
// patch in shared memory; their results are never used), D1 in TMEM.  Warp-level mma.sync is not used here: it shares
// the tensor pipe with tcgen05.mma and stalls the issuing warps while the consumer's MMAs run (ncu: the HMMAs waited
// on the pipe for a fifth of the producers' time), and it needs the operands in registers.
//   warps 0-5   producer epilogue: TMEM lane quarter = warp & 3, M tile = warp >> 2 (pixels f = 128*mt + 32*q + lane < 180):
//               D1 -> BN + leaky (zero outside the image: the 3x3 conv's padding) -> bf16 -> the three kw copies of A[tile&1]
//   warp 6      one elected lane: L2 prefetch, patch loads, the 8 producer MMAs of tile i+1, then the 18 consumer MMAs of tile i
//   warps 8-23  consumer epilogue, four warps per TMEM lane quarter with 16 channels each (the epilogue is latency-bound --
//               dependent chains at one instruction per ~9 cycles -- so it is spread over many warps, scale / shift in
//               registers): BN + leaky + residual (= centre of the patch, read from shared memory), one TMA store of
//               [2 rows][16 cols][16 channels] per warp and tile
// A patch stage is released by the sixteen consumer-epilogue warps (the last readers).
constexpr int BLK_PROD_WARPS = 6;
constexpr int BLK_MMA_WARP = 6;                                           // (warp 7 idles: the epilogue warps must start at a multiple of 4)
constexpr int BLK_EPI_WARP0 = 8, BLK_EPI_WARPS = 16;
constexpr int BLK_EPI_COLS = FUSE_COUT / (BLK_EPI_WARPS / 4);             // 16
constexpr int BLK_THREADS = (BLK_EPI_WARP0 + BLK_EPI_WARPS) * 32;         // 768
constexpr int BLK_W1_BYTES = FUSE_CMID * 128;                             // [32 couts][64 cin] bf16, SWIZZLE_128B: 4,096
constexpr int BLK_TMEM_COLS = 256;                                        // [0,128): two consumer accumulators; [128,256): two x (2 M tiles x 32)
constexpr int FUSE_SMEM_BLOCK = 1024 + FUSE_HEADER + FUSE_B_BYTES + BLK_W1_BYTES + 2 * BLOCK_A_BYTES + BLOCK_IN_STAGES * BLOCK_PATCH_STAGE + FUSE_EPI_BYTES;

struct BlockArgs {
  FuseArgs f;
  float scale1[FUSE_CMID], shift1[FUSE_CMID];      // producer BN by value
};

__global__ void __launch_bounds__(BLK_THREADS, 1)
block_fused_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut, const BlockArgs ba) {
  extern __shared__ __align__(1024) uint8_t fuse_smem_raw[];
  uint8_t* smem = fuse_smem_raw + ((1024u - (smem_u32(fuse_smem_raw) & 1023u)) & 1023u);
  const FuseArgs& a = ba.f;
  // header: the common barriers, then d1_full[2], d1_empty[2]
  FuseSmem s;
  s.a_full = reinterpret_cast<uint64_t*>(smem);
  s.mma_done = s.a_full + 2;
  s.acc_empty = s.mma_done + 2;
  s.b_full = s.acc_empty + 2;
  s.in_full = s.b_full + 1;
  s.in_empty = s.in_full + FUSE_MAX_IN_STAGES;
  uint64_t* d1_full = s.in_empty + FUSE_MAX_IN_STAGES;            // 15 + 4 barriers: bytes [0, 152)
  uint64_t* d1_empty = d1_full + 2;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(smem + 256);
  s.bs = smem + FUSE_HEADER;
  uint8_t* w1s = s.bs + FUSE_B_BYTES;
  s.a0 = w1s + BLK_W1_BYTES;
  s.in = s.a0 + 2 * BLOCK_A_BYTES;                                // 1024-aligned: 1024 + 36864 + 4096 + 61440
  s.epi = s.in + BLOCK_IN_STAGES * BLOCK_PATCH_STAGE;             // after the patches: the second M tile of the last stage reads into it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.a_full[i], BLK_PROD_WARPS);
      mbar_init(&s.mma_done[i], 1);
      mbar_init(&s.acc_empty[i], BLK_EPI_WARPS);
      mbar_init(&d1_full[i], 1);
      mbar_init(&d1_empty[i], BLK_PROD_WARPS);
    }
    mbar_init(s.b_full, 1);
    for (int i = 0; i < BLOCK_IN_STAGES; ++i) {
      mbar_init(&s.in_full[i], 1);
      mbar_init(&s.in_empty[i], BLK_EPI_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == BLK_MMA_WARP) {
    tmem_alloc<BLK_TMEM_COLS>(s.tmem_ptr);
    if (elect_one()) {
      mbar_expect_tx(s.b_full, (uint32_t)(FUSE_B_BYTES + BLK_W1_BYTES));
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(&tmB, s.b_full, s.bs + tap * FUSE_BTAP_BYTES, tap * FUSE_CMID, 0);
      tma_load_2d(&tmW1, s.b_full, w1s, 0, 0);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;
  const uint32_t patch_u32 = smem_u32(s.in);

  if (warp < BLK_PROD_WARPS) {
    // ---- producer epilogue: this thread owns patch pixel f of every tile ----
    const int q = warp & 3, mt = warp >> 2;
    const int f = 128 * mt + 32 * q + lane;
    const bool valid = f < BLOCK_PATCH_PIX;
    const int py = f / BLOCK_PATCH_W, px = f - py * BLOCK_PATCH_W;
    uint32_t dofs[3];                               // byte offset of this pixel's operand row in copy kw (column c = px - kw), chunk 0
    uint32_t dsw[3];                                // its SWIZZLE_64B chunk permutation: chunk j sits at j ^ dsw
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int c = px - kw;
      const bool v = valid && c >= 0 && c < FUSE_TW;
      const int row = py * FUSE_TW + c;
      dofs[kw] = v ? (uint32_t)(kw * BLOCK_COPY_BYTES + row * 64) : 0xFFFFFFFFu;
      dsw[kw] = (uint32_t)((row >> 1) & 3);
    }
    const bool leaky1 = a.leaky1 != 0;
    const bool dbg = a.dbg != nullptr && threadIdx.x == 0;
    long long t_stage = 0, t_wait = 0, t_conv = 0;
    const long long t_begin = dbg ? clk() : 0;
    TileWalk tw;
    tw.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int p0 = tw.p0(), q0 = tw.q0();
      long long t0 = dbg ? clk() : 0;
      mbar_wait(&d1_full[buf], (uint32_t)(it >> 1) & 1u);
      if (dbg) { const long long t1 = clk(); t_stage += t1 - t0; t0 = t1; }
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(128 + buf * 64 + mt * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d1_empty[buf]);
      bool inside = valid;
      if (p0 == 0 || q0 == 0 || p0 + FUSE_TH == a.H || q0 + FUSE_TW == a.W) {
        const int y = p0 - 1 + py, x = q0 - 1 + px;
        inside = inside && (unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W;      // else: the 3x3 conv's zero padding
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float y0, y1;
        bn_leaky2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), f2_pack(ba.scale1[2 * j], ba.scale1[2 * j + 1]),
                  f2_pack(ba.shift1[2 * j], ba.shift1[2 * j + 1]), leaky1, y0, y1);
        pk[j] = inside ? pack_bf16(y0, y1) : 0u;
      }
      if (it >= 2) {
        const long long tw0 = dbg ? clk() : 0;
        mbar_wait(&s.mma_done[buf], (uint32_t)((it >> 1) - 1) & 1u);       // the MMAs of tile it-2 have read A[buf]
        if (dbg) t_wait += clk() - tw0;
      }
      const uint32_t a_base = smem_u32(s.a0) + (uint32_t)(buf * BLOCK_A_BYTES);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        if (dofs[kw] != 0xFFFFFFFFu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) sts128(a_base + dofs[kw] + (((uint32_t)j ^ dsw[kw]) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_full[buf]);
      if (dbg) t_conv += clk() - t0;
      tw.advance();
    }
    if (dbg) {
      atomicAdd(&a.dbg[FUSE_DBG_PROD_STAGE], (unsigned long long)t_stage);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_WAIT_A], (unsigned long long)t_wait);
      atomicAdd(&a.dbg[FUSE_DBG_PROD_CONV], (unsigned long long)(t_conv - t_wait));
      atomicAdd(&a.dbg[FUSE_DBG_PROD_TOTAL], (unsigned long long)(clk() - t_begin));
    }
  } else if (warp >= BLK_EPI_WARP0) {
    // ---- consumer epilogue: TMEM lane quarter = warp & 3 (tile rows 2q, 2q+1), channels [16 * part, 16 * part + 16) ----
    const int quarter = warp & 3, ew = warp - BLK_EPI_WARP0, part = ew >> 2;
    const bool dbg = a.dbg != nullptr && ew == 1 && lane == 0;
    long long t_mma = 0;
    const long long t_begin = dbg ? clk() : 0;
    // residual position of this lane in the input patch: output pixel (2*quarter + lane/16, lane%16) -> patch pixel (+1, +1)
    const int res_f = (2 * quarter + (lane >> 4) + 1) * BLOCK_PATCH_W + (lane & 15) + 1;
    const uint32_t res_ofs = (uint32_t)res_f * 128u, res_sw = (uint32_t)res_f & 7u;
    int res_stage = 0;
    uint32_t res_phase = 0;
    const bool leaky2 = a.leaky2 != 0;
    uint64_t sc[BLK_EPI_COLS / 2], sh[BLK_EPI_COLS / 2];
#pragma unroll
    for (int j = 0; j < BLK_EPI_COLS / 2; ++j) {
      sc[j] = f2_pack(a.scale2[part * BLK_EPI_COLS + 2 * j], a.scale2[part * BLK_EPI_COLS + 2 * j + 1]);
      sh[j] = f2_pack(a.shift2[part * BLK_EPI_COLS + 2 * j], a.shift2[part * BLK_EPI_COLS + 2 * j + 1]);
    }
    TileWalk tw;
    tw.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      // staging: 32 pixels (this warp's two tile rows) x 32 bytes (16 channels), dense
      uint8_t* buf = s.epi + (ew * 2 + (it & 1)) * 1024;
      const uint32_t row = smem_u32(buf) + (uint32_t)lane * 32u;
      if (lane == 0) tma_store_wait_read<1>();       // the store of two tiles ago has read this staging buffer
      __syncwarp();
      const long long t0 = dbg ? clk() : 0;
      mbar_wait(&s.mma_done[acc], (uint32_t)(it >> 1) & 1u);
      if (dbg) t_mma += clk() - t0;
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * FUSE_COUT + part * BLK_EPI_COLS), v);
      const uint32_t res_row = patch_u32 + (uint32_t)(res_stage * BLOCK_PATCH_STAGE) + res_ofs;
      mbar_wait(&s.in_full[res_stage], res_phase);   // long complete (the MMAs consumed the patch): orders the reads below
      const uint4 r0 = lds128(res_row + (((uint32_t)(2 * part) ^ res_sw) << 4));
      const uint4 r1 = lds128(res_row + (((uint32_t)(2 * part + 1) ^ res_sw) << 4));
      tmem_ld_wait();
      const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y0, y1;
        bn_leaky2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), sc[j], sh[j], leaky2, y0, y1);
        y0 += __uint_as_float(rw[j] << 16);
        y1 += __uint_as_float(rw[j] & 0xFFFF0000u);
        pk[j] = pack_bf16(y0, y1);
      }
      sts128(row, pk[0], pk[1], pk[2], pk[3]);
      sts128(row + 16u, pk[4], pk[5], pk[6], pk[7]);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s.acc_empty[acc]);              // the accumulator is in registers / staging: hand it back early
        mbar_arrive(&s.in_empty[res_stage]);
        tma_store_4d(&tmOut, buf, part * BLK_EPI_COLS, tw.q0(), tw.p0() + 2 * quarter, tw.img);
        tma_store_commit();
      }
      if (++res_stage == BLOCK_IN_STAGES) { res_stage = 0; res_phase ^= 1u; }
      tw.advance();
    }
    if (lane == 0) tma_store_wait_all();
    if (dbg) {
      atomicAdd(&a.dbg[FUSE_DBG_EPI_WAIT_MMA], (unsigned long long)t_mma);
      atomicAdd(&a.dbg[FUSE_DBG_EPI_TOTAL], (unsigned long long)(clk() - t_begin));
      atomicAdd(&a.dbg[FUSE_DBG_TILES], (unsigned long long)it);
    }
  } else if (warp == BLK_MMA_WARP) {
    if (elect_one()) {
      const bool dbg_mma = a.dbg != nullptr;
      long long t_issue[2] = {0, 0};
      const long long t_begin = dbg_mma ? clk() : 0;
      TileWalk pf, l2;
      pf.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
      l2 = pf;
      int pf_tile = blockIdx.x, pf_stage = 0, l2_tile = blockIdx.x;
      uint32_t pf_phase = 0;
      bool pf_wrapped = false;
      auto prefetch_one = [&]() {
        if (pf_tile < a.n_tiles) {
          if (pf_wrapped) mbar_wait(&s.in_empty[pf_stage], pf_phase);
          mbar_expect_tx(&s.in_full[pf_stage], (uint32_t)BLOCK_PATCH_TX);
          tma_load_4d(&tmIn, &s.in_full[pf_stage], patch_u32 + (uint32_t)(pf_stage * BLOCK_PATCH_STAGE), 0, pf.q0() - 1, pf.p0() - 1, pf.img);
          pf.advance();
          pf_tile += gridDim.x;
          if (++pf_stage == BLOCK_IN_STAGES) { pf_stage = 0; if (pf_wrapped) pf_phase ^= 1u; pf_wrapped = true; }
        }
      };
      auto l2_one = [&]() {
        if (l2_tile < a.n_tiles) {
          tma_prefetch_l2_4d(&tmIn, 0, l2.q0() - 1, l2.p0() - 1, l2.img);
          l2.advance();
          l2_tile += gridDim.x;
        }
      };
      // producer conv of tile j: D1[j&1] = patch(stage j % S) x W1^T, two M tiles
      int l3_stage = 0;
      uint32_t l3_phase = 0;
      const uint32_t w1_addr = smem_u32(w1s);
      auto issue_l3 = [&](int j) {
        constexpr uint32_t idesc = make_idesc<FUSE_CMID>();
        mbar_wait(&s.in_full[l3_stage], l3_phase);
        if (j >= 2) mbar_wait(&d1_empty[j & 1], (uint32_t)((j >> 1) - 1) & 1u);
        tc_fence_after();
        const uint32_t pa = patch_u32 + (uint32_t)(l3_stage * BLOCK_PATCH_STAGE);
        const uint64_t db = make_kmajor_desc<64>(w1_addr);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t da = make_kmajor_desc<64>(pa + (uint32_t)(mt * 128 * 128));
          const uint32_t td = tmem_base + (uint32_t)(128 + (j & 1) * 64 + mt * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(td, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&d1_full[j & 1]);
        if (++l3_stage == BLOCK_IN_STAGES) { l3_stage = 0; l3_phase ^= 1u; }
      };
      for (int i = 0; i < FUSE_L2_AHEAD; ++i) l2_one();
      for (int i = 0; i < BLOCK_IN_STAGES - 2; ++i) prefetch_one();
      mbar_wait(s.b_full, 0);
      if ((int)blockIdx.x < a.n_tiles) issue_l3(0);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        l2_one();
        prefetch_one();
        if (tile + (int)gridDim.x < a.n_tiles) issue_l3(it + 1);
        fuse_issue_tile<false, BLOCK_A_BYTES>(s, tmem_base, it, dbg_mma ? t_issue : nullptr);
      }
      if (dbg_mma) {
        atomicAdd(&a.dbg[FUSE_DBG_MMA_WAIT_A], (unsigned long long)t_issue[0]);
        atomicAdd(&a.dbg[FUSE_DBG_MMA_WAIT_ACC], (unsigned long long)t_issue[1]);
        atomicAdd(&a.dbg[FUSE_DBG_MMA_TOTAL], (unsigned long long)(clk() - t_begin));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BLK_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<BLK_TMEM_COLS>(tmem_base);
  }
}

// ================================================================================================
// CONVPOOL: conv 3x3 s1 32->64 + BN + leaky + 2x2/2 max pool (Darknet-19's second conv, net/v2.py:23-24,
// net/layers.py:70-81) -- the consumer of the fused block without a producer conv, the pool taken in registers
// ================================================================================================
// The 10 x 18 x 32-channel input patch arrives as one TMA box (SWIZZLE_64B, zero outside the image = the conv's padding);
// six warps copy it into the three kw copies of the operand (no arithmetic), the 18 MMAs are those of the block kernel.
// Epilogue: a warp holds tile rows 2q (lanes 0-15) and 2q+1 (lanes 16-31), one pixel per lane, so the 2x2 window of pooled
// pixel (q, c/2) is lanes {c, c^1} x {row, row^16}: two shuffles and two packed bf16 maxima per register.  max commutes with
// the monotone bf16 rounding: bit-identical to conv -> bf16 -> maxpool2_kernel.  Only the pooled 4 x 8 x 64 tile is written
// (one TMA store of 8 pixels x 16 channels per warp); the full-resolution activation never exists.  The epilogue is the
// long pole -- per warp a chain of dependent shuffles -- and it is latency-, not throughput-bound (ncu: its warps were busy
// 96% of the time at one instruction per 9 cycles with 4 or 8 warps), hence sixteen warps of 16 channels each with their
// scale / shift in registers.
constexpr int CP_PROD_WARPS = 6;
constexpr int CP_MMA_WARP = 6;                                           // (warp 7 idles: the epilogue warps must start at a multiple of 4)
constexpr int CP_EPI_WARP0 = 8, CP_EPI_WARPS = 16;                       // four epilogue warps per TMEM lane quarter: 16 columns each
constexpr int CP_EPI_COLS = FUSE_COUT / (CP_EPI_WARPS / 4);              // 16
constexpr int CP_THREADS = (CP_EPI_WARP0 + CP_EPI_WARPS) * 32;           // 768
constexpr int CP_PATCH_TX = BLOCK_PATCH_PIX * 64;                        // 11,520
constexpr int CP_PATCH_STAGE = 12288;
constexpr int CP_IN_STAGES = 4;
constexpr int CP_EPI_BYTES = CP_EPI_WARPS * 2 * 256;                     // per epilogue warp two buffers of 8 pooled pixels x 32 B
constexpr int CP_BUFS = 4;                                               // operand buffers = TMEM accumulators (64 columns each)
constexpr int FUSE_SMEM_CONVPOOL = 1024 + FUSE_HEADER + FUSE_B_BYTES + CP_BUFS * BLOCK_A_BYTES + CP_IN_STAGES * CP_PATCH_STAGE + CP_EPI_BYTES;

__global__ void __launch_bounds__(CP_THREADS, 1)
convpool_fused_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const FuseArgs a) {
  extern __shared__ __align__(1024) uint8_t fuse_smem_raw[];
  uint8_t* smem = fuse_smem_raw + ((1024u - (smem_u32(fuse_smem_raw) & 1023u)) & 1023u);
  FuseSmem s;
  s.a_full = reinterpret_cast<uint64_t*>(smem);              // [4] [4] [4] [1] [4] [4]: 21 barriers, bytes [0, 168)
  s.mma_done = s.a_full + CP_BUFS;
  s.acc_empty = s.mma_done + CP_BUFS;
  s.b_full = s.acc_empty + CP_BUFS;
  s.in_full = s.b_full + 1;
  s.in_empty = s.in_full + FUSE_MAX_IN_STAGES;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(smem + 256);
  s.bs = smem + FUSE_HEADER;
  s.a0 = s.bs + FUSE_B_BYTES;
  s.in = s.a0 + CP_BUFS * BLOCK_A_BYTES;
  s.epi = s.in + CP_IN_STAGES * CP_PATCH_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < CP_BUFS; ++i) {
      mbar_init(&s.a_full[i], CP_PROD_WARPS);
      mbar_init(&s.mma_done[i], 1);
      mbar_init(&s.acc_empty[i], CP_EPI_WARPS);
    }
    mbar_init(s.b_full, 1);
    for (int i = 0; i < CP_IN_STAGES; ++i) {
      mbar_init(&s.in_full[i], 1);
      mbar_init(&s.in_empty[i], CP_PROD_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == CP_MMA_WARP) {
    tmem_alloc<CP_BUFS * FUSE_COUT>(s.tmem_ptr);
    if (elect_one()) {
      mbar_expect_tx(s.b_full, (uint32_t)FUSE_B_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(&tmB, s.b_full, s.bs + tap * FUSE_BTAP_BYTES, tap * FUSE_CMID, 0);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;
  const uint32_t patch_u32 = smem_u32(s.in);

  if (warp < CP_PROD_WARPS) {
    // ---- copy producers: this thread owns patch pixel f of every tile ----
    const int q = warp & 3, mt = warp >> 2;
    const int f = 128 * mt + 32 * q + lane;
    const bool valid = f < BLOCK_PATCH_PIX;
    const int py = f / BLOCK_PATCH_W, px = f - py * BLOCK_PATCH_W;
    const uint32_t src_ofs = (uint32_t)(valid ? f : 0) * 64u, src_sw = (uint32_t)(f >> 1) & 3u;
    uint32_t dofs[3], dsw[3];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int c = px - kw;
      const bool v = valid && c >= 0 && c < FUSE_TW;
      const int row = py * FUSE_TW + c;
      dofs[kw] = v ? (uint32_t)(kw * BLOCK_COPY_BYTES + row * 64) : 0xFFFFFFFFu;
      dsw[kw] = (uint32_t)((row >> 1) & 3);
    }
    int it = 0, in_stage = 0;
    uint32_t in_phase = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & (CP_BUFS - 1);
      const uint32_t patch = patch_u32 + (uint32_t)(in_stage * CP_PATCH_STAGE) + src_ofs;
      mbar_wait(&s.in_full[in_stage], in_phase);
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = lds128(patch + (((uint32_t)j ^ src_sw) << 4));
      if (it >= CP_BUFS) mbar_wait(&s.mma_done[buf], (uint32_t)((it / CP_BUFS) - 1) & 1u);       // the MMAs of tile it-4 have read A[buf]
      const uint32_t a_base = smem_u32(s.a0) + (uint32_t)(buf * BLOCK_A_BYTES);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        if (dofs[kw] != 0xFFFFFFFFu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) sts128(a_base + dofs[kw] + (((uint32_t)j ^ dsw[kw]) << 4), v[j].x, v[j].y, v[j].z, v[j].w);
        }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s.in_empty[in_stage]);          // the patch values are in the operand (the stores above consumed the loads)
        mbar_arrive(&s.a_full[buf]);
      }
      if (++in_stage == CP_IN_STAGES) { in_stage = 0; in_phase ^= 1u; }
    }
  } else if (warp >= CP_EPI_WARP0) {
    // ---- pooled epilogue ----
    const int quarter = warp & 3, ew = warp - CP_EPI_WARP0, part = ew >> 2;       // columns [16 * part, 16 * part + 16)
    const bool leaky2 = a.leaky2 != 0;
    const bool writer = (lane & 17) == 0;             // even column of the even row: lanes 0, 2, ..., 14
    uint64_t sc[CP_EPI_COLS / 2], sh[CP_EPI_COLS / 2];
#pragma unroll
    for (int j = 0; j < CP_EPI_COLS / 2; ++j) {
      sc[j] = f2_pack(a.scale2[part * CP_EPI_COLS + 2 * j], a.scale2[part * CP_EPI_COLS + 2 * j + 1]);
      sh[j] = f2_pack(a.shift2[part * CP_EPI_COLS + 2 * j], a.shift2[part * CP_EPI_COLS + 2 * j + 1]);
    }
    TileWalk tw;
    tw.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & (CP_BUFS - 1);
      uint8_t* buf = s.epi + (ew * 2 + (it & 1)) * 256;
      // staging: 8 pooled pixels x 32 bytes (16 channels), no swizzle (two 16-byte chunks per row)
      const uint32_t row = smem_u32(buf) + (uint32_t)(lane >> 1) * 32u;
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      mbar_wait(&s.mma_done[acc], (uint32_t)(it / CP_BUFS) & 1u);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * FUSE_COUT + part * CP_EPI_COLS), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.acc_empty[acc]);  // the accumulator is in registers: hand it back
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y0, y1;
        bn_leaky2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), sc[j], sh[j], leaky2, y0, y1);
        pk[j] = pack_bf16(y0, y1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, pk[j], 1);
        const __nv_bfloat162 mm = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&pk[j]), *reinterpret_cast<const __nv_bfloat162*>(&o));
        pk[j] = *reinterpret_cast<const uint32_t*>(&mm);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, pk[j], 16);
        const __nv_bfloat162 mm = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&pk[j]), *reinterpret_cast<const __nv_bfloat162*>(&o));
        pk[j] = *reinterpret_cast<const uint32_t*>(&mm);
      }
      if (writer) {
        sts128(row, pk[0], pk[1], pk[2], pk[3]);
        sts128(row + 16u, pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&tmOut, buf, part * CP_EPI_COLS, tw.q0() >> 1, (tw.p0() >> 1) + quarter, tw.img);
        tma_store_commit();
      }
      tw.advance();
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp == CP_MMA_WARP) {
    if (elect_one()) {
      TileWalk pf, l2;
      pf.init(a.tiles_w, a.tiles_h, blockIdx.x, gridDim.x);
      l2 = pf;
      int pf_tile = blockIdx.x, pf_stage = 0, l2_tile = blockIdx.x;
      uint32_t pf_phase = 0;
      bool pf_wrapped = false;
      auto prefetch_one = [&]() {
        if (pf_tile < a.n_tiles) {
          if (pf_wrapped) mbar_wait(&s.in_empty[pf_stage], pf_phase);
          mbar_expect_tx(&s.in_full[pf_stage], (uint32_t)CP_PATCH_TX);
          tma_load_4d(&tmIn, &s.in_full[pf_stage], patch_u32 + (uint32_t)(pf_stage * CP_PATCH_STAGE), 0, pf.q0() - 1, pf.p0() - 1, pf.img);
          pf.advance();
          pf_tile += gridDim.x;
          if (++pf_stage == CP_IN_STAGES) { pf_stage = 0; if (pf_wrapped) pf_phase ^= 1u; pf_wrapped = true; }
        }
      };
      auto l2_one = [&]() {
        if (l2_tile < a.n_tiles) {
          tma_prefetch_l2_4d(&tmIn, 0, l2.q0() - 1, l2.p0() - 1, l2.img);
          l2.advance();
          l2_tile += gridDim.x;
        }
      };
      for (int i = 0; i < FUSE_L2_AHEAD; ++i) l2_one();
      for (int i = 0; i < CP_IN_STAGES - 1; ++i) prefetch_one();
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        l2_one();
        prefetch_one();
        {   // 18 MMAs of tile `it` into accumulator / from operand buffer it % 4
          constexpr uint32_t idesc = make_idesc<FUSE_COUT>();
          const int buf = it & (CP_BUFS - 1);
          const uint32_t use = (uint32_t)(it / CP_BUFS);
          if (it == 0) mbar_wait(s.b_full, 0);
          if (it >= CP_BUFS) mbar_wait(&s.acc_empty[buf], (use - 1u) & 1u);
          mbar_wait(&s.a_full[buf], use & 1u);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(s.a0) + (uint32_t)(buf * BLOCK_A_BYTES);
          const uint32_t b_addr = smem_u32(s.bs);
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * FUSE_COUT);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int kh = tap / 3, kw = tap - kh * 3;
            const uint64_t da = make_kmajor_desc<FUSE_CMID>(a_addr + (uint32_t)(kw * BLOCK_COPY_BYTES + kh * FUSE_ROW_BYTES));
            const uint64_t db = make_kmajor_desc<FUSE_CMID>(b_addr + (uint32_t)(tap * FUSE_BTAP_BYTES));
#pragma unroll
            for (int k = 0; k < FUSE_CMID / 16; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
          }
          umma_commit(&s.mma_done[buf]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CP_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<CP_BUFS * FUSE_COUT>(tmem_base);
  }
}

}  // namespace yb
