// conv_tc.cuh -- NHWC bf16 implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA (sm_100a only).
//
// Replaces tf.layers.conv2d + batch_normalization + leaky_relu (+ the shortcut add, the route
// concat, the x2 nearest upsample and the space-to-depth reorg that follow it) of the reference's
// net/layers.py:17-67,84-116 with ONE kernel per conv:
//
//   D[m, co] = sum_{kh,kw,ci} X[n, p*s + kh - pad, q*s + kw - pad, ci] * Wt[co, (kh,kw,ci)]      m = (n,p,q)
//   Y = act(D * scale[co] + shift[co]) (+ residual[m, co])      -> bf16 (or fp32 for the heads)
//
// GEMM view: M = N*Ho*Wo pixels (tile 128 = the 128 TMEM lanes), N = Cout (tile BN = TMEM columns),
// K = k*k*Cin walked as (tap, 64-channel block).  Per K step the producer warp issues
//   * A: one TMA load of 128 pixels x BK channels.  1x1 convs use a tiled 2-D map over [pixels, C];
//        3x3 convs use an IM2COL map over (C,W,H,N): the hardware walks 128 output pixels across
//        rows and images from the tile's first pixel, applies the tap offset and the conv stride and
//        zero-fills the padding halo -- the reference's tf.pad/"SAME" costs nothing.
//   * B: one tiled TMA load of BN x BK weights (K-major, packed [Cout_pad][k*k*Cin] at load time).
// Both land in 128B-swizzled (64B for Cin=32) K-major shared memory, exactly the canonical UMMA
// operand layout, so a single elected thread issues BK/16 tcgen05.mma (M=128, N=BN, K=16) per step
// with fp32 accumulation in TMEM.  tcgen05.commit releases smem stages / publishes the accumulator
// through mbarriers; four epilogue warps read their TMEM lane quarter with tcgen05.ld and apply the
// whole epilogue in registers before storing straight to the consumer's buffer (concat slices,
// upsample replicas and reorg scatter are just different store addresses).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace yb {

enum OutMode { OUT_PLAIN = 0, OUT_UPSAMPLE2 = 1, OUT_REORG2 = 2, OUT_POOL2 = 3 };     // OUT_POOL2: first conv (mma.sync) only

struct ConvArgs {
  int M;              // valid output pixels = n * Ho * Wo
  int Ho, Wo;         // output spatial size
  int cout;           // real output channels
  int taps;           // ksize * kw
  int ksize;          // kernel height
  int kw;             // kernel width   } differ from ksize / conv_stride / pad only for the input-pair view of a stride-2 conv
  int stride_w;       // stride along W } (engine.cu, Op::px_pair == 2)
  int pad_w;          // low-side padding along W
  int kc_blocks;      // Cin / BK
  int conv_stride;    // 1 or 2
  int pad;            // low-side padding ((k-1)/2)
  int im2col;         // 1: A through the im2col map; 0: tiled [pixels, C] map (1x1 stride 1)
  const float* scale; // [cout_pad]
  const float* shift; // [cout_pad]
  int leaky;
  const __nv_bfloat16* res;  // residual (shortcut) or nullptr; pixel stride res_ld, already channel-offset
  int res_ld;
  void* out;          // already channel-offset
  int out_ld;         // elements per output pixel
  int out_f32;        // 1: float32 store (heads), 0: bf16
  int out_mode;       // OutMode
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {   // <= N most recent store groups may still read smem
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Programmatic dependent launch: the next kernel in the stream may start its prologue (barrier init, TMEM
// allocation, weight loads) once every CTA of this grid has passed launch_dependents; it must not touch anything the
// previous grids wrote (or still read) before griddep_wait() returns.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- 2-CTA (cta_group::2) helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on a barrier of the pair's leader CTA (cluster address `mbar`)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t mbar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(const CUtensorMap* m, uint32_t mbar, void* dst, int c, int w, int h,
                                                        int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs] * B[smem of both CTAs]: M = 256 (128 rows per CTA), issued by the leader
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this smem offset in every CTA of `mask` once the leader's prior MMAs are complete
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
template <int BN, int M>
__device__ __forceinline__ constexpr uint32_t make_idesc_m() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, single CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (quarter*32 + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (= 8 rows * row bytes) | [46,48) version = 1 (sm_100)
//   [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  constexpr uint32_t row_bytes = BK * 2;
  constexpr uint64_t layout = (row_bytes == 128) ? 2ull : 4ull;
  constexpr uint64_t sbo = (8u * row_bytes) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6)=1, A bf16 [7,10)=1,
// B bf16 [10,13)=1, A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// packed fp32 pairs (sm_100): one instruction, two IEEE operations
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// y = acc * scale + shift; leaky: max(y, 0.1 y) -- the scalar epilogue's operations (FFMA, FMUL, FMNMX), two lanes at a time
__device__ __forceinline__ void bn_leaky2(float a0, float a1, uint64_t sc, uint64_t sh, bool leaky, float& y0, float& y1) {
  const uint64_t y = f2_fma(f2_pack(a0, a1), sc, sh);
  f2_unpack(y, y0, y1);
  if (leaky) {
    float z0, z1;
    f2_unpack(f2_mul(y, f2_pack(0.1f, 0.1f)), z0, z1);
    y0 = fmaxf(y0, z0); y1 = fmaxf(y1, z1);
  }
}

// ================================================================================================
// Persistent kernel: one CTA per SM loops over output tiles.  Two TMEM accumulator buffers let the
// epilogue of tile i (TMEM -> registers -> global) overlap the mainloop of tile i+1, the TMA producer
// runs ahead across tile boundaries, and the per-CTA fixed costs (TMEM alloc, barrier init, descriptor
// prefetch, scale/shift staging) are paid once per SM instead of once per tile.
//   warp 0        TMA producer (one elected lane)
//   warp 1        MMA issuer (one elected lane) + TMEM owner
//   warps 2..9    epilogue: TMEM lane quarter = warp % 4; the two warps of a quarter split the columns
// ================================================================================================
constexpr int CONV_TCP_THREADS = 320;
constexpr int CONV_TCP_EPI_WARPS = 8;
constexpr int CONV_TCP_MAX_STAGES = 16;
constexpr int CONV_TCP_MAX_COUT_PAD = 1024;
// dynamic shared memory: [1024 B barriers | scale,shift (8 KB) | optional stationary B | stages]
constexpr int CONV_TCP_HEADER = 1024 + 2 * CONV_TCP_MAX_COUT_PAD * 4;
constexpr int CONV_TCP_SMEM_MAX = 232448;                             // 227 KB opt-in limit
constexpr int CONV_TCP_TILE_BUDGET = CONV_TCP_SMEM_MAX - 1024 - CONV_TCP_HEADER;
// TMA epilogue staging: per epilogue warp two 32-row x 64-byte (SWIZZLE_64B) sub-tiles
constexpr int CONV_TCP_EPI_BUF = 32 * 64;
constexpr int CONV_TCP_EPI_BYTES = CONV_TCP_EPI_WARPS * 2 * CONV_TCP_EPI_BUF;

struct PersistArgs {
  int n_tiles_n, n_tiles, cout_pad;
  int n_stages;        // pipeline depth (runtime: fills the shared memory that is left)
  int ksub;            // BK-blocks per pipeline stage: one barrier round trip (and one pass of the issue loop) feeds
                       // ksub * BK/16 MMAs, so narrow tiles still give the tensor pipe >= ~500 cycles per hand-shake
  int b_stationary;    // 1: the whole [BN x K] weight matrix is loaded once per CTA and stays in shared memory
  int tma_epi;         // 1: bf16 plain output through smem + TMA store, residual through TMA load
  int solo_issue;      // 1: a single thread runs the MMA issue loop; 0: the whole warp walks it, electing per stage
  int ablate;          // debug/roofline probes (results are wrong when non-zero): 1 = epilogue does nothing,
                       // 2 = no MMAs, 4 = no A loads, 8 = no B loads
  int splitk;          // > 1: every tile is computed by `splitk` work units (consecutive units = the parts of one tile, so
                       // they run on neighbouring CTAs at the same time), each walking num_k / splitk K blocks; the raw
                       // fp32 accumulators go to `ws`, the last part to arrive (per epilogue warp, counted in `ws_cnt`)
                       // sums all parts in part order -- deterministic -- and runs the epilogue.  Latency mode for grids
                       // that fill a fraction of the SMs (small batches): fp32 summation order differs from splitk = 1.
  float* ws;           // [n_tiles * splitk][pair rank][128][BN] partial accumulators
  unsigned int* ws_cnt;// [n_tiles][pair rank][8 epilogue warps] arrival counters (zero between launches: the last part resets them)
  unsigned long long* dbg;   // optional [16] cycle counters summed over all CTAs (see CONV_DBG_*); nullptr = off
};
// cycle counters: who waits on whom inside the persistent conv kernel
enum ConvDbg {
  CONV_DBG_PROD_WAIT_EMPTY = 0, CONV_DBG_PROD_TOTAL, CONV_DBG_MMA_WAIT_FULL, CONV_DBG_MMA_WAIT_TMEM, CONV_DBG_MMA_TOTAL,
  CONV_DBG_EPI_WAIT_ACC, CONV_DBG_EPI_WAIT_RES, CONV_DBG_EPI_WAIT_BUF, CONV_DBG_EPI_TMEM_LD, CONV_DBG_EPI_MATH_STORE,
  CONV_DBG_EPI_TOTAL, CONV_DBG_EPI_FENCE_STORE, CONV_DBG_COUNT
};
__device__ __forceinline__ long long clk() { return clock64(); }

// PAIR = true: launched as 2-CTA clusters.  The pair computes a 256(M) x BN(N) tile with tcgen05.mma.cta_group::2:
// CTA r holds the A rows of M tile 2j+r and the B rows [r*BN/2, (r+1)*BN/2) of the N tile, so every SM ingests half
// of B per K step (BN=256: 32 KB instead of 48 KB) and one MMA instruction covers 256 rows -- issuing an MMA costs
// ~120 cycles whatever its shape, which is what bounds the narrow-N layers.  The leader (rank 0) issues the MMAs, both
// CTAs' TMA loads complete on the leader's full barriers, the leader's commits release the smem stages / publish the
// accumulators in both CTAs, and both CTAs' epilogue warps hand the accumulator back on the leader's tmem_empty barrier.
// SPLIT = true: the split-K instantiation (PersistArgs::splitk > 1 allowed); the plain one carries none of its code.
template <int BN, int BK, bool PAIR, bool SPLIT = false>
__global__ void __launch_bounds__(CONV_TCP_THREADS, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes, const ConvArgs a,
                       const PersistArgs pa) {
  static_assert(!PAIR || BN >= 64, "cta_group::2 needs N >= 64 here (32 B rows per CTA)");
  constexpr int A_BYTES = 128 * BK * 2;
  constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;      // bytes of B this CTA loads per K step
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles.  The offset is added to the __shared__ array itself (not to an
  // integer-cast pointer) so the compiler keeps every access in the shared address space (LDS/STS, not generic LD/ST).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + CONV_TCP_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + CONV_TCP_MAX_STAGES;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                 // [2]
  uint64_t* b_full_bar = tmem_empty_bar + 2;                    // [1]
  uint64_t* res_bar = b_full_bar + 1;                           // [EPI_WARPS][2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 2 * CONV_TCP_EPI_WARPS);
  // BN scale / shift per channel pair, interleaved {scale[c], scale[c+1], shift[c], shift[c+1]}: one 16-byte broadcast load
  // per pair in the epilogue
  float4* s_ss = reinterpret_cast<float4*>(smem + 1024);
  const int num_k = a.taps * a.kc_blocks;
  const int n_stages = pa.n_stages;
  const bool bstat = pa.b_stationary != 0;
  uint8_t* epi_smem = smem + CONV_TCP_HEADER;                                 // CONV_TCP_EPI_BYTES when tma_epi
  uint8_t* b_stat = epi_smem + (pa.tma_epi ? CONV_TCP_EPI_BYTES : 0);        // num_k * B_BYTES when stationary
  uint8_t* stages = b_stat + (bstat ? num_k * B_BYTES : 0);
  const int sub_bytes = A_BYTES + (bstat ? 0 : B_BYTES);      // one BK-block of A (and B) inside a stage
  const int ksub = pa.ksub;
  const int stage_bytes = ksub * sub_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], PAIR ? 2 : 1);            // pair: both producers arrive on the leader's barrier
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], (PAIR ? 2 : 1) * CONV_TCP_EPI_WARPS);
    }
    mbar_init(b_full_bar, PAIR ? 2 : 1);
    for (int s = 0; s < 2 * CONV_TCP_EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
    if (pa.tma_epi) {
      tma_prefetch_desc(&tmOut);
      if (a.res != nullptr) tma_prefetch_desc(&tmRes);
    }
    fence_barrier_init();
  }
  if (PAIR) cluster_sync_all();                        // peer barriers are initialised before anything remote arrives
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair<TMEM_COLS>(tmem_ptr_smem);
    else tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  }
  for (int i = threadIdx.x; i < pa.cout_pad; i += CONV_TCP_THREADS) {
    reinterpret_cast<float*>(s_ss)[(i >> 1) * 4 + (i & 1)] = a.scale[i];
    reinterpret_cast<float*>(s_ss)[(i >> 1) * 4 + 2 + (i & 1)] = a.shift[i];
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // tile walk: single CTA -> tiles blockIdx.x, +gridDim.x, ...; pair -> pair-tiles (cluster id), M tile = 2j + rank
  const int walk_start = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int walk_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // split-K: work unit = (tile, part); num_k is a multiple of splitk (engine.cu: launch_conv_tcp)
  const int splitk = (SPLIT && pa.splitk > 1) ? pa.splitk : 1;
  const int n_units = pa.n_tiles * splitk;
  const int k_per = num_k / splitk;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      if (bstat) {      // weights do not depend on the previous kernel: fetched before the grid dependency resolves
        if (PAIR) {     // each CTA keeps its half of the N tile; both halves complete on the leader's barrier
          const uint32_t bb = mapa_u32(smem_u32(b_full_bar), 0);
          if (rank == 0) mbar_expect_tx(b_full_bar, 2u * (uint32_t)(num_k * B_BYTES));
          else mbar_arrive_cluster(bb);
          for (int kb = 0; kb < num_k; ++kb) tma_load_2d_pair(&tmB, bb, b_stat + kb * B_BYTES, kb * BK, (int)rank * (BN / 2));
        } else {
          mbar_expect_tx(b_full_bar, (uint32_t)(num_k * B_BYTES));
          for (int kb = 0; kb < num_k; ++kb) tma_load_2d(&tmB, b_full_bar, b_stat + kb * B_BYTES, kb * BK, 0);
        }
      }
      griddep_wait();
      const bool load_a = !(pa.ablate & 4), load_b = !bstat && !(pa.ablate & 8);
      const uint32_t tx_bytes = (load_a ? (uint32_t)A_BYTES : 0u) + (load_b ? (uint32_t)B_BYTES : 0u);
      const bool dbg = pa.dbg != nullptr;
      long long t_wait = 0;
      const long long t_begin = dbg ? clk() : 0;
      int stage = 0;
      uint32_t phase = 0;
      const int hw = a.Ho * a.Wo;
      const uint32_t full0 = PAIR ? mapa_u32(smem_u32(full_bar), 0) : 0u;       // leader's full_bar[0]
      // KS: BK-blocks per stage as a compile-time constant (0 = runtime value, any remainder).  The producer's loop
      // latency bounds the TMA issue rate, so the common cases are kept free of the runtime inner loop.
      auto produce = [&](auto ks_tag) {
        constexpr int KS = decltype(ks_tag)::value;
      for (int unit = walk_start; unit < n_units; unit += walk_step) {
          const int tile = splitk > 1 ? unit / splitk : unit;
          const int kb_lo = splitk > 1 ? (unit - tile * splitk) * k_per : 0, kb_hi = kb_lo + k_per;
          const int tile_mj = tile / pa.n_tiles_n;
          const int tile_m = PAIR ? 2 * tile_mj + (int)rank : tile_mj;
          const int n0 = (tile - tile_mj * pa.n_tiles_n) * BN;
          const int m0 = tile_m * 128;
          int img = 0, base_w = 0, base_h = 0;
          if (a.im2col) {
            img = m0 / hw;
            const int rem = m0 - img * hw;
            const int p0 = rem / a.Wo;
            base_w = (rem - p0 * a.Wo) * a.stride_w - a.pad_w;
            base_h = p0 * a.conv_stride - a.pad;
          }
          int cb = 0, kh = 0, kw = 0;
          if (kb_lo) {                                 // this part starts inside the K walk: (tap, channel block) of K block kb_lo
            const int tap = kb_lo / a.kc_blocks;
            cb = kb_lo - tap * a.kc_blocks;
            kh = tap / a.kw; kw = tap - kh * a.kw;
          }
          for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += (KS ? KS : ksub)) {
            const int cnt = KS ? KS : ((kb_hi - kb0 < ksub) ? kb_hi - kb0 : ksub);      // BK-blocks in this stage
            const long long t0 = dbg ? clk() : 0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (dbg) t_wait += clk() - t0;
            uint8_t* st_base = stages + stage * stage_bytes;
            const uint32_t fb = PAIR ? full0 + (uint32_t)stage * 8u : 0u;
            if (PAIR) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * tx_bytes * (uint32_t)cnt);   // bytes of both CTAs
              else mbar_arrive_cluster(fb);
            } else {
              mbar_expect_tx(&full_bar[stage], tx_bytes * (uint32_t)cnt);
            }
            for (int j = 0; j < cnt; ++j) {
              uint8_t* sa = st_base + j * sub_bytes;
              const int kb = kb0 + j;
              if (PAIR) {
                if (load_a) {
                  if (a.im2col) tma_load_im2col_4d_pair(&tmA, fb, sa, cb * BK, base_w, base_h, img, (uint16_t)kw, (uint16_t)kh);
                  else tma_load_2d_pair(&tmA, fb, sa, cb * BK, m0);
                }
                if (load_b) tma_load_2d_pair(&tmB, fb, sa + A_BYTES, kb * BK, n0 + (int)rank * (BN / 2));
              } else {
                if (load_a) {
                  if (a.im2col) tma_load_im2col_4d(&tmA, &full_bar[stage], sa, cb * BK, base_w, base_h, img, (uint16_t)kw, (uint16_t)kh);
                  else tma_load_2d(&tmA, &full_bar[stage], sa, cb * BK, m0);
                }
                if (load_b) tma_load_2d(&tmB, &full_bar[stage], sa + A_BYTES, kb * BK, n0);
              }
              if (++cb == a.kc_blocks) { cb = 0; if (++kw == a.kw) { kw = 0; ++kh; } }
            }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
      };
      if (ksub == 1) produce(std::integral_constant<int, 1>{});
      else if (ksub == 2 && k_per % 2 == 0) produce(std::integral_constant<int, 2>{});
      else if (ksub == 3 && k_per % 3 == 0) produce(std::integral_constant<int, 3>{});
      else if (ksub == 9 && k_per % 9 == 0) produce(std::integral_constant<int, 9>{});
      else produce(std::integral_constant<int, 0>{});
      if (dbg) {
        atomicAdd(&pa.dbg[CONV_DBG_PROD_WAIT_EMPTY], (unsigned long long)t_wait);
        atomicAdd(&pa.dbg[CONV_DBG_PROD_TOTAL], (unsigned long long)(clk() - t_begin));
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------ MMA issuer (pair: leader CTA only) ------------------------------
    // solo: one elected thread runs the whole issue loop (tcgen05.mma / commit are single-thread instructions);
    // otherwise all lanes walk the loop and elect an issuer per stage.
    auto issue_loop = [&](const bool solo, auto ks_tag) {
      constexpr int KS = decltype(ks_tag)::value;
      constexpr uint32_t idesc = PAIR ? make_idesc_m<BN, 256>() : make_idesc<BN>();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (bstat) {
        mbar_wait(b_full_bar, 0);
        tc_fence_after();
      }
      const uint32_t b_stat_addr = smem_u32(b_stat);
      const uint32_t stages_addr = smem_u32(stages);
      const bool do_mma = !(pa.ablate & 2);
      const bool dbg = pa.dbg != nullptr;
      long long t_full = 0, t_tmem = 0;
      const long long t_begin = dbg ? clk() : 0;
      for (int unit = walk_start; unit < n_units; unit += walk_step, ++it) {
        const int kb_lo = splitk > 1 ? (unit % splitk) * k_per : 0, kb_hi = kb_lo + k_per;
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        long long t0 = dbg ? clk() : 0;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);     // epilogue has drained this accumulator
        if (dbg) t_tmem += clk() - t0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += (KS ? KS : ksub)) {
          const int cnt = KS ? KS : ((kb_hi - kb0 < ksub) ? kb_hi - kb0 : ksub);
          t0 = dbg ? clk() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (dbg) t_full += clk() - t0;
          tc_fence_after();
          if (solo || elect_one()) {
            const uint32_t st_base = stages_addr + (uint32_t)(stage * stage_bytes);
            if (do_mma) {
              for (int j = 0; j < cnt; ++j) {
                const uint32_t sa = st_base + (uint32_t)(j * sub_bytes);
                const uint64_t da = make_kmajor_desc<BK>(sa);
                const uint64_t db = make_kmajor_desc<BK>(bstat ? b_stat_addr + (uint32_t)((kb0 + j) * B_BYTES) : sa + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (>>4) address field
                  const uint32_t accum = ((kb0 + j - kb_lo) | k) != 0 ? 1u : 0u;
                  if (PAIR) umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, accum);
                  else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, accum);
                }
              }
            }
            if (PAIR) {
              umma_commit_pair(&empty_bar[stage], 3);                          // frees this smem stage in both CTAs
              if (kb0 + cnt == kb_hi) umma_commit_pair(&tmem_full_bar[acc], 3);  // accumulator complete
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb0 + cnt == kb_hi) umma_commit(&tmem_full_bar[acc]);
            }
          }
          if (!solo) __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg && (solo || lane == 0)) {
        atomicAdd(&pa.dbg[CONV_DBG_MMA_WAIT_FULL], (unsigned long long)t_full);
        atomicAdd(&pa.dbg[CONV_DBG_MMA_WAIT_TMEM], (unsigned long long)t_tmem);
        atomicAdd(&pa.dbg[CONV_DBG_MMA_TOTAL], (unsigned long long)(clk() - t_begin));
      }
    };
    auto issue = [&](const bool solo) {
      if (ksub == 1) issue_loop(solo, std::integral_constant<int, 1>{});
      else if (ksub == 2 && k_per % 2 == 0) issue_loop(solo, std::integral_constant<int, 2>{});
      else if (ksub == 3 && k_per % 3 == 0) issue_loop(solo, std::integral_constant<int, 3>{});
      else if (ksub == 9 && k_per % 9 == 0) issue_loop(solo, std::integral_constant<int, 9>{});
      else issue_loop(solo, std::integral_constant<int, 0>{});
    };
    if (pa.solo_issue) {
      if (elect_one()) issue(true);
      __syncwarp();
    } else {
      issue(false);
    }
  } else if (warp >= 2) {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                 // 0 or 1
    const uint32_t tmem_empty0 = PAIR ? mapa_u32(smem_u32(tmem_empty_bar), 0) : 0u;   // leader's tmem_empty_bar[0]
    constexpr int NCH = BN / 32;                      // 32-column chunks per tile
    constexpr int CH_PER = (NCH + 1) / 2;
    const int ch_begin0 = half * CH_PER;
    const int ch_end0 = (ch_begin0 + CH_PER < NCH) ? ch_begin0 + CH_PER : NCH;
    const int hw = a.Ho * a.Wo;
    int it = 0;
    int epi_idx = 0;            // running sub-tile counter of this warp (selects the staging buffer)
    uint32_t res_par = 0;       // phase bits of this warp's two residual barriers
    bool res_ahead = false;     // the residual of this warp's next sub-tile is already on its way (TMA epilogue)
    griddep_wait();             // residual reads and output stores must follow the previous kernels
    const bool dbg = pa.dbg != nullptr;
    long long t_acc = 0, t_res = 0, t_buf = 0, t_ld = 0, t_math = 0, t_fs = 0;
    const long long t_begin = dbg ? clk() : 0;
    for (int unit = walk_start; unit < n_units; unit += walk_step, ++it) {
      const int tile = splitk > 1 ? unit / splitk : unit;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      // a 32-column tile has a single chunk per lane quarter: the two warps of a quarter take alternate tiles
      const int ch_begin = (NCH == 1) ? 0 : ch_begin0;
      const int ch_end = (NCH == 1) ? (((it & 1) == half) ? 1 : 0) : ch_end0;
      if (pa.ablate & 1) {      // probe: accumulator handed straight back
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(tmem_empty0 + (uint32_t)acc * 8u);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
        }
        continue;
      }
      const int tile_mj = tile / pa.n_tiles_n;
      const int tile_m = PAIR ? 2 * tile_mj + (int)rank : tile_mj;
      const int n0 = (tile - tile_mj * pa.n_tiles_n) * BN;
      const int m = tile_m * 128 + quarter * 32 + lane;
      const bool valid = m < a.M;
      long long opix[4];
      int n_dst = 1, ch_extra = 0;
      if (a.out_mode == OUT_PLAIN) {
        opix[0] = m;
      } else {
        const int img = m / hw;
        const int rem = m - img * hw;
        const int p = rem / a.Wo;
        const int q = rem - p * a.Wo;
        if (a.out_mode == OUT_UPSAMPLE2) {
          const long long W2 = 2 * a.Wo;
          const long long base = ((long long)img * 2 * a.Ho + 2 * p) * W2 + 2 * q;
          opix[0] = base; opix[1] = base + 1; opix[2] = base + W2; opix[3] = base + W2 + 1;
          n_dst = 4;
        } else {
          opix[0] = ((long long)img * (a.Ho >> 1) + (p >> 1)) * (a.Wo >> 1) + (q >> 1);
          ch_extra = ((p & 1) * 2 + (q & 1)) * a.cout;
        }
      }
      if (pa.tma_epi) {
        // ---------- TMA epilogue: 32 rows x 32 channels per step through swizzled smem ----------
        const int ew = warp - 2;
        uint8_t* bufs = epi_smem + ew * 2 * CONV_TCP_EPI_BUF;
        uint64_t* rbar = res_bar + ew * 2;
        const int m_warp = tile_m * 128 + quarter * 32;                 // first row of this warp
        const bool has_res_t = a.res != nullptr;
        // this lane's row inside the 32x64B sub-tile: 16-byte chunk c lives at chunk (c ^ ((lane >> 1) & 3))
        const uint32_t row_off = (uint32_t)lane * 64u;
        const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
        auto issue_res = [&](int chunk, int idx) {                       // lane 0 only
          const int b = idx & 1;
          tma_store_wait_read<0>();                                      // the store that last used this buffer has read it
          mbar_expect_tx(&rbar[b], CONV_TCP_EPI_BUF);
          tma_load_2d(&tmRes, &rbar[b], bufs + b * CONV_TCP_EPI_BUF, n0 + chunk * 32, m_warp);
        };
        // the first chunk's residual was requested at the end of this warp's previous tile (res_ahead), except for its
        // first tile: a narrow tile has one chunk per warp, so a request made here would expose the full memory latency
        if (has_res_t && ch_begin < ch_end && !res_ahead && lane == 0) issue_res(ch_begin, epi_idx);
        if (ch_begin < ch_end) res_ahead = false;
        long long t0 = dbg ? clk() : 0;
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        if (dbg) t_acc += clk() - t0;
        tc_fence_after();
#pragma unroll 1
        for (int chunk = ch_begin; chunk < ch_end; ++chunk, ++epi_idx) {
          const int b = epi_idx & 1;
          uint8_t* buf = bufs + b * CONV_TCP_EPI_BUF;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + chunk * 32), v);
          uint4 rcur[4];
          if (has_res_t) {
            t0 = dbg ? clk() : 0;
            if (chunk + 1 < ch_end && lane == 0) issue_res(chunk + 1, epi_idx + 1);
            if (dbg) { const long long t1 = clk(); t_buf += t1 - t0; t0 = t1; }
            mbar_wait(&rbar[b], (res_par >> b) & 1u);
            if (dbg) t_res += clk() - t0;
            res_par ^= (1u << b);
#pragma unroll
            for (int g = 0; g < 4; ++g) rcur[g] = *reinterpret_cast<const uint4*>(buf + row_off + (((uint32_t)g ^ sw) << 4));
          } else {
            t0 = dbg ? clk() : 0;
            if (lane == 0) tma_store_wait_read<1>();                     // the store two steps ago used this buffer
            __syncwarp();
            if (dbg) t_buf += clk() - t0;
          }
          t0 = dbg ? clk() : 0;
          tmem_ld_wait();
          if (dbg) { const long long t1 = clk(); t_ld += t1 - t0; t0 = t1; }
          const int cbase = n0 + chunk * 32;
          float f[32];
          {   // BN + leaky on packed fp32 pairs (FFMA2 / FMUL2: the scalar operations, two per instruction)
            const bool leaky = a.leaky != 0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float4 ss = s_ss[(cbase + j) >> 1];
              bn_leaky2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), f2_pack(ss.x, ss.y), f2_pack(ss.z, ss.w), leaky, f[j], f[j + 1]);
            }
          }
          if (has_res_t) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t rw[4] = {rcur[g].x, rcur[g].y, rcur[g].z, rcur[g].w};
#pragma unroll
              for (int j = 0; j < 4; ++j)      // two bf16 -> fp32 and one packed add per channel pair
                f2_unpack(f2_add(f2_pack(f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]),
                                 f2_pack(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xFFFF0000u))),
                          f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]);
            }
          }
          if (a.out_f32) {
            // fp32 heads: the 32-column chunk goes out as two stores of 16 floats (64-byte rows, same staging buffers)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint8_t* hb = bufs + ((epi_idx + hf) & 1) * CONV_TCP_EPI_BUF;
              if (hf == 1) {
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
              }
#pragma unroll
              for (int g = 0; g < 4; ++g)
                *reinterpret_cast<float4*>(hb + row_off + (((uint32_t)g ^ sw) << 4)) =
                    make_float4(f[hf * 16 + g * 4], f[hf * 16 + g * 4 + 1], f[hf * 16 + g * 4 + 2], f[hf * 16 + g * 4 + 3]);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmOut, hb, cbase + hf * 16, m_warp);
                tma_store_commit();
              }
            }
            ++epi_idx;               // two staging buffers were consumed (the loop header adds the other one)
            if (dbg) t_fs += clk() - t0;
            continue;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            __nv_bfloat162 b0 = __floats2bfloat162_rn(f[g * 8 + 0], f[g * 8 + 1]);
            __nv_bfloat162 b1 = __floats2bfloat162_rn(f[g * 8 + 2], f[g * 8 + 3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g * 8 + 4], f[g * 8 + 5]);
            __nv_bfloat162 b3 = __floats2bfloat162_rn(f[g * 8 + 6], f[g * 8 + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&b0);
            pk.y = *reinterpret_cast<uint32_t*>(&b1);
            pk.z = *reinterpret_cast<uint32_t*>(&b2);
            pk.w = *reinterpret_cast<uint32_t*>(&b3);
            *reinterpret_cast<uint4*>(buf + row_off + (((uint32_t)g ^ sw) << 4)) = pk;
          }
          if (dbg) { const long long t1 = clk(); t_math += t1 - t0; t0 = t1; }
          fence_proxy_async();                 // make the generic-proxy smem writes visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, buf, cbase, m_warp);    // rows >= the tensor's row count and channels >= Cout are clipped
            tma_store_commit();
          }
          if (dbg) t_fs += clk() - t0;
        }
        if (has_res_t && ch_begin < ch_end) {
          // residual of the first chunk of this warp's next tile, into the buffer the store before last has left
          const int nt = tile + ((NCH == 1) ? 2 : 1) * walk_step;
          if (nt < pa.n_tiles) {
            if (lane == 0) {
              const int nmj = nt / pa.n_tiles_n;
              const int nm = (PAIR ? 2 * nmj + (int)rank : nmj) * 128 + quarter * 32;
              const int nn0 = (nt - nmj * pa.n_tiles_n) * BN + ((NCH == 1) ? 0 : ch_begin0) * 32;
              const int b = epi_idx & 1;
              tma_store_wait_read<1>();
              mbar_expect_tx(&rbar[b], CONV_TCP_EPI_BUF);
              tma_load_2d(&tmRes, &rbar[b], bufs + b * CONV_TCP_EPI_BUF, nn0, nm);
            }
            res_ahead = true;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(tmem_empty0 + (uint32_t)acc * 8u);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
        }
        continue;
      }
      // residual of the first chunk is requested before waiting for the accumulator
      uint4 rnext[4];
      const bool has_res = a.res != nullptr;
      auto load_res = [&](int chunk, uint4 (&r)[4]) {
        const int cbase = n0 + chunk * 32;
        const __nv_bfloat16* rp = a.res + (long long)m * a.res_ld + cbase;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (valid && cbase + g * 8 < a.cout) r[g] = __ldg(reinterpret_cast<const uint4*>(rp + g * 8));
          else r[g] = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      // BN + leaky + residual + store of one 32-column chunk of this lane's row
      auto finish_chunk = [&](int chunk, const uint32_t (&v)[32], const uint4 (&rcur)[4]) {
        const int cbase = n0 + chunk * 32;
        if (valid && cbase < a.cout) {
          float f[32];
          {
            const bool leaky = a.leaky != 0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float4 ss = s_ss[(cbase + j) >> 1];
              bn_leaky2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), f2_pack(ss.x, ss.y), f2_pack(ss.z, ss.w), leaky, f[j], f[j + 1]);
            }
          }
          if (has_res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t rw[4] = {rcur[g].x, rcur[g].y, rcur[g].z, rcur[g].w};
#pragma unroll
              for (int j = 0; j < 4; ++j)      // two bf16 -> fp32 and one packed add per channel pair
                f2_unpack(f2_add(f2_pack(f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]),
                                 f2_pack(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xFFFF0000u))),
                          f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]);
            }
          }
          if (a.out_f32) {
            float* op = reinterpret_cast<float*>(a.out) + opix[0] * a.out_ld + cbase;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (cbase + g * 4 < a.out_ld)
                *reinterpret_cast<float4*>(op + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
            }
          } else {
            uint4 pk[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              __nv_bfloat162 b0 = __floats2bfloat162_rn(f[g * 8 + 0], f[g * 8 + 1]);
              __nv_bfloat162 b1 = __floats2bfloat162_rn(f[g * 8 + 2], f[g * 8 + 3]);
              __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g * 8 + 4], f[g * 8 + 5]);
              __nv_bfloat162 b3 = __floats2bfloat162_rn(f[g * 8 + 6], f[g * 8 + 7]);
              pk[g].x = *reinterpret_cast<uint32_t*>(&b0);
              pk[g].y = *reinterpret_cast<uint32_t*>(&b1);
              pk[g].z = *reinterpret_cast<uint32_t*>(&b2);
              pk[g].w = *reinterpret_cast<uint32_t*>(&b3);
            }
            for (int d = 0; d < n_dst; ++d) {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + opix[d] * a.out_ld + ch_extra + cbase;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (cbase + g * 8 < a.cout) *reinterpret_cast<uint4*>(op + g * 8) = pk[g];
              }
            }
          }
        }
      };
      if constexpr (SPLIT) {
        // ---------- split-K: park the raw accumulator, the last part to arrive reduces and finishes ----------
        constexpr int R = PAIR ? 2 : 1;
        const int part = unit - tile * splitk;
        // workspace layout: per (unit, pair rank) 128 x BN floats as [lane quarter][chunk][float4 index g][lane] -- a warp's
        // store / load of one float4 per lane covers 512 contiguous bytes (lane-per-row addressing would touch 32 lines)
        const size_t lane_in_unit = ((size_t)rank * 128 * BN) + (size_t)quarter * NCH * 1024 + (size_t)lane * 4;
        float* wrow = pa.ws + ((size_t)(tile * splitk + part) * R) * 128 * BN + lane_in_unit;
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = ch_begin; chunk < ch_end; ++chunk) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + chunk * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            __stcg(reinterpret_cast<float4*>(wrow + chunk * 1024 + g * 128),
                   make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(tmem_empty0 + (uint32_t)acc * 8u);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
        }
        __threadfence();
        __syncwarp();
        unsigned int* cnt = pa.ws_cnt + ((size_t)tile * R + rank) * CONV_TCP_EPI_WARPS + (warp - 2);
        unsigned int arrived = 0;
        if (lane == 0) arrived = atomicAdd(cnt, 1u);
        arrived = __shfl_sync(0xffffffffu, arrived, 0);
        if (arrived == (unsigned)(splitk - 1)) {
          if (lane == 0) *cnt = 0u;                            // ready for the next launch
          __threadfence();
          const float* rrow = pa.ws + ((size_t)(tile * splitk) * R) * 128 * BN + lane_in_unit;
          const size_t part_stride = (size_t)R * 128 * BN;
#pragma unroll 1
          for (int chunk = ch_begin; chunk < ch_end; ++chunk) {
            float sacc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sacc[j] = 0.0f;
            for (int p = 0; p < splitk; ++p) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(rrow + (size_t)p * part_stride + chunk * 1024 + g * 128));
                sacc[g * 4] += t.x; sacc[g * 4 + 1] += t.y; sacc[g * 4 + 2] += t.z; sacc[g * 4 + 3] += t.w;
              }
            }
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(sacc[j]);
            uint4 rcur[4];
            if (has_res) load_res(chunk, rcur);
            finish_chunk(chunk, v, rcur);
          }
        }
        continue;
      }
      if (has_res && ch_begin < ch_end) load_res(ch_begin, rnext);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = ch_begin; chunk < ch_end; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + chunk * 32), v);
        uint4 rcur[4];
        if (has_res) {
#pragma unroll
          for (int g = 0; g < 4; ++g) rcur[g] = rnext[g];
          if (chunk + 1 < ch_end) load_res(chunk + 1, rnext);
        }
        tmem_ld_wait();
        finish_chunk(chunk, v, rcur);
      }
      // this warp has finished reading the accumulator: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(tmem_empty0 + (uint32_t)acc * 8u);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
      }
    }
    if (dbg && lane == 0) {
      atomicAdd(&pa.dbg[CONV_DBG_EPI_WAIT_ACC], (unsigned long long)t_acc);
      atomicAdd(&pa.dbg[CONV_DBG_EPI_WAIT_RES], (unsigned long long)t_res);
      atomicAdd(&pa.dbg[CONV_DBG_EPI_WAIT_BUF], (unsigned long long)t_buf);
      atomicAdd(&pa.dbg[CONV_DBG_EPI_TMEM_LD], (unsigned long long)t_ld);
      atomicAdd(&pa.dbg[CONV_DBG_EPI_MATH_STORE], (unsigned long long)t_math);
      atomicAdd(&pa.dbg[CONV_DBG_EPI_TOTAL], (unsigned long long)(clk() - t_begin));
      atomicAdd(&pa.dbg[CONV_DBG_EPI_FENCE_STORE], (unsigned long long)t_fs);
    }
  }
  if (pa.tma_epi && warp >= 2 && lane == 0) tma_store_wait_all();
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // pair: nobody exits while the peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace yb
