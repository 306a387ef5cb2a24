// engine.cu -- libyolo_b200.so: plan compiler, activation arena, weight packing, launch sequencing and
// the C ABI declared in include/yolo_b200.h.  Host code only orchestrates; all arithmetic of the TEST
// hot path (net/yolo.py:83,86 of the reference) runs in the kernels of conv_tc.cuh, aux_kernels.cuh
// and post.cuh.  There is no CPU fallback: every compute entry point fails with YB_ERR_CUDA when no
// CUDA device is usable.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "../../include/yolo_b200.h"
#include "aux_kernels.cuh"
#include "conv_tc.cuh"
#include "conv_fused.cuh"
#include "post.cuh"
#include "preprocess.cuh"

namespace yb {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define YB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      cudaGetLastError();                                                                          \
      return fail(YB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                              \
  } while (0)
#define YB_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != YB_OK) return _r; \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

static uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// ------------------------------------------------------------------------------------------------
// driver entry points for tensor-map encoding (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_EncodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_EncodeTiled g_encode_tiled = nullptr;
static PFN_EncodeIm2col g_encode_im2col = nullptr;

static int load_driver_fns() {
  if (g_encode_tiled && g_encode_im2col) return YB_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
  g_encode_tiled = (PFN_EncodeTiled)fn;
  fn = nullptr;
  YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(YB_ERR_CUDA, "cuTensorMapEncodeIm2col not available in this driver");
  g_encode_im2col = (PFN_EncodeIm2col)fn;
  return YB_OK;
}

// generic 2-D bf16 map: [rows, cols] with row pitch ld (elements), box = box_rows x box_cols
static int make_map_2d(CUtensorMap* tm, const void* base, long long rows, int cols, int ld, int box_rows, int box_cols,
                       CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%d ld=%d box=%dx%d", (int)r, rows, cols, ld, box_rows, box_cols);
  return YB_OK;
}

static int make_map_2d_f32(CUtensorMap* tm, const void* base, long long rows, int cols, int ld, int box_rows, int box_cols,
                           CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed (%d): rows=%lld cols=%d ld=%d box=%dx%d", (int)r, rows, cols, ld, box_rows, box_cols);
  return YB_OK;
}

static CUtensorMapSwizzle swizzle_for(int bk) { return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B; }

// [rows, cols] bf16 row-major with row pitch ld (elements); box = box_rows x bk
static int make_tiled_map(CUtensorMap* tm, const void* base, long long rows, int cols, int ld, int box_rows, int bk) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%d ld=%d box=%dx%d", (int)r, rows, cols, ld, box_rows, bk);
  return YB_OK;
}

// activation [N,H,W,C] bf16 (pixel pitch ld) as an im2col map: 128 output pixels x bk channels per load
static int make_im2col_map(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int ld, int ksize, int stride,
                           int pad_lo, int bk, int kw = 0, int stride_w = 0, int pad_w = -1) {
  if (kw <= 0) kw = ksize;
  if (stride_w <= 0) stride_w = stride;
  if (pad_w < 0) pad_w = pad_lo;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  const int pad_hi = (ksize - 1) - pad_lo, pad_hi_w = (kw - 1) - pad_w;
  int lower[2] = {-pad_w, -pad_lo};                                // {W, H}: offset of the first window's origin
  int upper[2] = {pad_hi_w - (kw - 1), pad_hi - (ksize - 1)};      // far corner: last window origin relative to the edge
  cuuint32_t es[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper,
                               (cuuint32_t)bk, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d ld=%d k=%d s=%d", (int)r, N, H, W, C, ld, ksize, stride);
  // Same workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): drivers <= 13.1 set a
  // descriptor bit that breaks im2col loads of tensors smaller than 128 KiB.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  const unsigned long long bytes = (unsigned long long)N * H * W * ld * 2;
  if (drv <= 13010 && bytes < 131072ull) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return YB_OK;
}

// ------------------------------------------------------------------------------------------------
// post-processing context (decode + sort + NMS), shared by the engine and the stand-alone yb_post_* API
// ------------------------------------------------------------------------------------------------
struct PostCtx {
  int max_batch = 0, rows = 0, rows_pow2 = 0, box_len = 0, n_scales = 0, v2 = 0;
  ScaleDesc scales[POST_MAX_SCALES];
  float *prob = nullptr, *x = nullptr, *y = nullptr;
  double *w = nullptr, *h = nullptr;
  int* cls = nullptr;
  unsigned long long* keys = nullptr;
  void* sorted_boxes = nullptr;
  float4 *sorted_f4 = nullptr, *sorted_f4s = nullptr, *sorted_f4i = nullptr, *sorted_ky = nullptr, *sorted_kx = nullptr;
  float2* sorted_area = nullptr;
  int* sorted_cls = nullptr;
  unsigned long long* sorted_qidx = nullptr;
  unsigned char* flags = nullptr;
  int *order = nullptr, *n_keep = nullptr, *n_cand = nullptr;
  DetOut* dets = nullptr;
  size_t dets_cap = 0;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  bool smem_set = false;
  bool bulk_attr_set = false;     // per context (= per device): opt-in shared memory of decode_v3_bulk_kernel

  int alloc(int max_batch_, int rows_, int box_len_, int v2_) {
    max_batch = max_batch_; rows = rows_; box_len = box_len_; v2 = v2_;
    rows_pow2 = 1;
    while (rows_pow2 < rows) rows_pow2 <<= 1;
    const size_t nr = (size_t)max_batch * rows;
    YB_CUDA(cudaMalloc(&prob, nr * 4)); YB_CUDA(cudaMalloc(&x, nr * 4)); YB_CUDA(cudaMalloc(&y, nr * 4));
    YB_CUDA(cudaMalloc(&w, nr * 8)); YB_CUDA(cudaMalloc(&h, nr * 8)); YB_CUDA(cudaMalloc(&cls, nr * 4));
    if (rows_pow2 > NMS_SMEM_KEYS) YB_CUDA(cudaMalloc(&keys, (size_t)max_batch * rows_pow2 * 8));
    YB_CUDA(cudaMalloc(&sorted_boxes, nr * sizeof(BoxC<double>)));
    YB_CUDA(cudaMalloc(&sorted_f4, nr * sizeof(float4)));
    YB_CUDA(cudaMalloc(&sorted_f4s, nr * sizeof(float4)));
    YB_CUDA(cudaMalloc(&sorted_ky, nr * sizeof(float4)));
    YB_CUDA(cudaMalloc(&sorted_kx, nr * sizeof(float4)));
    YB_CUDA(cudaMalloc(&sorted_f4i, nr * sizeof(float4)));
    YB_CUDA(cudaMalloc(&sorted_area, nr * sizeof(float2)));
    YB_CUDA(cudaMalloc(&sorted_cls, nr * 4)); YB_CUDA(cudaMalloc(&flags, nr));
    YB_CUDA(cudaMalloc(&sorted_qidx, nr * 8));
    YB_CUDA(cudaMalloc(&order, nr * 4));
    YB_CUDA(cudaMalloc(&n_keep, (size_t)max_batch * 4)); YB_CUDA(cudaMalloc(&n_cand, (size_t)max_batch * 4));
    dets_cap = 0;
    for (int i = 0; i < 3; ++i) YB_CUDA(cudaEventCreate(&ev[i]));
    return YB_OK;
  }
  void release() {
    cudaFree(prob); cudaFree(x); cudaFree(y); cudaFree(w); cudaFree(h); cudaFree(cls); cudaFree(keys);
    cudaFree(sorted_boxes); cudaFree(sorted_f4); cudaFree(sorted_f4s); cudaFree(sorted_ky); cudaFree(sorted_kx); cudaFree(sorted_f4i); cudaFree(sorted_area); cudaFree(sorted_cls); cudaFree(sorted_qidx); cudaFree(flags); cudaFree(order); cudaFree(n_keep); cudaFree(n_cand);
    cudaFree(dets);
    for (int i = 0; i < 3; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
  }

  // scales[] must carry base/img_stride/cell_stride for this call
  int decode(cudaStream_t st, const ScaleDesc* sc, int n, float thr) {
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < n_scales; ++i) a.sc[i] = sc[i];
    a.n_scales = n_scales; a.rows = rows; a.box_len = box_len; a.n_images = n; a.v2 = v2; a.thr = thr;
    a.prob = prob; a.x = x; a.y = y; a.w = w; a.h = h; a.cls = cls;
    if (!v2) {
      // Dense candidate sets on a contiguous [n * rows][5 + C] tensor (yb_post_run's layout; BASELINE config 5) take the
      // bulk kernel: whole rows staged through shared memory, one lane per row.  Sparse sets (real detections: a few
      // candidates per image) keep the kernel that touches one 32-byte sector per non-candidate row.  The choice is made
      // from the score threshold (a low threshold is what makes a head dense); YB_DECODE_BULK = 0 / 1 forces it.
      bool contiguous = (reinterpret_cast<uintptr_t>(sc[0].base) & 15) == 0 && sc[0].row_begin == 0;
      for (int i = 0; i < n_scales && contiguous; ++i)
        contiguous = sc[i].base == sc[0].base + (long long)sc[i].row_begin * box_len && sc[i].cell_stride == sc[i].na * box_len &&
                     sc[i].img_stride == (long long)rows * box_len;
      const char* force = getenv("YB_DECODE_BULK");
      const bool bulk = contiguous && (force ? atoi(force) != 0 : thr < 0.02f);
      const int bulk_smem = DECODE_BULK_WARPS * 32 * box_len * 4;
      if (bulk && bulk_smem <= 100 * 1024) {
        if (!bulk_attr_set) {
          YB_CUDA(cudaFuncSetAttribute(decode_v3_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          bulk_attr_set = true;
        }
        decode_v3_bulk_kernel<<<ceil_div((long long)n * rows, DECODE_BULK_WARPS * 32), DECODE_BULK_WARPS * 32, bulk_smem, st>>>(a, sc[0].base);
      } else {       // one lane per row (+ cooperative decode of the candidates)
        decode_v3_kernel<<<ceil_div((long long)n * rows, 256), 256, 0, st>>>(a);
      }
    } else {         // one warp per row: the softmax score needs the whole row anyway
      const long long warps = (long long)n * rows;
      decode_kernel<<<ceil_div(warps * 32, 256), 256, 0, st>>>(a);
    }
    YB_CUDA(cudaGetLastError());
    return YB_OK;
  }

  template <typename T, bool XY64, bool WH64>
  int nms_launch(cudaStream_t st, int n, const NmsArgs& a) {
    auto kern = sort_nms_kernel<T, XY64, WH64>;
    const int smem = NMS_DYN_SMEM;
    YB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<n, NMS_THREADS, smem, st>>>(a);
    YB_CUDA(cudaGetLastError());
    return YB_OK;
  }

  NmsArgs nms_args(double thr, int per_class) const {
    NmsArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = rows; a.per_class = per_class; a.thr = thr;
    { const char* dbg = getenv("YB_NMS_DEBUG"); a.debug = dbg ? atoi(dbg) : 0; }
    a.prob = prob; a.x = x; a.y = y; a.w = w; a.h = h; a.cls = cls;
    a.keys = keys; a.rows_pow2 = rows_pow2; a.sorted_boxes = sorted_boxes; a.sorted_f4 = sorted_f4; a.sorted_f4s = sorted_f4s; a.sorted_ky = sorted_ky; a.sorted_kx = sorted_kx; a.sorted_f4i = sorted_f4i; a.sorted_area = sorted_area; a.sorted_cls = sorted_cls; a.sorted_qidx = sorted_qidx;
    a.flags = flags; a.order = order; a.n_keep = n_keep; a.n_cand = n_cand;
    return a;
  }

  // engine path: float32 x,y + float64 w,h, IoU in float64 (the reference's dtypes)
  int nms(cudaStream_t st, int n, double iou_thr, int mode) {
    YB_CUDA(cudaMemsetAsync(flags, 0, (size_t)n * rows, st));
    return nms_launch<double, false, true>(st, n, nms_args(iou_thr, mode == YB_NMS_PER_CLASS));
  }

  // out[n][max_per_image] on the device; at most min(n_keep, max_per_image) records per image are written
  int gather(cudaStream_t st, int n, int max_per_image) {
    const size_t need = (size_t)n * max_per_image;
    if (need > dets_cap) {
      YB_CUDA(cudaStreamSynchronize(st));
      cudaFree(dets); dets = nullptr; dets_cap = 0;
      YB_CUDA(cudaMalloc(&dets, need * sizeof(DetOut)));
      dets_cap = need;
    }
    dim3 grid(ceil_div(std::min(max_per_image, rows), 128), n);
    gather_dets_kernel<<<grid, 128, 0, st>>>(rows, max_per_image, order, n_keep, prob, x, y, w, h, cls, dets);
    YB_CUDA(cudaGetLastError());
    return YB_OK;
  }
};

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
struct Shape { int h = 0, w = 0, c = 0; };
struct View {
  int buf = -1;       // index into bufs, -2 = network input, -1 = none (fused away / not a tensor)
  int coff = 0, ld = 0, h = 0, w = 0, c = 0;
  bool f32 = false;
};
struct Buf {
  size_t bytes = 0, off = 0;
  int first = INT_MAX, last = -1;
  bool keep = false;
};
enum OpKind { OP_CONV = 0, OP_MAXPOOL, OP_ADD, OP_UPSAMPLE, OP_REORG, OP_COPY };
enum ConvPath { PATH_TC = 0, PATH_DIRECT = 1, PATH_SIMT = 2, PATH_FUSED = 3 };
enum FuseKind { FUSE_NONE = 0, FUSE_STEM = 1, FUSE_BLOCK = 2, FUSE_CONVPOOL = 3 };

// How one tcgen05 conv is launched.  compile_plan fills in a heuristic; yb_engine_autotune replaces it with the
// fastest measured candidate.
struct ConvCfg {
  int bn = 0;          // N tile (TMEM columns per accumulator): 32, 64, 128 or 256
  int pair = 0;        // 1: cta_group::2 CTA pairs computing 256 x bn tiles (bn >= 64)
  int bstat = -1;      // weight-stationary B: -1 = heuristic (whenever it fits), 0 = off, 1 = on if it fits
  int tma_epi = -1;    // TMA-store epilogue: -1 = heuristic, 0 = off, 1 = on if the output allows it
  int ksub = 0;        // BK-blocks per pipeline stage: 0 = heuristic (enough MMA work per barrier hand-shake)
};

struct Op {
  int kind = 0, layer = -1;
  View in, in2, out;
  // conv
  int cout = 0, cout_pad = 0, cin = 0, ksize = 0, stride = 1, pad = 0, leaky = 0, has_res = 0, out_mode = 0;
  int Ho = 0, Wo = 0, path = PATH_TC, bn_max = 0, bk = 0;
  // Pixel-pair view (px_pair = 1): a stride-1 conv with 32 input channels is launched as the same conv over PAIRS of
  // horizontally adjacent pixels -- [N,H,W,32] read as [N,H,W/2,64], [N,H,W,Cout] written as [N,H,W/2,2*Cout], the
  // weights re-packed with structural zeros (load_weights).  Same bytes, same results bit for bit (the extra products
  // are exact zeros added to an fp32 sum), but 128-byte im2col rows instead of 64-byte ones -- the TMA engine delivers
  // about one row per 2-3 cycles whatever its length -- and an N tile twice as wide, which an M=128 MMA gets for free.
  // cin / cout / Wo and the views below then describe the paired problem.
  // px_pair = 2: the stride-2 3x3 conv with 32 input channels reads its INPUT in pairs ([N,H,W/2,64]); output pixel x
  // needs input pixels 2x-1, 2x, 2x+1 = the second pixel of pair x-1 and both of pair x, i.e. a 3 (high) x 2 (wide) kernel
  // over pairs with stride (2, 1) and padding 1 on the left only.  Six 128-byte im2col rows per pixel instead of nine
  // 64-byte ones; the output side is unchanged.
  int px_pair = 0;
  int kw = 0, stride_w = 0, pad_w = 0;      // kernel width, stride and low padding along W (= ksize, stride, pad unless px_pair == 2)
  // Two-layer fusion (conv_fused.cuh): this op is the 3x3 consumer conv and also computes its producer conv
  // ops[fuse_src] (which is then never launched: skip; FUSE_CONVPOOL has no producer: fuse_src = -1) -- FUSE_STEM: first conv 3->32 + 3x3 s2 32->64;
  // FUSE_BLOCK: 1x1 64->32 + 3x3 s1 32->64 + shortcut.
  int fuse_kind = FUSE_NONE, fuse_src = -1;
  bool skip = false;
  ConvCfg cfg;
  int launched_stages = 0, launched_bstat = 0, launched_tma_epi = 0, launched_ksub = 0, launched_splitk = 1;   // what the last launch resolved to
  float tuned_ms = 0.f, default_ms = 0.f;                              // autotune: best candidate vs the heuristic
  __nv_bfloat16* d_wt = nullptr;            // [cout_pad][K], cout_pad = round_up(cout, bn_max)
  float *d_wt32 = nullptr, *d_scale = nullptr, *d_shift = nullptr;
  std::vector<float> h_scale, h_shift;      // host copies (the fused kernels take the consumer's BN by value)
  alignas(64) CUtensorMap tmA, tmOut, tmRes;
  alignas(64) CUtensorMap tmOut4;           // fused ops: output as (C, W, H, N), box 64 x 16 x 2 x 1
  alignas(64) CUtensorMap tmIn4;            // fused block: producer input (= residual) as (C, W, H, N), box 64 x 18 x 10 x 1
  alignas(64) CUtensorMap tmW1;             // fused block: the 1x1 producer's weights [32][64], one box
  alignas(64) CUtensorMap tmB[4];           // weight maps with box rows 32, 64, 128, 256 (128 doubles as the pair half)
  bool tma_epi = false;
  // generic
  int factor = 0;
};

struct Head {
  int yolo_layer = -1, conv_layer = -1;
  View view;
  int h = 0, w = 0, na = 0, row_begin = 0;
  float anchors[2 * YB_MAX_ANCHORS];
};

}  // namespace yb

// ------------------------------------------------------------------------------------------------
// device-side image preprocessing (cv2.resize INTER_LINEAR + BGR->RGB), shared by the engine and yb_resize_bgr2rgb
// ------------------------------------------------------------------------------------------------
namespace yb {

// cv2's coefficient tables (modules/imgproc/src/resize.cpp, cv::resize -> resizeGeneric_ for CV_8U INTER_LINEAR),
// computed in the same types: double scale, float fraction, short weights with 11 fractional bits.
static void build_resize_table(int src, int dst, bool clamp_weights, std::vector<ResizeTab>& out) {
  const double inv_scale = (double)dst / (double)src;
  const double scale = 1. / inv_scale;
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= (float)s;
    if (clamp_weights) {                      // columns only: the row index is clamped where it is used instead
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    ResizeTab t;
    t.ofs = s;
    t.c0 = (short)std::lrintf((1.f - f) * 2048.f);      // cvRound: round half to even
    t.c1 = (short)std::lrintf(f * 2048.f);
    out.push_back(t);
  }
}

struct PreCtx {
  unsigned char* raw[2] = {nullptr, nullptr};     // device copies of the caller's images (double-buffered)
  size_t raw_cap[2] = {0, 0};
  ResizeImage* d_imgs[2] = {nullptr, nullptr};
  ResizeTab* d_tabs[2] = {nullptr, nullptr};
  size_t imgs_cap[2] = {0, 0}, tabs_cap[2] = {0, 0};
  // Pinned staging of the caller's (pageable) images: a few host threads gather them into page-locked memory, chunk by
  // chunk, and each chunk goes out as one DMA while the next is gathered.  A cudaMemcpyAsync straight from pageable memory
  // runs at ~10 GB/s and blocks the caller for its whole length (128 decoded 720p frames: 34 ms per step against 9 ms of
  // compute).
  unsigned char* pinned[2] = {nullptr, nullptr};
  size_t pinned_cap[2] = {0, 0};
  cudaEvent_t pin_done[2] = {nullptr, nullptr};   // the DMAs out of pinned[i] have completed

  void release() {
    for (int i = 0; i < 2; ++i) {
      cudaFree(raw[i]); cudaFree(d_imgs[i]); cudaFree(d_tabs[i]); raw[i] = nullptr; d_imgs[i] = nullptr; d_tabs[i] = nullptr; raw_cap[i] = imgs_cap[i] = tabs_cap[i] = 0;
      if (pin_done[i]) { cudaEventSynchronize(pin_done[i]); cudaEventDestroy(pin_done[i]); pin_done[i] = nullptr; }
      if (pinned[i]) cudaFreeHost(pinned[i]);
      pinned[i] = nullptr; pinned_cap[i] = 0;
    }
  }

  // Uploads n BGR images (host memory) on copy stream `cs`, builds their tables and enqueues the resize kernel on `ks`
  // after the copies (event `ev`).  dst: device [n, dh, dw, 3] uint8 RGB.
  int run(int slot, cudaStream_t cs, cudaStream_t ks, cudaEvent_t ev, const void* const* images, const int* heights, const int* widths,
          const int* strides, int n, int dh, int dw, unsigned char* dst) {
    size_t total = 0;
    std::vector<size_t> offs(n);
    for (int i = 0; i < n; ++i) {
      if (!images[i] || heights[i] <= 0 || widths[i] <= 0) return fail(YB_ERR_INVALID, "image %d: bad pointer or size", i);
      const int stride = strides ? strides[i] : widths[i] * 3;
      if (stride < widths[i] * 3) return fail(YB_ERR_INVALID, "image %d: row stride %d smaller than %d", i, stride, widths[i] * 3);
      offs[i] = total;
      total += ((size_t)heights[i] * stride + 255) & ~(size_t)255;
    }
    if (total > raw_cap[slot]) {
      YB_CUDA(cudaStreamSynchronize(ks));
      cudaFree(raw[slot]); raw[slot] = nullptr; raw_cap[slot] = 0;
      YB_CUDA(cudaMalloc(&raw[slot], total + total / 4));
      raw_cap[slot] = total + total / 4;
    }
    std::vector<ResizeImage> imgs(n);
    std::vector<ResizeTab> tabs;
    struct Key { int sh, sw, xtab, ytab; };
    std::vector<Key> seen;
    // upload: small batches straight from the caller's memory, large ones through the pinned staging buffer
    const bool staged = total >= ((size_t)4 << 20);
    if (staged) {
      // the DMAs of this buffer's previous use (two calls ago) must have read it
      if (!pin_done[slot]) YB_CUDA(cudaEventCreateWithFlags(&pin_done[slot], cudaEventDisableTiming));
      else YB_CUDA(cudaEventSynchronize(pin_done[slot]));
      if (total > pinned_cap[slot]) {
        if (pinned[slot]) cudaFreeHost(pinned[slot]);
        pinned[slot] = nullptr; pinned_cap[slot] = 0;
        YB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&pinned[slot]), total + total / 4, cudaHostAllocDefault));
        pinned_cap[slot] = total + total / 4;
      }
      const unsigned hc = std::thread::hardware_concurrency();
      const char* gt = getenv("YB_GATHER_THREADS");
      const int n_thr = gt ? std::max(1, atoi(gt)) : (int)std::max(1u, std::min(16u, hc ? hc : 1u));
      const int n_chunks = std::min(n, 4);
      for (int c = 0; c < n_chunks; ++c) {
        const int i0 = (int)((long long)n * c / n_chunks), i1 = (int)((long long)n * (c + 1) / n_chunks);
        auto gather = [&](int t) {
          for (int i = i0 + t; i < i1; i += n_thr) {
            const int stride = strides ? strides[i] : widths[i] * 3;
            memcpy(pinned[slot] + offs[i], images[i], (size_t)heights[i] * stride);
          }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_thr && i0 + t < i1; ++t) pool.emplace_back(gather, t);
        gather(0);
        for (auto& th : pool) th.join();
        const size_t b0 = offs[i0], b1 = (i1 < n) ? offs[i1] : total;
        YB_CUDA(cudaMemcpyAsync(raw[slot] + b0, pinned[slot] + b0, b1 - b0, cudaMemcpyHostToDevice, cs));
      }
      YB_CUDA(cudaEventRecord(pin_done[slot], cs));
    }
    for (int i = 0; i < n; ++i) {
      const int stride = strides ? strides[i] : widths[i] * 3;
      if (!staged) YB_CUDA(cudaMemcpyAsync(raw[slot] + offs[i], images[i], (size_t)heights[i] * stride, cudaMemcpyHostToDevice, cs));
      ResizeImage& im = imgs[i];
      im.src = raw[slot] + offs[i]; im.sh = heights[i]; im.sw = widths[i]; im.stride = stride;
      im.area2x = (widths[i] == 2 * dw && heights[i] == 2 * dh) ? 1 : 0;
      im.xtab = im.ytab = 0;
      if (im.area2x) continue;
      const Key* hit = nullptr;
      for (const Key& k : seen) if (k.sh == im.sh && k.sw == im.sw) { hit = &k; break; }
      if (hit) { im.xtab = hit->xtab; im.ytab = hit->ytab; continue; }
      im.xtab = (int)tabs.size();
      build_resize_table(im.sw, dw, true, tabs);
      im.ytab = (int)tabs.size();
      build_resize_table(im.sh, dh, false, tabs);
      seen.push_back(Key{im.sh, im.sw, im.xtab, im.ytab});
    }
    if (tabs.empty()) tabs.push_back(ResizeTab{0, 0, 0});
    if ((size_t)n > imgs_cap[slot]) {
      YB_CUDA(cudaStreamSynchronize(ks));
      cudaFree(d_imgs[slot]); d_imgs[slot] = nullptr;
      YB_CUDA(cudaMalloc(&d_imgs[slot], (size_t)n * sizeof(ResizeImage)));
      imgs_cap[slot] = n;
    }
    if (tabs.size() > tabs_cap[slot]) {
      YB_CUDA(cudaStreamSynchronize(ks));
      cudaFree(d_tabs[slot]); d_tabs[slot] = nullptr;
      YB_CUDA(cudaMalloc(&d_tabs[slot], tabs.size() * 2 * sizeof(ResizeTab)));
      tabs_cap[slot] = tabs.size() * 2;
    }
    // descriptors and tables are small: synchronous copies from these temporaries on the copy stream
    YB_CUDA(cudaMemcpyAsync(d_imgs[slot], imgs.data(), (size_t)n * sizeof(ResizeImage), cudaMemcpyHostToDevice, cs));
    YB_CUDA(cudaMemcpyAsync(d_tabs[slot], tabs.data(), tabs.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice, cs));
    // No synchronisation: the descriptor and table vectors are pageable, so those copies have left host memory when the
    // calls return, and the images were either copied the same way or gathered into the pinned buffer above.
    YB_CUDA(cudaEventRecord(ev, cs));
    YB_CUDA(cudaStreamWaitEvent(ks, ev, 0));
    dim3 grid(ceil_div(dw, 128), dh, n);
    resize_bgr2rgb_kernel<<<grid, 128, 0, ks>>>(d_imgs[slot], d_tabs[slot], dh, dw, dst);
    YB_CUDA(cudaGetLastError());
    return YB_OK;
  }
};

}  // namespace yb

using namespace yb;

struct yb_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<yb_layer> plan;
  int H = 0, W = 0, C = 0, max_batch = 0, decode_mode = 0, num_classes = 0;
  bool keep_all = false;
  int bn_max = 256;
  bool b_stationary = true;
  bool tma_epilogue = true;
  bool cta_pairs = true;
  bool pdl = true;              // programmatic dependent launch between consecutive tcgen05 convs
  bool fuse_upsample = true;    // conv epilogue writes the 2x2 replicas itself (YB_FUSE_UPSAMPLE=0: separate copy kernel)
  int fuse_stem = 1, fuse_block = 1;   // conv_fused.cuh (YB_FUSE_STEM / YB_FUSE_BLOCK: 0 = off, 1 = on unless YB_KEEP_ALL, 2 = always)
  int first_resident[4] = {1, 1, 1, 1};   // co-resident blocks per SM of conv_first_mma_kernel<float / uint8> (+2: pooled variant)
  int fuse_pool = 1;                   // first conv + 2x2 max pool in one kernel (YB_FUSE_POOL; off under YB_KEEP_ALL unless 2)
  int pixel_pairs = 1;          // Op::px_pair: 1 = the Cin=32 stride-1 convs (default), 2 = also the stride-2 one (measured
                                // slower: 0.423 -> 0.441 ms, K grows by a third), 0 = plain views (YB_PIXEL_PAIRS)
  int solo_issue = 1;           // MMA issue loop run by one thread (1) or by the whole warp electing per stage (0)
  int ablate = 0;               // debug probes of the persistent conv kernel (see PersistArgs::ablate)
  unsigned long long* dbg_counters = nullptr;   // device [CONV_DBG_COUNT] cycle counters while "cycles" is switched on
  // Split-K latency mode (option "split_k" / YB_SPLIT_K, off by default: it changes the fp32 summation order with the
  // batch size): partial accumulators of the convs whose grids fill a fraction of the SMs, and their arrival counters
  int split_k = 0;
  float* splitk_ws = nullptr; size_t splitk_ws_bytes = 0;
  unsigned int* splitk_cnt = nullptr; size_t splitk_cnt_n = 0;
  int num_sms = 148;
  std::string tune_report;      // JSON written by yb_engine_autotune
  std::vector<Shape> shape;
  std::vector<View> view;
  std::vector<Buf> bufs;
  std::vector<Op> ops;
  std::vector<Head> heads;
  int rows = 0, box_len = 0;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  void* input_dev[2] = {nullptr, nullptr};   // double-buffered staging for host images
  cudaStream_t copy_stream = nullptr;        // H2D of batch i+1 overlaps the compute of batch i
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_input_free[2] = {nullptr, nullptr};
  int stage_toggle = 0;
  int last_input_reader = 0;                 // index of the last op that reads the network input
  cudaEvent_t ev_input_read = nullptr;       // recorded behind the last reader of the network input (yb_engine_order_before)
  cudaEvent_t marks[8] = {nullptr};
  cudaEvent_t ev_fetch[2] = {nullptr, nullptr};
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_events;      // (n_ops + 1) per profiled forward
  const void* cur_input = nullptr;
  int cur_input_dtype = YB_F32;
  struct StemMap { const void* ptr; int dtype; alignas(64) CUtensorMap map; };
  std::deque<StemMap> stem_maps;            // launch_fused: input tensor maps per (address, element type); deque: stable addresses
  float* d_u8_lut = nullptr;
  float* scratch_f32 = nullptr;   // read_output / read_layer staging
  size_t scratch_floats = 0;
  PostCtx post;
  PreCtx pre;
  bool last_input_is_u8_staged = false;      // the last forward's input sits in input_dev[] as uint8 (forward_raw)
  bool weights_loaded = false;
  int last_n = 0;
  bool detected = false;
  int conv_impl = 0;
  int fwd_launches = 0, det_launches = 0;
  // CUDA graphs of the forward (one per batch size / input buffer / input dtype).  Small batches are launch-bound:
  // 75 launches at ~9 us each against a few hundred us of device work.  A key is run eagerly the first time (one-time
  // attribute calls, autotuned configurations settle), captured the second time, replayed afterwards.
  struct FwdGraph { int n; const void* input; int dtype; int seen; cudaGraphExec_t exec; };
  std::vector<FwdGraph> graphs;
  int graph_mode = -1;            // -1: automatic (n <= YB_GRAPH_AUTO_MAX_N), 0: never, 1: always
  int graph_replays = 0;
};
constexpr int YB_GRAPH_AUTO_MAX_N = 32;

static void clear_graphs(yb_engine* e) {
  for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

namespace yb {

static void* view_ptr(const yb_engine* e, const View& v) {
  if (v.buf == -2) return const_cast<void*>(e->cur_input);
  if (v.buf < 0) return nullptr;
  return e->arena + e->bufs[v.buf].off + (size_t)v.coff * (v.f32 ? 4 : 2);
}

static inline int bn_index(int bn) { return bn == 32 ? 0 : bn == 64 ? 1 : bn == 128 ? 2 : 3; }

struct LaunchEnv {
  int device, num_sms;
  bool allow_bstat, allow_tma_epi, pdl;
  int ablate;
  unsigned long long* dbg;
  int solo_issue;
  int split_k;                 // 1: grids that fill a fraction of the SMs are split along K (latency mode)
  float* ws; size_t ws_bytes;  // split-K workspace
  unsigned int* ws_cnt; size_t ws_cnt_n;
};

// Split-K factor for a grid of `tiles` work units on `cap` CTAs (or CTA pairs): the S in {1,2,3,4} dividing the K walk
// that minimises waves / S, a part keeping at least `min_blocks` K blocks; S > 1 must win by more than its overhead
// (a round trip of the accumulator through L2).
static int choose_splitk(int tiles, int cap, int num_k, int ksub, int bk, int bn) {
  // time model in units of one K block of one tile (BK/16 MMAs of ~BN/2 cycles, at least 128 cycles each for M = 128 in
  // SS mode): waves * blocks per part + a split's overhead (accumulator out to L2, fence, S parts back in, ~4 us = 7,500
  // cycles) expressed in the same unit
  static const double overhead_cycles = getenv("YB_SPLIT_K_OVERHEAD") ? atof(getenv("YB_SPLIT_K_OVERHEAD")) : 7500.0;
  static const double min_gain = getenv("YB_SPLIT_K_GAIN") ? atof(getenv("YB_SPLIT_K_GAIN")) : 0.85;
  const double block_cycles = (bk / 16) * 128.0;
  const double overhead = overhead_cycles / block_cycles;
  int best = 1;
  double best_t = (double)ceil_div(tiles, cap) * num_k;
  for (int S = 2; S <= 4; ++S) {
    if (num_k % (S * ksub) != 0) continue;
    const double t = (double)ceil_div(tiles * S, cap) * (num_k / S) + overhead;
    if (t < min_gain * best_t) { best_t = t; best = S; }
  }
  return best;
}

// Resolves a ConvCfg into the launch parameters of conv_tc_persist_kernel<BN,BK,PAIR> and launches it.
template <int BN, int BK, bool PAIR>
static int launch_conv_tcp(cudaStream_t st, Op& op, const ConvArgs& a, const LaunchEnv& env, const ConvCfg& cfg) {
  constexpr bool CAN_SPLIT = BK == 64 && BN >= 64;     // the split-K instantiations that exist
  auto kern = conv_tc_persist_kernel<BN, BK, PAIR>;
  const int a_bytes = 128 * BK * 2, b_bytes = (PAIR ? BN / 2 : BN) * BK * 2;
  const int num_k = a.taps * a.kc_blocks;
  PersistArgs pa;
  memset(&pa, 0, sizeof(pa));
  const int tiles_m = PAIR ? ceil_div(ceil_div(a.M, 128), 2) : ceil_div(a.M, 128);    // pair: pairs of M tiles
  pa.cout_pad = round_up(op.cout, BN);
  pa.n_tiles_n = pa.cout_pad / BN;
  pa.n_tiles = tiles_m * pa.n_tiles_n;
  pa.ablate = env.ablate;
  pa.dbg = env.dbg;
  pa.solo_issue = env.solo_issue;
  // TMA-store epilogue.  Heuristic: the staging buffers cost one pipeline stage at BN=256, worth it while the epilogue
  // is the long pole (K <= 1152) or the pair kernel halves the operand bytes anyway.
  bool tma_epi = env.allow_tma_epi && op.tma_epi && cfg.tma_epi != 0;
  if (cfg.tma_epi < 0) tma_epi = tma_epi && (BN < 256 || PAIR || num_k * BK <= 1152);
  pa.tma_epi = tma_epi ? 1 : 0;
  const int budget = CONV_TCP_TILE_BUDGET - (pa.tma_epi ? CONV_TCP_EPI_BYTES : 0);
  // weight-stationary when one N tile covers Cout and at least 4 A stages still fit next to the weights
  const long long b_total = (long long)num_k * b_bytes;
  pa.b_stationary = (env.allow_bstat && cfg.bstat != 0 && pa.n_tiles_n == 1 && b_total + 4ll * a_bytes <= budget) ? 1 : 0;
  const int sub_bytes = a_bytes + (pa.b_stationary ? 0 : b_bytes);
  const int avail = budget - (pa.b_stationary ? (int)b_total : 0);
  // BK-blocks per stage.  The issue loop and the barrier round trip cost a few hundred cycles per stage whatever the
  // tile, so a stage should carry ~512 cycles of tensor work: one BK-block is (BK/16) MMAs of BN/2 cycles each.
  int ksub = cfg.ksub;
  if (ksub <= 0) {
    // smallest divisor of the K walk that reaches the target (3 or 9 taps, 2 or 4 channel blocks ...), as long as two
    // stages of it fit; otherwise the largest divisor that does
    const int target = std::min(num_k, std::max(1, 512 / ((BK / 16) * (BN / 2))));
    ksub = 1;
    for (int d = 1; d <= num_k; ++d) {
      if (num_k % d != 0 || avail / (d * sub_bytes) < 2) continue;
      ksub = d;
      if (d >= target) break;
    }
  } else {
    ksub = std::min(ksub, num_k);
    while (ksub > 1 && avail / (ksub * sub_bytes) < 2) --ksub;
  }
  // split-K (latency mode): direct epilogue only, N tiles of at least two chunks (the warps of a quarter then own fixed columns)
  const int cap = PAIR ? env.num_sms / 2 : env.num_sms;
  pa.splitk = 1;
  if (CAN_SPLIT && env.split_k && env.ws && pa.n_tiles < 2 * cap) {
    int S = choose_splitk(pa.n_tiles, cap, num_k, ksub, BK, BN);
    const size_t unit_bytes = (size_t)(PAIR ? 2 : 1) * 128 * BN * 4;
    while (S > 1 && ((size_t)pa.n_tiles * S * unit_bytes > env.ws_bytes || num_k % (S * ksub) != 0)) --S;
    if ((size_t)pa.n_tiles * (PAIR ? 2 : 1) * CONV_TCP_EPI_WARPS > env.ws_cnt_n) S = 1;
    if (S > 1 && pa.tma_epi) {       // the staging buffers return to the pipeline
      pa.tma_epi = 0;
    }
    pa.splitk = S;
    pa.ws = env.ws; pa.ws_cnt = env.ws_cnt;
  }
  pa.ksub = ksub;
  const int stage_bytes = ksub * sub_bytes;
  pa.n_stages = std::min(CONV_TCP_MAX_STAGES, avail / stage_bytes);
  if (pa.n_stages < 2) return fail(YB_ERR_INVALID, "persistent conv: shared memory too small for BN=%d BK=%d", BN, BK);
  const int smem = 1024 + CONV_TCP_HEADER + (pa.tma_epi ? CONV_TCP_EPI_BYTES : 0) + (pa.b_stationary ? (int)b_total : 0) +
                   pa.n_stages * stage_bytes;
  // (the opt-in shared-memory limit of every instantiation is raised once per engine: set_kernel_attrs)
  op.launched_stages = pa.n_stages; op.launched_bstat = pa.b_stationary; op.launched_tma_epi = pa.tma_epi; op.launched_ksub = pa.ksub;
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  const int n_units = pa.n_tiles * pa.splitk;
  op.launched_splitk = pa.splitk;
  lc.gridDim = PAIR ? dim3(2 * std::min(n_units, env.num_sms / 2)) : dim3(std::min(n_units, env.num_sms));
  lc.blockDim = dim3(CONV_TCP_THREADS);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (env.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  lc.attrs = attr; lc.numAttrs = na;
  if constexpr (CAN_SPLIT) {
    if (pa.splitk > 1) {
      YB_CUDA(cudaLaunchKernelEx(&lc, conv_tc_persist_kernel<BN, BK, PAIR, true>, op.tmA, op.tmB[bn_index(PAIR ? BN / 2 : BN)], op.tmOut, op.tmRes, a, pa));
      return YB_OK;
    }
  }
  YB_CUDA(cudaLaunchKernelEx(&lc, kern, op.tmA, op.tmB[bn_index(PAIR ? BN / 2 : BN)], op.tmOut, op.tmRes, a, pa));
  return YB_OK;
}

static int dispatch_conv_tcp(cudaStream_t st, Op& op, const ConvArgs& a, const LaunchEnv& env, const ConvCfg& cfg) {
  if (op.cout_pad > CONV_TCP_MAX_COUT_PAD) return fail(YB_ERR_INVALID, "tcgen05 conv supports at most %d output channels", CONV_TCP_MAX_COUT_PAD);
  if (cfg.pair) {
    if (ceil_div(a.M, 128) < 2) return fail(YB_ERR_INVALID, "CTA pairs need two M tiles");
#define YB_CASE(BN_, BK_) \
    if (cfg.bn == BN_ && op.bk == BK_) return launch_conv_tcp<BN_, BK_, true>(st, op, a, env, cfg);
    YB_CASE(256, 64) YB_CASE(128, 64) YB_CASE(64, 64) YB_CASE(128, 32) YB_CASE(64, 32)
#undef YB_CASE
    return fail(YB_ERR_INVALID, "no CTA-pair conv instantiation for BN=%d BK=%d", cfg.bn, op.bk);
  }
#define YB_CASE(BN_, BK_) \
  if (cfg.bn == BN_ && op.bk == BK_) return launch_conv_tcp<BN_, BK_, false>(st, op, a, env, cfg);
  YB_CASE(256, 64) YB_CASE(128, 64) YB_CASE(64, 64) YB_CASE(32, 64)
  YB_CASE(128, 32) YB_CASE(64, 32) YB_CASE(32, 32)
#undef YB_CASE
  return fail(YB_ERR_INVALID, "no tcgen05 conv instantiation for BN=%d BK=%d", cfg.bn, op.bk);
}

// Heuristic configuration (used until yb_engine_autotune has measured the alternatives): the widest N tile; CTA pairs
// (cta_group::2) where the mainloop matters: every 3x3 conv (narrow N tiles are MMA-issue bound, wide ones operand-
// bandwidth bound) and the 1x1 convs with N tile 256 and K >= 512.
static ConvCfg default_cfg(const Op& op, int max_batch, bool allow_pair) {
  ConvCfg c;
  c.bn = op.bn_max;
  const long long M = (long long)max_batch * op.Ho * op.Wo;
  const int K = op.ksize * op.kw * op.cin;
  c.pair = (allow_pair && c.bn >= 64 && M > 128 && op.bk == 64 && (op.ksize > 1 || (c.bn == 256 && K >= 512))) ? 1 : 0;
  return c;
}

// Every launch configuration worth measuring for this conv.
static std::vector<ConvCfg> candidate_cfgs(const Op& op, int n, bool allow_pair) {
  std::vector<ConvCfg> v;
  const long long M = (long long)n * op.Ho * op.Wo;
  for (int bn = op.bn_max; bn >= 32 && bn >= op.bn_max / 4; bn >>= 1) {
    for (int pair = 0; pair <= 1; ++pair) {
      if (pair && !(allow_pair && bn >= 64 && M > 128)) continue;
      for (int bstat = 0; bstat <= 1; ++bstat) {
        if (bstat && round_up(op.cout, bn) != bn) continue;
        for (int te = 0; te <= 1; ++te) {
          if (te && !op.tma_epi) continue;
          ConvCfg c;
          c.bn = bn; c.pair = pair; c.bstat = bstat; c.tma_epi = te;
          v.push_back(c);
        }
      }
    }
  }
  return v;
}

// The fused stem reads the network input through a 3-D tiled tensor map over (W*3 elements, H, N): one per input
// address and element type (the two staging slots, or whatever device pointer the caller passes), cached.
static int stem_input_map(yb_engine* e, const void* ptr, int dtype, const CUtensorMap** out) {
  for (auto& c : e->stem_maps) if (c.ptr == ptr && c.dtype == dtype) { *out = &c.map; return YB_OK; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0)
    return fail(YB_ERR_INVALID, "the fused first layers read the input by TMA: a device input must be 16-byte aligned (or set YB_FUSE_STEM=0)");
  const size_t es = dtype == YB_U8 ? 1 : 4;
  if (e->stem_maps.size() >= 64) e->stem_maps.clear();          // callers cycling through many device buffers
  e->stem_maps.emplace_back();
  auto& c = e->stem_maps.back();
  c.ptr = ptr; c.dtype = dtype;
  cuuint64_t dims[3] = {(cuuint64_t)e->W * 3, (cuuint64_t)e->H, (cuuint64_t)e->max_batch};
  cuuint64_t strides[2] = {(cuuint64_t)e->W * 3 * es, (cuuint64_t)e->H * e->W * 3 * es};
  cuuint32_t box[3] = {(cuuint32_t)(dtype == YB_U8 ? STEM_RAW_ROW_U8 : STEM_PATCH_PITCH), (cuuint32_t)STEM_PATCH_ROWS, 1};
  cuuint32_t est[3] = {1, 1, 1};
  CUresult r = g_encode_tiled(&c.map, dtype == YB_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr),
                              dims, strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { e->stem_maps.pop_back(); return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled(network input) failed (%d)", (int)r); }
  *out = &c.map;
  return YB_OK;
}

// conv_fused.cuh: one persistent CTA per SM over 8 x 16 output tiles
static int launch_fused(yb_engine* e, Op& op, int n) {
  const Op& pp = e->ops[op.fuse_kind == FUSE_CONVPOOL ? 0 : op.fuse_src];      // (unused for conv + pool)
  FuseArgs a;
  memset(&a, 0, sizeof(a));
  a.n_img = n; a.Ho = op.Ho; a.Wo = op.Wo; a.H = pp.in.h; a.W = pp.in.w;
  a.scale1 = pp.d_scale; a.shift1 = pp.d_shift;
  if ((int)op.h_scale.size() < FUSE_COUT || (int)op.h_shift.size() < FUSE_COUT) return fail(YB_ERR_STATE, "layer %d: fused conv launched before its weights were loaded", op.layer);
  memcpy(a.scale2, op.h_scale.data(), sizeof(a.scale2));
  memcpy(a.shift2, op.h_shift.data(), sizeof(a.shift2));
  a.leaky1 = pp.leaky; a.leaky2 = op.leaky;
  a.tiles_h = op.Ho / FUSE_TH; a.tiles_w = op.Wo / FUSE_TW; a.n_tiles = n * a.tiles_h * a.tiles_w;
  a.dbg = e->dbg_counters;
  const int grid = std::min(a.n_tiles, e->num_sms);
  const CUtensorMap& tmB = op.tmB[bn_index(FUSE_COUT)];
  if (op.fuse_kind == FUSE_CONVPOOL) {
    a.H = op.Ho; a.W = op.Wo;
    convpool_fused_kernel<<<grid, CP_THREADS, FUSE_SMEM_CONVPOOL, e->stream>>>(op.tmIn4, tmB, op.tmOut4, a);
  } else if (op.fuse_kind == FUSE_STEM) {
    a.w1 = pp.d_wt32;
    const CUtensorMap* tmIn = nullptr;
    YB_TRY(stem_input_map(e, e->cur_input, e->cur_input_dtype, &tmIn));
    if (e->cur_input_dtype == YB_U8) stem_fused_kernel<true><<<grid, FUSE_THREADS, fuse_smem_stem<true>(), e->stream>>>(*tmIn, tmB, op.tmOut4, a);
    else stem_fused_kernel<false><<<grid, FUSE_THREADS, fuse_smem_stem<false>(), e->stream>>>(*tmIn, tmB, op.tmOut4, a);
  } else {
    BlockArgs ba;
    ba.f = a;
    if ((int)pp.h_scale.size() < FUSE_CMID || (int)pp.h_shift.size() < FUSE_CMID) return fail(YB_ERR_STATE, "layer %d: fused conv launched before its weights were loaded", pp.layer);
    memcpy(ba.scale1, pp.h_scale.data(), sizeof(ba.scale1));
    memcpy(ba.shift1, pp.h_shift.data(), sizeof(ba.shift1));
    block_fused_kernel<<<grid, BLK_THREADS, FUSE_SMEM_BLOCK, e->stream>>>(op.tmIn4, op.tmW1, tmB, op.tmOut4, ba);
  }
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

template <int BN, int BK, bool PAIR>
static int set_tcp_attr() {
  YB_CUDA(cudaFuncSetAttribute(conv_tc_persist_kernel<BN, BK, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_TCP_SMEM_MAX));
  if constexpr (BK == 64 && BN >= 64)
    YB_CUDA(cudaFuncSetAttribute(conv_tc_persist_kernel<BN, BK, PAIR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_TCP_SMEM_MAX));
  return YB_OK;
}
// Per-device function attributes (opt-in shared-memory limits) and occupancy figures, set when an engine is created on
// the device -- no process-global bookkeeping, so engines on several devices can be driven from several threads.
// Split-K workspace: 2 * capacity work units of the widest pair tile (2 x 128 x 256 floats each), counters for every
// (tile, pair rank, epilogue warp).
static int set_split_k(yb_engine* e, bool on) {
  if (on && !e->splitk_ws) {
    const size_t units = 2 * (size_t)e->num_sms;
    e->splitk_ws_bytes = units * 2 * 128 * 256 * 4;
    e->splitk_cnt_n = units * 2 * CONV_TCP_EPI_WARPS;
    YB_CUDA(cudaMalloc(&e->splitk_ws, e->splitk_ws_bytes));
    YB_CUDA(cudaMalloc(&e->splitk_cnt, e->splitk_cnt_n * sizeof(unsigned int)));
    YB_CUDA(cudaMemset(e->splitk_cnt, 0, e->splitk_cnt_n * sizeof(unsigned int)));
  }
  e->split_k = on ? 1 : 0;
  return YB_OK;
}

static int set_kernel_attrs(yb_engine* e) {
  YB_TRY((set_tcp_attr<256, 64, true>())); YB_TRY((set_tcp_attr<128, 64, true>())); YB_TRY((set_tcp_attr<64, 64, true>()));
  YB_TRY((set_tcp_attr<128, 32, true>())); YB_TRY((set_tcp_attr<64, 32, true>()));
  YB_TRY((set_tcp_attr<256, 64, false>())); YB_TRY((set_tcp_attr<128, 64, false>())); YB_TRY((set_tcp_attr<64, 64, false>()));
  YB_TRY((set_tcp_attr<32, 64, false>())); YB_TRY((set_tcp_attr<128, 32, false>())); YB_TRY((set_tcp_attr<64, 32, false>()));
  YB_TRY((set_tcp_attr<32, 32, false>()));
  YB_CUDA(cudaFuncSetAttribute(stem_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fuse_smem_stem<true>()));
  YB_CUDA(cudaFuncSetAttribute(stem_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fuse_smem_stem<false>()));
  YB_CUDA(cudaFuncSetAttribute(block_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSE_SMEM_BLOCK));
  YB_CUDA(cudaFuncSetAttribute(convpool_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSE_SMEM_CONVPOOL));
  const int smem_first = (FIRST_ROWS + 2) * FIRST_ROW_ELEMS(e->W) * 2, smem_pool = (2 * FIRST_ROWS + 2) * FIRST_ROW_ELEMS(e->W) * 2;
  if (smem_first <= 200 * 1024) {
    YB_CUDA(cudaFuncSetAttribute(conv_first_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    YB_CUDA(cudaFuncSetAttribute(conv_first_mma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->first_resident[0], conv_first_mma_kernel<false, false>, 256, smem_first));
    YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->first_resident[1], conv_first_mma_kernel<true, false>, 256, smem_first));
  }
  if (smem_pool <= 200 * 1024) {
    YB_CUDA(cudaFuncSetAttribute(conv_first_mma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    YB_CUDA(cudaFuncSetAttribute(conv_first_mma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->first_resident[2], conv_first_mma_kernel<false, true>, 256, smem_pool));
    YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->first_resident[3], conv_first_mma_kernel<true, true>, 256, smem_pool));
  }
  for (int i = 0; i < 4; ++i) if (e->first_resident[i] < 1) e->first_resident[i] = 1;
  return YB_OK;
}

static int run_op(yb_engine* e, Op& op, int n, const ConvCfg* cfg_override = nullptr) {
  cudaStream_t st = e->stream;
  if (op.skip) return YB_OK;                    // computed inside its consumer (Op::fuse_kind)
  if (op.kind == OP_CONV && op.path == PATH_FUSED) return launch_fused(e, op, n);
  if (op.kind == OP_CONV) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.M = n * op.Ho * op.Wo; a.Ho = op.Ho; a.Wo = op.Wo; a.cout = op.cout; a.taps = op.ksize * op.kw; a.ksize = op.ksize;
    a.kw = op.kw; a.stride_w = op.stride_w; a.pad_w = op.pad_w;
    a.conv_stride = op.stride; a.pad = op.pad; a.scale = op.d_scale; a.shift = op.d_shift; a.leaky = op.leaky;
    a.res = op.has_res ? reinterpret_cast<const __nv_bfloat16*>(view_ptr(e, op.in2)) : nullptr;
    a.res_ld = op.in2.ld;
    a.out = view_ptr(e, op.out); a.out_ld = op.out.ld; a.out_f32 = op.out.f32 ? 1 : 0; a.out_mode = op.out_mode;
    InView in{view_ptr(e, op.in), op.in.ld, op.in.h, op.in.w, op.in.c};
    int path = op.path;
    if (path == PATH_TC && e->conv_impl == 1) path = PATH_SIMT;
    if (path == PATH_TC) {
      a.kc_blocks = op.cin / op.bk;
      a.im2col = !(op.ksize == 1 && op.stride == 1);
      const LaunchEnv env{e->device, e->num_sms, e->b_stationary, e->tma_epilogue, e->pdl, e->ablate, e->dbg_counters, e->solo_issue,
                          e->split_k, e->splitk_ws, e->splitk_ws_bytes, e->splitk_cnt, e->splitk_cnt_n};
      ConvCfg cfg = cfg_override ? *cfg_override : op.cfg;
      if (cfg.pair && ceil_div(a.M, 128) < 2) cfg.pair = 0;          // a batch too small to form a pair of M tiles
      YB_TRY(dispatch_conv_tcp(st, op, a, env, cfg));
    } else if (path == PATH_DIRECT) {
      const bool u8 = e->cur_input_dtype == YB_U8 && op.in.buf == -2;
      const bool pool = op.out_mode == OUT_POOL2;               // the 2x2 max pool behind the conv is applied in registers
      const int rows_per_block = pool ? 2 * FIRST_ROWS : FIRST_ROWS;
      const int smem_first = (rows_per_block + 2) * FIRST_ROW_ELEMS(op.in.w) * 2;
      const bool mma_ok = op.ksize == 3 && op.cin == 3 && op.stride == 1 && op.cout % 32 == 0 && !op.out.f32 && !op.has_res &&
                          (op.out_mode == OUT_PLAIN || pool) && op.out.ld % 8 == 0 && op.out.coff % 8 == 0 && e->conv_impl == 0 &&
                          op.in.ld == 3 && (op.in.w * 3) % 4 == 0 && smem_first <= 200 * 1024;
      if (pool && !mma_ok) return fail(YB_ERR_INVALID, "layer %d: the fused conv + max-pool needs the tensor-core first-conv kernel", op.layer);
      if (mma_ok) {
        const int groups = n * ceil_div(op.Ho, rows_per_block);     // stride 1: output rows == input rows
        auto launch = [&](auto kern) -> int {
          const int res = e->first_resident[(pool ? 2 : 0) + (u8 ? 1 : 0)];          // co-resident blocks per SM (set_kernel_attrs)
          // exactly one resident wave: the blocks walk the row groups with a grid stride, a partial second wave
          // would run at a fraction of the occupancy
          dim3 grid(std::min(groups, e->num_sms * res / std::max(1, op.cout / 32)), op.cout / 32);
          if (grid.x < 1) grid.x = 1;
          kern<<<grid, 256, smem_first, st>>>(in, op.d_wt32, a, e->d_u8_lut, groups);
          YB_CUDA(cudaGetLastError());
          return YB_OK;
        };
        if (pool) return u8 ? launch(conv_first_mma_kernel<true, true>) : launch(conv_first_mma_kernel<false, true>);
        if (u8) return launch(conv_first_mma_kernel<true, false>);
        return launch(conv_first_mma_kernel<false, false>);
      }
      const int smem = a.taps * op.cin * 32 * 4;
      dim3 grid(ceil_div(a.M, 128), ceil_div(op.cout, 32));
      if (e->cur_input_dtype == YB_U8 && op.in.buf == -2)
        conv_direct_kernel<true><<<grid, 128, smem, st>>>(in, op.d_wt32, a, e->d_u8_lut);
      else
        conv_direct_kernel<false><<<grid, 128, smem, st>>>(in, op.d_wt32, a, e->d_u8_lut);
      YB_CUDA(cudaGetLastError());
    } else {
      const long long total = (long long)a.M * ((op.cout + 3) / 4);
      conv_simt_kernel<<<ceil_div(total, 256), 256, 0, st>>>(in, op.d_wt, a);
      YB_CUDA(cudaGetLastError());
    }
    return YB_OK;
  }
  const __nv_bfloat16* ip = reinterpret_cast<const __nv_bfloat16*>(view_ptr(e, op.in));
  __nv_bfloat16* opn = reinterpret_cast<__nv_bfloat16*>(view_ptr(e, op.out));
  if (op.kind == OP_MAXPOOL) {
    const long long total = (long long)n * op.out.h * op.out.w * (op.in.c / 8);
    maxpool2_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ip, op.in.ld, op.in.h, op.in.w, op.in.c, opn, op.out.ld, total);
  } else if (op.kind == OP_ADD) {
    const long long total = (long long)n * op.out.h * op.out.w * op.out.c;
    add_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ip, op.in.ld, reinterpret_cast<const __nv_bfloat16*>(view_ptr(e, op.in2)),
                                                      op.in2.ld, opn, op.out.ld, op.out.c, total);
  } else if (op.kind == OP_UPSAMPLE) {
    if (op.factor == 2 && op.in.c % 8 == 0 && op.in.ld % 8 == 0 && op.in.coff % 8 == 0 && op.out.ld % 8 == 0 && op.out.coff % 8 == 0) {
      const long long total = (long long)n * op.in.h * op.in.w * (op.in.c / 8);
      upsample2_vec8_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ip, op.in.ld, op.in.h, op.in.w, op.in.c / 8, opn, op.out.ld, total);
    } else {
      const long long total = (long long)n * op.out.h * op.out.w * op.out.c;
      upsample_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ip, op.in.ld, op.in.h, op.in.w, op.in.c, op.factor, opn, op.out.ld, total);
    }
  } else if (op.kind == OP_REORG) {
    const long long total = (long long)n * op.in.h * op.in.w * op.in.c;
    reorg_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ip, op.in.ld, op.in.h, op.in.w, op.in.c, op.factor, opn, op.out.ld, total);
  } else if (op.kind == OP_COPY) {
    const long long total = (long long)n * op.out.h * op.out.w * op.in.c;
    copy_channels_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ip, op.in.ld, opn, op.out.ld, op.in.c, total);
  }
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

// ---- plan compilation ----
static int compile_plan(yb_engine* e) {
  const int n = (int)e->plan.size();
  const std::vector<yb_layer>& P = e->plan;
  e->shape.assign(n, Shape());
  e->view.assign(n, View());
  std::vector<std::vector<int>> consumers(n);
  if (n < 2 || P[0].kind != YB_INPUT) return fail(YB_ERR_INVALID, "plan must start with an input layer");
  for (int i = 0; i < n; ++i) {
    const yb_layer& l = P[i];
    if (l.n_src < 0 || l.n_src > YB_MAX_SRC) return fail(YB_ERR_INVALID, "layer %d: bad n_src", i);
    for (int s = 0; s < l.n_src; ++s) {
      if (l.src[s] < 0 || l.src[s] >= i) return fail(YB_ERR_INVALID, "layer %d: source %d is not an earlier layer", i, l.src[s]);
      consumers[l.src[s]].push_back(i);
    }
    Shape& sh = e->shape[i];
    auto S = [&](int k) -> const Shape& { return e->shape[l.src[k]]; };
    switch (l.kind) {
      case YB_INPUT: sh.h = e->H; sh.w = e->W; sh.c = e->C; break;
      case YB_CONV:
        if (l.n_src != 1 || l.filters <= 0 || l.ksize <= 0 || l.stride <= 0) return fail(YB_ERR_INVALID, "layer %d: bad conv", i);
        if (l.ksize % 2 == 0) return fail(YB_ERR_INVALID, "layer %d: even conv kernels are not supported", i);
        sh.h = l.stride > 1 ? (S(0).h - 1) / l.stride + 1 : S(0).h;
        sh.w = l.stride > 1 ? (S(0).w - 1) / l.stride + 1 : S(0).w;
        sh.c = l.filters;
        break;
      case YB_MAXPOOL:
        if (l.n_src != 1 || l.ksize != 2 || l.stride != 2 || (S(0).h & 1) || (S(0).w & 1) || (S(0).c & 7))
          return fail(YB_ERR_INVALID, "layer %d: only 2x2/2 max pool on even maps with C%%8==0 is supported", i);
        sh.h = S(0).h / 2; sh.w = S(0).w / 2; sh.c = S(0).c;
        break;
      case YB_ROUTE:
        if (l.n_src < 1) return fail(YB_ERR_INVALID, "layer %d: route without sources", i);
        sh.h = S(0).h; sh.w = S(0).w; sh.c = 0;
        for (int s = 0; s < l.n_src; ++s) {
          if (S(s).h != sh.h || S(s).w != sh.w) return fail(YB_ERR_INVALID, "layer %d: route inputs differ in size", i);
          sh.c += S(s).c;
        }
        break;
      case YB_REORG:
        if (l.n_src != 1 || l.stride < 1 || S(0).h % l.stride || S(0).w % l.stride) return fail(YB_ERR_INVALID, "layer %d: bad reorg", i);
        sh.h = S(0).h / l.stride; sh.w = S(0).w / l.stride; sh.c = S(0).c * l.stride * l.stride;
        break;
      case YB_SHORTCUT:
        if (l.n_src != 2 || S(0).h != S(1).h || S(0).w != S(1).w || S(0).c != S(1).c) return fail(YB_ERR_INVALID, "layer %d: bad shortcut", i);
        sh = S(0);
        break;
      case YB_UPSAMPLE:
        if (l.n_src != 1 || l.stride < 1) return fail(YB_ERR_INVALID, "layer %d: bad upsample", i);
        sh.h = S(0).h * l.stride; sh.w = S(0).w * l.stride; sh.c = S(0).c;
        break;
      case YB_YOLO:
        if (l.n_src != 1 || l.n_anchors < 1 || l.n_anchors > YB_MAX_ANCHORS) return fail(YB_ERR_INVALID, "layer %d: bad yolo layer", i);
        if (P[l.src[0]].kind != YB_CONV) return fail(YB_ERR_INVALID, "layer %d: a yolo layer must follow a conv", i);
        if (S(0).c != l.n_anchors * (5 + e->num_classes))
          return fail(YB_ERR_INVALID, "layer %d: head has %d channels, expected %d", i, S(0).c, l.n_anchors * (5 + e->num_classes));
        sh = S(0);
        break;
      case YB_DETECTION: sh = Shape(); break;
      default: return fail(YB_ERR_INVALID, "layer %d: unknown kind %d", i, l.kind);
    }
  }
  // --- fusion decisions ---
  std::vector<int> absorbed_into(n, -1);   // conv i writes directly into the view of layer absorbed_into[i]
  std::vector<int> res_src(n, -1);
  std::vector<int> out_mode(n, OUT_PLAIN);
  std::vector<bool> is_head(n, false);
  for (int i = 0; i < n; ++i) {
    if (P[i].kind != YB_CONV) continue;
    const std::vector<int>& cs = consumers[i];
    bool all_yolo = !cs.empty();
    for (int c : cs) all_yolo = all_yolo && P[c].kind == YB_YOLO;
    if (all_yolo || (cs.empty() && i == n - 1)) { is_head[i] = true; continue; }
    if (cs.size() != 1) continue;
    const int j = cs[0];
    const bool tc_ok = true;
    if (P[j].kind == YB_SHORTCUT && P[j].src[0] == i && P[j].src[1] != i && tc_ok) { absorbed_into[i] = j; res_src[i] = P[j].src[1]; }
    // The replicated 2x2 store is fused into the conv epilogue.  Timed alone its direct stores cost more than a plain
    // TMA epilogue plus the 16-byte-vectorised copy kernel, but inside the PDL-overlapped step the two variants are
    // indistinguishable (14,885 vs 14,904 img/s), so the variant with two launches less stays the default.
    else if (e->fuse_upsample && P[j].kind == YB_UPSAMPLE && P[j].stride == 2) { absorbed_into[i] = j; out_mode[i] = OUT_UPSAMPLE2; }
    else if (P[j].kind == YB_REORG && P[j].stride == 2 && (e->shape[i].h % 2 == 0) && (e->shape[i].w % 2 == 0)) { absorbed_into[i] = j; out_mode[i] = OUT_REORG2; }
    // Darknet-19's first conv + max pool (net/v2.py:20-22): pooled in the registers of the tensor-core first-conv kernel
    else if (P[j].kind == YB_MAXPOOL && P[P[i].src[0]].kind == YB_INPUT && e->C == 3 && P[i].ksize == 3 && P[i].stride == 1 &&
             P[i].filters % 32 == 0 && P[i].batch_norm && e->shape[i].h % 2 == 0 && e->shape[i].w % 2 == 0 && (e->W * 3) % 4 == 0 &&
             (2 * FIRST_ROWS + 2) * FIRST_ROW_ELEMS(e->W) * 2 <= 200 * 1024 &&
             (e->fuse_pool == 2 || (e->fuse_pool == 1 && !e->keep_all))) { absorbed_into[i] = j; out_mode[i] = OUT_POOL2; }
    // Darknet-19's second conv + max pool (net/v2.py:23-24): conv 3x3 32 -> 64 on the tcgen05 consumer of conv_fused.cuh,
    // pooled in the epilogue's registers (convpool_fused_kernel).  The input must be a plain activation of its own.
    else if (P[j].kind == YB_MAXPOOL && P[j].stride == 2 && P[j].ksize == 2 && (P[P[i].src[0]].kind == YB_MAXPOOL || P[P[i].src[0]].kind == YB_CONV) &&
             P[i].ksize == 3 && P[i].stride == 1 && P[i].filters == FUSE_COUT && P[i].batch_norm && e->shape[P[i].src[0]].c == FUSE_CMID &&
             e->shape[i].h % FUSE_TH == 0 && e->shape[i].w % FUSE_TW == 0 &&
             (e->fuse_pool == 2 || (e->fuse_pool == 1 && !e->keep_all))) { absorbed_into[i] = j; out_mode[i] = OUT_POOL2; }
  }
  std::vector<bool> absorbs(n, false);
  for (int i = 0; i < n; ++i) if (absorbed_into[i] >= 0) absorbs[absorbed_into[i]] = true;
  auto root_of = [&](int s) { while (P[s].kind == YB_ROUTE && P[s].n_src == 1) s = P[s].src[0]; return s; };
  // --- concat placement ---
  struct Place { int route = -1, coff = 0; };
  std::vector<Place> placed(n);
  for (int j = 0; j < n; ++j) {
    if (P[j].kind != YB_ROUTE || P[j].n_src < 2) continue;
    int coff = 0;
    for (int s = 0; s < P[j].n_src; ++s) {
      const int r = root_of(P[j].src[s]);
      const int k = P[r].kind;
      const bool placeable = (k == YB_CONV && absorbed_into[r] < 0 && !is_head[r]) || k == YB_MAXPOOL || k == YB_SHORTCUT ||
                             k == YB_UPSAMPLE || k == YB_REORG;
      if (placeable && placed[r].route < 0 && (coff % 8) == 0 && (e->shape[r].c % 8) == 0) { placed[r].route = j; placed[r].coff = coff; }
      coff += e->shape[P[j].src[s]].c;
    }
  }
  // --- views and buffers ---
  auto new_buf = [&](const Shape& sh, int ld, bool f32) {
    Buf b;
    b.bytes = (size_t)e->max_batch * sh.h * sh.w * ld * (f32 ? 4 : 2);
    b.keep = e->keep_all;
    e->bufs.push_back(b);
    return (int)e->bufs.size() - 1;
  };
  std::vector<int> route_buf(n, -1);
  for (int j = 0; j < n; ++j)
    if (P[j].kind == YB_ROUTE && P[j].n_src >= 2) route_buf[j] = new_buf(e->shape[j], round_up(e->shape[j].c, 8), false);
  for (int i = 0; i < n; ++i) {
    const Shape& sh = e->shape[i];
    View v;
    v.h = sh.h; v.w = sh.w; v.c = sh.c;
    const int k = P[i].kind;
    if (k == YB_INPUT) { v.buf = -2; v.ld = sh.c; v.f32 = true; }
    else if (k == YB_DETECTION) { v.buf = -1; }
    else if (k == YB_YOLO) { v = e->view[P[i].src[0]]; }
    else if (k == YB_ROUTE && P[i].n_src == 1) { v = e->view[P[i].src[0]]; }
    else if (k == YB_ROUTE) { v.buf = route_buf[i]; v.ld = round_up(sh.c, 8); v.coff = 0; }
    else if (k == YB_CONV && absorbed_into[i] >= 0) { v.buf = -1; }
    else if (placed[i].route >= 0) { v.buf = route_buf[placed[i].route]; v.ld = round_up(e->shape[placed[i].route].c, 8); v.coff = placed[i].coff; }
    else if (k == YB_CONV && is_head[i]) { v.f32 = true; v.ld = round_up(sh.c, 4); v.buf = new_buf(sh, v.ld, true); e->bufs[v.buf].keep = true; }
    else { v.ld = round_up(sh.c, 8); v.buf = new_buf(sh, v.ld, false); }
    e->view[i] = v;
  }
  // --- ops ---
  auto touch = [&](const View& v, int op_idx, bool write) {
    if (v.buf < 0) return;
    Buf& b = e->bufs[v.buf];
    if (write) b.first = std::min(b.first, op_idx);
    b.first = std::min(b.first, op_idx);
    b.last = std::max(b.last, op_idx);
  };
  for (int i = 0; i < n; ++i) {
    const yb_layer& l = P[i];
    Op op;
    op.layer = i;
    bool emit = false;
    if (l.kind == YB_CONV) {
      const int dst = absorbed_into[i] >= 0 ? absorbed_into[i] : i;
      op.kind = OP_CONV;
      op.in = e->view[l.src[0]];
      if (op.in.buf == -1) return fail(YB_ERR_INVALID, "layer %d: input %d was fused away", i, l.src[0]);
      op.out = e->view[dst];
      op.cin = e->shape[l.src[0]].c; op.cout = l.filters; op.ksize = l.ksize; op.stride = l.stride;
      op.pad = (l.ksize - 1) / 2; op.leaky = l.leaky; op.out_mode = out_mode[i];
      op.kw = op.ksize; op.stride_w = op.stride; op.pad_w = op.pad;
      op.Ho = e->shape[i].h; op.Wo = e->shape[i].w;
      if (res_src[i] >= 0) { op.has_res = 1; op.in2 = e->view[res_src[i]]; if (op.in2.buf == -1 || op.in2.f32) return fail(YB_ERR_INVALID, "layer %d: bad residual source", i); }
      const bool in_bf16 = !op.in.f32;
      const bool tma_ok = in_bf16 && (op.cin % 32 == 0) && (op.in.ld % 8 == 0) && (op.in.coff % 8 == 0) &&
                          (l.ksize == 1 || l.ksize == 3) && (l.stride == 1 || l.stride == 2) &&
                          (op.out.f32 || (op.cout % 8 == 0 && op.out.ld % 8 == 0 && op.out.coff % 8 == 0)) &&
                          (!op.has_res || (op.in2.ld % 8 == 0 && op.in2.coff % 8 == 0));
      // ---- two-layer fusion (conv_fused.cuh): this 3x3 conv over 32 channels also computes the conv that feeds it ----
      const bool fuse_shape = tma_ok && l.ksize == 3 && op.cin == 32 && op.cout == FUSE_COUT && op.out_mode == OUT_PLAIN && !op.out.f32 &&
                              op.Ho % FUSE_TH == 0 && op.Wo % FUSE_TW == 0 && !e->ops.empty() && e->ops.back().kind == OP_CONV &&
                              e->ops.back().layer == l.src[0] && consumers[l.src[0]].size() == 1 && !e->ops.back().skip &&
                              e->ops.back().out_mode == OUT_PLAIN && !e->ops.back().has_res && e->ops.back().cout == FUSE_CMID;
      if (fuse_shape) {
        Op& pp = e->ops.back();
        const bool stem = l.stride == 2 && !op.has_res && pp.path == PATH_DIRECT && pp.ksize == 3 && pp.stride == 1 && pp.cin == 3 &&
                          pp.in.buf == -2 && pp.in.ld == 3 && pp.in.h == 2 * op.Ho && pp.in.w == 2 * op.Wo &&
                          (e->fuse_stem == 2 || (e->fuse_stem == 1 && !e->keep_all));
        const bool block = l.stride == 1 && op.has_res && pp.path == PATH_TC && pp.ksize == 1 && pp.stride == 1 && pp.cin == 64 &&
                           !pp.px_pair && !pp.in.f32 && pp.in.buf >= 0 && pp.in.ld % 8 == 0 && pp.in.coff % 8 == 0 &&
                           op.in2.buf == pp.in.buf && op.in2.coff == pp.in.coff && op.in2.ld == pp.in.ld &&
                           (e->fuse_block == 2 || (e->fuse_block == 1 && !e->keep_all));
        if (stem || block) {
          op.fuse_kind = stem ? FUSE_STEM : FUSE_BLOCK;
          op.fuse_src = (int)e->ops.size() - 1;
          pp.skip = true;
          if (pp.out.buf >= 0) { e->bufs[pp.out.buf].first = INT_MAX; e->bufs[pp.out.buf].last = -1; }   // never materialised
          e->view[pp.layer].buf = -1;
          op.in = pp.in;
        }
      }
      if (op.out_mode == OUT_POOL2 && op.in.buf != -2) {        // conv 32 -> 64 + max pool: convpool_fused_kernel
        if (!(tma_ok && op.cin == FUSE_CMID && op.cout == FUSE_COUT && op.in.buf >= 0 && op.in.ld == FUSE_CMID && op.in.coff == 0 &&
              !op.out.f32 && !op.has_res))
          return fail(YB_ERR_INVALID, "layer %d: conv + max pool fusion needs a plain 32-channel bf16 input (pitch %d, offset %d)", i, op.in.ld, op.in.coff);
        op.fuse_kind = FUSE_CONVPOOL;
      }
      if (op.fuse_kind) {
        op.path = PATH_FUSED;
        op.bk = FUSE_CMID; op.bn_max = FUSE_COUT;
        op.cfg = ConvCfg(); op.cfg.bn = FUSE_COUT;
      } else
      if (e->pixel_pairs && tma_ok && op.cin == 32 && l.stride == 1 && op.out_mode == OUT_PLAIN && !op.out.f32 &&
          op.cout % 32 == 0 && op.cout <= 128 && op.Wo % 2 == 0 && op.in.w == op.Wo && op.in.ld == op.cin && op.in.coff == 0 &&
          op.out.ld == op.cout && op.out.coff == 0 && (!op.has_res || (op.in2.ld == op.cout && op.in2.coff == 0 && op.in2.w == op.Wo))) {
        op.px_pair = 1;
        op.Wo /= 2;
        op.in.w /= 2; op.in.c *= 2; op.in.ld *= 2;
        op.out.w /= 2; op.out.c *= 2; op.out.ld *= 2;
        if (op.has_res) { op.in2.w /= 2; op.in2.c *= 2; op.in2.ld *= 2; }
        op.cin *= 2; op.cout *= 2;
      }
      if (e->pixel_pairs >= 2 && !op.px_pair && !op.fuse_kind && tma_ok && op.cin == 32 && l.ksize == 3 && l.stride == 2 && op.in.w % 2 == 0 &&
          op.in.w == 2 * op.Wo && op.in.ld == op.cin && op.in.coff == 0) {
        op.px_pair = 2;
        op.in.w /= 2; op.in.c *= 2; op.in.ld *= 2;
        op.cin *= 2;
        op.kw = 2; op.stride_w = 1; op.pad_w = 1;
      }
      if (op.fuse_kind) {
        // path, tile and weight packing were fixed above (plain [cout][9*32] bf16 weights, N tile 64, BK 32)
      } else if (op.in.f32 || op.in.buf == -2) {
        if (op.cin > 8) return fail(YB_ERR_INVALID, "layer %d: fp32 conv input with %d channels is not supported", i, op.cin);
        op.path = PATH_DIRECT;
      } else if (tma_ok) {
        op.path = PATH_TC;
        op.bk = (op.cin % 64 == 0) ? 64 : 32;
        int bn = 32;
        while (bn < e->bn_max && bn < op.cout) bn <<= 1;
        if (op.bk == 32 && bn > 128) bn = 128;
        op.bn_max = bn;
        op.cfg = default_cfg(op, e->max_batch, e->cta_pairs);
      } else {
        // No silent detour: the CUDA-core kernel is a debug cross-check, an order of magnitude slower than the tcgen05 path.
        op.path = PATH_SIMT;
        fprintf(stderr, "libyolo_b200: warning: layer %d (conv %dx%d s%d, %d -> %d channels, input pitch %d, channel offset %d) does not "
                        "meet the TMA alignment rules (channels %% 32, pitches and offsets %% 8) and runs on the CUDA-core conv kernel\n",
                i, l.ksize, l.ksize, l.stride, op.cin, op.cout, op.in.ld, op.in.coff);
      }
      op.cout_pad = round_up(op.cout, (op.path == PATH_TC || op.path == PATH_FUSED) ? op.bn_max : 4);
      emit = true;
    } else if (l.kind == YB_MAXPOOL && !absorbs[i]) {
      op.kind = OP_MAXPOOL; op.in = e->view[l.src[0]]; op.out = e->view[i]; emit = true;
      if (op.in.f32 || op.in.buf < 0) return fail(YB_ERR_INVALID, "layer %d: max pool needs a bf16 activation input", i);
    } else if (l.kind == YB_SHORTCUT && !absorbs[i]) {
      op.kind = OP_ADD; op.in = e->view[l.src[0]]; op.in2 = e->view[l.src[1]]; op.out = e->view[i]; emit = true;
      if (op.in.buf < 0 || op.in2.buf < 0 || op.in.f32 || op.in2.f32) return fail(YB_ERR_INVALID, "layer %d: bad shortcut operands", i);
    } else if (l.kind == YB_UPSAMPLE && !absorbs[i]) {
      op.kind = OP_UPSAMPLE; op.in = e->view[l.src[0]]; op.out = e->view[i]; op.factor = l.stride; emit = true;
      if (op.in.buf < 0 || op.in.f32) return fail(YB_ERR_INVALID, "layer %d: bad upsample operand", i);
    } else if (l.kind == YB_REORG && !absorbs[i]) {
      op.kind = OP_REORG; op.in = e->view[l.src[0]]; op.out = e->view[i]; op.factor = l.stride; emit = true;
      if (op.in.buf < 0 || op.in.f32) return fail(YB_ERR_INVALID, "layer %d: bad reorg operand", i);
    } else if (l.kind == YB_ROUTE && l.n_src >= 2) {
      int coff = 0;
      for (int s = 0; s < l.n_src; ++s) {
        const int r = root_of(l.src[s]);
        if (!(placed[r].route == i && placed[r].coff == coff)) {
          Op cp;
          cp.kind = OP_COPY; cp.layer = i; cp.in = e->view[l.src[s]]; cp.out = e->view[i];
          cp.out.coff += coff; cp.out.c = cp.in.c;
          if (cp.in.buf < 0 || cp.in.f32) return fail(YB_ERR_INVALID, "layer %d: cannot concatenate source %d", i, l.src[s]);
          const int idx = (int)e->ops.size();
          touch(cp.in, idx, false); touch(cp.out, idx, true);
          e->ops.push_back(cp);
        }
        coff += e->shape[l.src[s]].c;
      }
    }
    if (emit) {
      const int idx = (int)e->ops.size();
      touch(op.in, idx, false);
      if (op.kind == OP_ADD || op.has_res) touch(op.in2, idx, false);
      touch(op.out, idx, true);
      e->ops.push_back(op);
    }
  }
  // --- heads ---
  int row = 0;
  for (int i = 0; i < n; ++i) {
    if (P[i].kind != YB_YOLO) continue;
    Head hd;
    hd.yolo_layer = i; hd.conv_layer = P[i].src[0]; hd.view = e->view[hd.conv_layer];
    if (!hd.view.f32 || hd.view.buf < 0) return fail(YB_ERR_INVALID, "layer %d: yolo input must be a head conv consumed only by yolo layers", i);
    hd.h = e->shape[i].h; hd.w = e->shape[i].w; hd.na = P[i].n_anchors; hd.row_begin = row;
    memcpy(hd.anchors, P[i].anchors, sizeof(float) * 2 * hd.na);
    row += hd.h * hd.w * hd.na;
    e->heads.push_back(hd);
  }
  if (e->heads.empty()) return fail(YB_ERR_INVALID, "plan has no yolo layer (append one for YOLOv2)");
  if ((int)e->heads.size() > POST_MAX_SCALES) return fail(YB_ERR_INVALID, "too many yolo layers");
  e->rows = row;
  e->box_len = 5 + e->num_classes;
  if (e->rows > 65536 * 16) return fail(YB_ERR_INVALID, "too many head rows per image");
  // --- arena: first-fit over live intervals ---
  struct Live { size_t off, bytes; int last; };
  std::vector<Live> live;
  std::vector<int> order_idx;
  for (int b = 0; b < (int)e->bufs.size(); ++b) if (e->bufs[b].last >= 0) order_idx.push_back(b);
  std::sort(order_idx.begin(), order_idx.end(), [&](int a, int b) { return e->bufs[a].first < e->bufs[b].first; });
  size_t top = 0;
  for (int b : order_idx) {
    Buf& B = e->bufs[b];
    const size_t need = (B.bytes + 1023) & ~(size_t)1023;
    std::vector<Live> still;
    for (const Live& l : live) if (l.last >= B.first) still.push_back(l);
    live.swap(still);
    std::sort(live.begin(), live.end(), [](const Live& a, const Live& b) { return a.off < b.off; });
    size_t off = 0;
    for (const Live& l : live) {
      if (off + need <= l.off) break;
      off = std::max(off, l.off + l.bytes);
    }
    B.off = off;
    live.push_back(Live{off, need, B.keep ? INT_MAX : B.last});
    top = std::max(top, off + need);
  }
  e->arena_bytes = top + 1024;
  return YB_OK;
}

static int build_tensor_maps(yb_engine* e) {
  for (Op& op : e->ops) {
    if (op.kind == OP_CONV && op.path == PATH_FUSED) {
      const int K = 9 * FUSE_CMID;
      memset(op.tmB, 0, sizeof(op.tmB));
      YB_TRY(make_tiled_map(&op.tmB[bn_index(FUSE_COUT)], op.d_wt, op.cout_pad, K, K, FUSE_COUT, FUSE_CMID));
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r;
      if (op.fuse_kind == FUSE_CONVPOOL) {
        // pooled output (C, Wo/2, Ho/2, N), one box of 8 pixels x 64 channels per epilogue warp; input (C = 32, W, H, N), one
        // 10 x 18 patch per tile, 64-byte rows
        cuuint64_t odims[4] = {(cuuint64_t)op.cout, (cuuint64_t)(op.Wo / 2), (cuuint64_t)(op.Ho / 2), (cuuint64_t)e->max_batch};
        cuuint64_t ostr[3] = {(cuuint64_t)op.out.ld * 2, (cuuint64_t)(op.Wo / 2) * op.out.ld * 2, (cuuint64_t)(op.Ho / 2) * (op.Wo / 2) * op.out.ld * 2};
        cuuint32_t obox[4] = {(cuuint32_t)CP_EPI_COLS, 8, 1, 1};          // per epilogue warp: 8 pooled pixels x 16 channels
        r = g_encode_tiled(&op.tmOut4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, view_ptr(e, op.out), odims, ostr, obox, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled(pooled output) failed (%d) for layer %d", (int)r, op.layer);
        cuuint64_t idims[4] = {(cuuint64_t)FUSE_CMID, (cuuint64_t)op.Wo, (cuuint64_t)op.Ho, (cuuint64_t)e->max_batch};
        cuuint64_t istr[3] = {(cuuint64_t)op.in.ld * 2, (cuuint64_t)op.Wo * op.in.ld * 2, (cuuint64_t)op.Ho * op.Wo * op.in.ld * 2};
        cuuint32_t ibox[4] = {(cuuint32_t)FUSE_CMID, (cuuint32_t)BLOCK_PATCH_W, (cuuint32_t)BLOCK_PATCH_H, 1};
        r = g_encode_tiled(&op.tmIn4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, view_ptr(e, op.in), idims, istr, ibox, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled(conv + pool input) failed (%d) for layer %d", (int)r, op.layer);
        continue;
      }
      cuuint64_t dims[4] = {(cuuint64_t)op.cout, (cuuint64_t)op.Wo, (cuuint64_t)op.Ho, (cuuint64_t)e->max_batch};
      cuuint64_t strides[3] = {(cuuint64_t)op.out.ld * 2, (cuuint64_t)op.Wo * op.out.ld * 2, (cuuint64_t)op.Ho * op.Wo * op.out.ld * 2};
      // stem: one box of [2 rows][16 cols][64 channels] per epilogue warp (128-byte rows, SWIZZLE_128B); block: sixteen
      // epilogue warps, [2 rows][16 cols][16 channels] each, dense 32-byte rows
      const bool blk = op.fuse_kind == FUSE_BLOCK;
      cuuint32_t box[4] = {(cuuint32_t)(blk ? BLK_EPI_COLS : FUSE_COUT), (cuuint32_t)FUSE_TW, 2, 1};
      r = g_encode_tiled(&op.tmOut4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, view_ptr(e, op.out), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, blk ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled(4-D output) failed (%d) for layer %d", (int)r, op.layer);
      memset(&op.tmIn4, 0, sizeof(op.tmIn4));
      if (op.fuse_kind == FUSE_BLOCK) {       // the producer's input patch (its centre is the residual): 10 x 18 pixels x 64 channels
        cuuint64_t idims[4] = {64, (cuuint64_t)op.Wo, (cuuint64_t)op.Ho, (cuuint64_t)e->max_batch};
        cuuint64_t istrides[3] = {(cuuint64_t)op.in2.ld * 2, (cuuint64_t)op.Wo * op.in2.ld * 2, (cuuint64_t)op.Ho * op.Wo * op.in2.ld * 2};
        cuuint32_t ibox[4] = {64, (cuuint32_t)BLOCK_PATCH_W, (cuuint32_t)BLOCK_PATCH_H, 1};
        r = g_encode_tiled(&op.tmIn4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, view_ptr(e, op.in2), idims, istrides, ibox, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(YB_ERR_CUDA, "cuTensorMapEncodeTiled(4-D block input) failed (%d) for layer %d", (int)r, op.layer);
        const Op& pp = e->ops[op.fuse_src];
        YB_TRY(make_tiled_map(&op.tmW1, pp.d_wt, pp.cout_pad, 64, 64, FUSE_CMID, 64));
      }
      continue;
    }
    if (op.kind != OP_CONV || op.path != PATH_TC) continue;
    const void* in_ptr = view_ptr(e, op.in);
    if (op.ksize == 1 && op.stride == 1) {
      YB_TRY(make_tiled_map(&op.tmA, in_ptr, (long long)e->max_batch * op.in.h * op.in.w, op.cin, op.in.ld, 128, op.bk));
    } else {
      YB_TRY(make_im2col_map(&op.tmA, in_ptr, e->max_batch, op.in.h, op.in.w, op.cin, op.in.ld, op.ksize, op.stride, op.pad, op.bk,
                             op.kw, op.stride_w, op.pad_w));
    }
    const int K = op.ksize * op.kw * op.cin;
    memset(op.tmB, 0, sizeof(op.tmB));
    for (int bn = 32; bn <= op.bn_max; bn <<= 1) YB_TRY(make_tiled_map(&op.tmB[bn_index(bn)], op.d_wt, op.cout_pad, K, K, bn, op.bk));
    // TMA epilogue (plain bf16 outputs): 32 rows x 32 channels per store, 64-byte swizzle
    op.tma_epi = false;
    memset(&op.tmOut, 0, sizeof(op.tmOut));
    memset(&op.tmRes, 0, sizeof(op.tmRes));
    if (op.out_mode == OUT_PLAIN && !op.out.f32) {
      const long long rows = (long long)e->max_batch * op.Ho * op.Wo;
      YB_TRY(make_map_2d(&op.tmOut, view_ptr(e, op.out), rows, op.cout, op.out.ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
      if (op.has_res) YB_TRY(make_map_2d(&op.tmRes, view_ptr(e, op.in2), rows, op.cout, op.in2.ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
      op.tma_epi = true;
    } else if (op.out_mode == OUT_PLAIN && op.out.f32 && !op.has_res && op.out.ld % 4 == 0 && op.out.coff == 0) {
      // fp32 heads: 32 rows x 16 floats per store (the padded channels up to ld are written as well, like the direct path)
      const long long rows = (long long)e->max_batch * op.Ho * op.Wo;
      YB_TRY(make_map_2d_f32(&op.tmOut, view_ptr(e, op.out), rows, op.out.ld, op.out.ld, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B));
      op.tma_epi = true;
    }
  }
  return YB_OK;
}

static int ensure_scratch(yb_engine* e, size_t floats) {
  if (floats <= e->scratch_floats) return YB_OK;
  if (e->scratch_f32) cudaFree(e->scratch_f32);
  e->scratch_f32 = nullptr; e->scratch_floats = 0;
  YB_CUDA(cudaMalloc(&e->scratch_f32, floats * 4));
  e->scratch_floats = floats;
  return YB_OK;
}

static void engine_scales(const yb_engine* e, ScaleDesc* sc) {
  for (size_t i = 0; i < e->heads.size(); ++i) {
    const Head& hd = e->heads[i];
    ScaleDesc& s = sc[i];
    s.base = reinterpret_cast<const float*>(view_ptr(e, hd.view));
    s.cell_stride = hd.view.ld;
    s.img_stride = (long long)hd.h * hd.w * hd.view.ld;
    s.h = hd.h; s.w = hd.w; s.na = hd.na; s.row_begin = hd.row_begin;
    memcpy(s.anchors, hd.anchors, sizeof(float) * 2 * hd.na);
  }
}

static int set_device(int device) {
  int count = 0;
  cudaError_t er = cudaGetDeviceCount(&count);
  if (er != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(YB_ERR_CUDA, "no usable CUDA device (%s); libyolo_b200 has no CPU fallback", er == cudaSuccess ? "device count is 0" : cudaGetErrorString(er));
  }
  if (device < 0 || device >= count) return fail(YB_ERR_INVALID, "device %d out of range (%d visible)", device, count);
  YB_CUDA(cudaSetDevice(device));
  return YB_OK;
}

}  // namespace yb

// ================================================================================================
// autotuner
// ================================================================================================
// Average device time of `reps` back-to-back launches of one op (after one untimed launch), CUDA events on the engine's stream.
static int time_op(yb_engine* e, Op& op, int n, const ConvCfg* cfg, int reps, float* ms_out) {
  YB_TRY(run_op(e, op, n, cfg));
  YB_CUDA(cudaEventRecord(e->marks[6], e->stream));
  for (int r = 0; r < reps; ++r) YB_TRY(run_op(e, op, n, cfg));
  YB_CUDA(cudaEventRecord(e->marks[7], e->stream));
  YB_CUDA(cudaEventSynchronize(e->marks[7]));
  float ms = 0.f;
  YB_CUDA(cudaEventElapsedTime(&ms, e->marks[6], e->marks[7]));
  *ms_out = ms / (float)reps;
  return YB_OK;
}

extern "C" {
static int forward_impl(yb_engine* e, const void* images, int dtype, int mem, int n, cudaEvent_t* evs, int* n_ev);
}

// Makes sure every activation buffer holds finite, realistic data for a batch of n (op timings on zeros or on
// uninitialised memory would not be representative): one forward over pseudo-random uint8 images.
static int prime_activations(yb_engine* e, int n) {
  const long long total = (long long)n * e->H * e->W * e->C;
  fill_u8_hash_kernel<<<ceil_div(total, 256), 256, 0, e->stream>>>(reinterpret_cast<unsigned char*>(e->input_dev[0]), total, 0x9E3779B9u);
  YB_CUDA(cudaGetLastError());
  YB_TRY(forward_impl(e, e->input_dev[0], YB_U8, YB_MEM_DEVICE, n, nullptr, nullptr));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  return YB_OK;
}

static int autotune(yb_engine* e, int n, int reps) {
  YB_TRY(prime_activations(e, n));
  const bool was_prof = e->prof_on;
  e->prof_on = false;
  struct Done { long long sig[12]; ConvCfg cfg; float best, def; };
  std::vector<Done> done;
  std::string& js = e->tune_report;
  js = "{\"batch\": " + std::to_string(n) + ", \"reps\": " + std::to_string(reps) + ", \"ops\": [";
  bool first_op = true;
  char line[512];
  for (size_t oi = 0; oi < e->ops.size(); ++oi) {
    Op& op = e->ops[oi];
    if (op.kind != OP_CONV || op.path != PATH_TC || op.skip) continue;
    const long long sig[12] = {op.cin, op.cout, op.ksize, op.stride, op.Ho, op.Wo, op.has_res, op.out_mode, op.out.f32 ? 1 : 0,
                               op.leaky, op.in.ld, op.out.ld};
    const Done* hit = nullptr;
    for (const Done& d : done) if (memcmp(d.sig, sig, sizeof(sig)) == 0) { hit = &d; break; }
    if (hit) { op.cfg = hit->cfg; op.tuned_ms = hit->best; op.default_ms = hit->def; continue; }
    const ConvCfg def = default_cfg(op, n, e->cta_pairs);
    float def_ms = 0.f;
    YB_TRY(time_op(e, op, n, &def, reps, &def_ms));
    std::vector<ConvCfg> cands = candidate_cfgs(op, n, e->cta_pairs);
    snprintf(line, sizeof(line), "%s\n {\"op\": %d, \"layer\": %d, \"cin\": %d, \"cout\": %d, \"k\": %d, \"stride\": %d, \"ho\": %d, \"wo\": %d, "
             "\"res\": %d, \"out_mode\": %d, \"default_ms\": %.5f, \"candidates\": [", first_op ? "" : ",", (int)oi, op.layer, op.cin, op.cout,
             op.ksize, op.stride, op.Ho, op.Wo, op.has_res, op.out_mode, def_ms);
    js += line;
    first_op = false;
    float best = 1e30f;
    ConvCfg best_cfg = def;
    for (size_t ci = 0; ci < cands.size(); ++ci) {
      float ms = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {       // best of three short bursts
        float t = 0.f;
        YB_TRY(time_op(e, op, n, &cands[ci], reps, &t));
        ms = std::min(ms, t);
      }
      snprintf(line, sizeof(line), "%s{\"bn\": %d, \"pair\": %d, \"bstat\": %d, \"tma_epi\": %d, \"stages\": %d, \"ksub\": %d, \"ms\": %.5f}", ci ? ", " : "",
               cands[ci].bn, cands[ci].pair, op.launched_bstat, op.launched_tma_epi, op.launched_stages, op.launched_ksub, ms);
      js += line;
      if (ms < best) { best = ms; best_cfg = cands[ci]; }
    }
    // second pass: BK-blocks per stage for the best configuration found so far
    {
      const int num_k = op.ksize * op.kw * (op.cin / op.bk);
      const int ks[] = {1, 2, 3, 4, 6, 9};
      ConvCfg base = best_cfg;
      for (int k : ks) {
        if (k > num_k) continue;
        ConvCfg c = base;
        c.ksub = k;
        float ms = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          float t = 0.f;
          YB_TRY(time_op(e, op, n, &c, reps, &t));
          ms = std::min(ms, t);
        }
        if (op.launched_ksub != k) continue;            // did not fit: resolved to something already measured
        snprintf(line, sizeof(line), ", {\"bn\": %d, \"pair\": %d, \"bstat\": %d, \"tma_epi\": %d, \"stages\": %d, \"ksub\": %d, \"ms\": %.5f}",
                 c.bn, c.pair, op.launched_bstat, op.launched_tma_epi, op.launched_stages, k, ms);
        js += line;
        if (ms < best) { best = ms; best_cfg = c; }
      }
    }
    // keep the heuristic unless a candidate is measurably (>1.5%) faster: avoids flapping on noise
    float def_best = def_ms;
    for (int rep = 0; rep < 2; ++rep) { float t = 0.f; YB_TRY(time_op(e, op, n, &def, reps, &t)); def_best = std::min(def_best, t); }
    if (best > def_best * 0.985f) { best_cfg = def; best = def_best; }
    snprintf(line, sizeof(line), "], \"chosen\": {\"bn\": %d, \"pair\": %d, \"bstat\": %d, \"tma_epi\": %d, \"ksub\": %d, \"ms\": %.5f}}", best_cfg.bn, best_cfg.pair,
             best_cfg.bstat, best_cfg.tma_epi, best_cfg.ksub, best);
    js += line;
    op.cfg = best_cfg; op.tuned_ms = best; op.default_ms = def_best;
    Done d;
    memcpy(d.sig, sig, sizeof(sig)); d.cfg = best_cfg; d.best = best; d.def = def_best;
    done.push_back(d);
  }
  js += "\n]}";
  e->prof_on = was_prof;
  e->last_n = 0;            // activations were clobbered: a forward must precede the next read/detect
  e->detected = false;
  return YB_OK;
}

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

// ---- CRC-32C (Castagnoli), the checksum of TensorFlow's tensor-bundle checkpoints (host only, no GPU needed) ----
static uint32_t g_crc_table[8][256];
static bool g_crc_ready = false;
static void crc32c_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFFu];
  g_crc_ready = true;
}
#if defined(__x86_64__)
__attribute__((target("sse4.2"))) static uint32_t crc32c_hw(uint32_t crc, const unsigned char* p, size_t n) {
  uint64_t c = crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) { c = __builtin_ia32_crc32qi((uint32_t)c, *p++); --n; }
  while (n >= 8) { uint64_t v; memcpy(&v, p, 8); c = __builtin_ia32_crc32di(c, v); p += 8; n -= 8; }
  while (n) { c = __builtin_ia32_crc32qi((uint32_t)c, *p++); --n; }
  return (uint32_t)c;
}
#endif
static uint32_t crc32c_sw(uint32_t crc, const unsigned char* p, size_t n) {   // slicing-by-8
  if (!g_crc_ready) crc32c_init();
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = g_crc_table[7][lo & 0xFF] ^ g_crc_table[6][(lo >> 8) & 0xFF] ^ g_crc_table[5][(lo >> 16) & 0xFF] ^ g_crc_table[4][lo >> 24] ^
          g_crc_table[3][hi & 0xFF] ^ g_crc_table[2][(hi >> 8) & 0xFF] ^ g_crc_table[1][(hi >> 16) & 0xFF] ^ g_crc_table[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ g_crc_table[0][(crc ^ *p++) & 0xFFu];
  return crc;
}

// CRC-32C of data[0..n) continuing from `seed` (0 for a fresh checksum), as crc32c::Extend in TensorFlow/LevelDB.
// impl: 0 = fastest available (SSE4.2 when the CPU has it), 1 = portable table version (used to cross-check).
int yb_crc32c(const void* data, size_t n, uint32_t seed, int impl, uint32_t* out) {
  if ((!data && n) || !out) return fail(YB_ERR_INVALID, "yb_crc32c: bad argument");
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = seed ^ 0xFFFFFFFFu;
#if defined(__x86_64__)
  if (impl == 0 && __builtin_cpu_supports("sse4.2")) { *out = crc32c_hw(c, p, n) ^ 0xFFFFFFFFu; return YB_OK; }
#endif
  *out = crc32c_sw(c, p, n) ^ 0xFFFFFFFFu;
  return YB_OK;
}

int yb_abi_version(void) { return YB_ABI_VERSION; }
const char* yb_last_error(void) { return g_err; }

int yb_device_count(int* count) {
  if (!count) return fail(YB_ERR_INVALID, "count is NULL");
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
  *count = c;
  return YB_OK;
}

int yb_engine_create(const yb_layer* plan, int n_layers, int in_h, int in_w, int in_c, int max_batch, int device,
                     int decode_mode, int num_classes, yb_engine** out) {
  if (!plan || !out || n_layers < 2 || in_h <= 0 || in_w <= 0 || in_c <= 0 || max_batch <= 0 || num_classes <= 0)
    return fail(YB_ERR_INVALID, "yb_engine_create: bad argument");
  YB_TRY(set_device(device));
  YB_TRY(load_driver_fns());
  yb_engine* e = new yb_engine();
  e->device = device; e->H = in_h; e->W = in_w; e->C = in_c; e->max_batch = max_batch;
  e->decode_mode = decode_mode; e->num_classes = num_classes;
  e->plan.assign(plan, plan + n_layers);
  const char* ka = getenv("YB_KEEP_ALL");
  e->keep_all = ka && atoi(ka) != 0;
  const char* fu = getenv("YB_FUSE_UPSAMPLE");
  if (fu) e->fuse_upsample = atoi(fu) != 0;
  const char* fs = getenv("YB_FUSE_STEM");
  if (fs) e->fuse_stem = atoi(fs);
  const char* fb = getenv("YB_FUSE_BLOCK");
  if (fb) e->fuse_block = atoi(fb);
  const char* sk = getenv("YB_SPLIT_K");
  if (sk && atoi(sk) != 0) e->split_k = -1;          // workspace allocated once the device is known (below)
  const char* fp = getenv("YB_FUSE_POOL");
  if (fp) e->fuse_pool = atoi(fp);
  const char* pp = getenv("YB_PIXEL_PAIRS");
  if (pp) e->pixel_pairs = atoi(pp);
  const char* pd = getenv("YB_PDL");
  if (pd) e->pdl = atoi(pd) != 0;
  const char* cp = getenv("YB_PAIR");
  if (cp) e->cta_pairs = atoi(cp) != 0;
  const char* te = getenv("YB_TMA_EPI");
  if (te) e->tma_epilogue = atoi(te) != 0;
  const char* bs = getenv("YB_BSTAT");
  if (bs) e->b_stationary = atoi(bs) != 0;
  cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, device);
  const char* bm = getenv("YB_BN_MAX");
  if (bm && (atoi(bm) == 32 || atoi(bm) == 64 || atoi(bm) == 128 || atoi(bm) == 256)) e->bn_max = atoi(bm);
  int r = compile_plan(e);
  if (r == YB_OK) {
    auto go = [&]() -> int {
      YB_TRY(set_kernel_attrs(e));
      if (e->split_k < 0) YB_TRY(set_split_k(e, true));
      YB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
      YB_CUDA(cudaMalloc(&e->arena, e->arena_bytes));
      YB_CUDA(cudaMemset(e->arena, 0, e->arena_bytes));
      YB_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        YB_CUDA(cudaMalloc(&e->input_dev[i], (size_t)max_batch * in_h * in_w * in_c * 4));
        YB_CUDA(cudaEventCreateWithFlags(&e->ev_copied[i], cudaEventDisableTiming));
        YB_CUDA(cudaEventCreateWithFlags(&e->ev_input_free[i], cudaEventDisableTiming));
        YB_CUDA(cudaEventCreateWithFlags(&e->ev_fetch[i], cudaEventDisableTiming));
      }
      for (int i = 0; i < 8; ++i) YB_CUDA(cudaEventCreate(&e->marks[i]));
      YB_CUDA(cudaEventCreateWithFlags(&e->ev_input_read, cudaEventDisableTiming));
      e->last_input_reader = 0;
      for (size_t i = 0; i < e->ops.size(); ++i)
        if (e->ops[i].in.buf == -2 || e->ops[i].in2.buf == -2) e->last_input_reader = (int)i;
      float lut[256];
      for (int i = 0; i < 256; ++i) lut[i] = (float)((double)i / 255.0);   // image / 255. (net/base.py:153) then the fp32 feed cast
      YB_CUDA(cudaMalloc(&e->d_u8_lut, sizeof(lut)));
      YB_CUDA(cudaMemcpy(e->d_u8_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
      YB_TRY(e->post.alloc(max_batch, e->rows, e->box_len, decode_mode == YB_DECODE_V2));
      e->post.n_scales = (int)e->heads.size();
      return YB_OK;
    };
    r = go();
  }
  if (r != YB_OK) { yb_engine_destroy(e); return r; }
  *out = e;
  return YB_OK;
}

void yb_engine_destroy(yb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  clear_graphs(e);
  for (Op& op : e->ops) { cudaFree(op.d_wt); cudaFree(op.d_wt32); cudaFree(op.d_scale); cudaFree(op.d_shift); }
  if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
  cudaFree(e->dbg_counters);
  cudaFree(e->splitk_ws); cudaFree(e->splitk_cnt);
  cudaFree(e->arena); cudaFree(e->input_dev[0]); cudaFree(e->input_dev[1]); cudaFree(e->d_u8_lut); cudaFree(e->scratch_f32);
  e->post.release();
  e->pre.release();
  for (int i = 0; i < 2; ++i) {
    if (e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]);
    if (e->ev_input_free[i]) cudaEventDestroy(e->ev_input_free[i]);
    if (e->ev_fetch[i]) cudaEventDestroy(e->ev_fetch[i]);
  }
  for (int i = 0; i < 8; ++i) if (e->marks[i]) cudaEventDestroy(e->marks[i]);
  if (e->ev_input_read) cudaEventDestroy(e->ev_input_read);
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int yb_engine_load_weights(yb_engine* e, const float* stream, size_t n, size_t* consumed) {
  if (!e || !stream) return fail(YB_ERR_INVALID, "yb_engine_load_weights: bad argument");
  YB_TRY(set_device(e->device));
  clear_graphs(e);
  size_t need = 0;
  for (const Op& op : e->ops)
    if (op.kind == OP_CONV) {
      const size_t co = op.px_pair == 1 ? op.cout / 2 : op.cout, ci = op.px_pair ? op.cin / 2 : op.cin;     // the file's dimensions
      need += (size_t)(e->plan[op.layer].batch_norm ? 4 : 1) * co + co * ci * op.ksize * op.ksize;
    }
  if (n < need) {
    if (consumed) *consumed = 0;
    return fail(YB_ERR_SHORT_WEIGHTS, "weight stream has %zu floats, the plan needs %zu", n, need);
  }
  size_t read = 0;
  std::vector<uint16_t> packed;
  std::vector<float> packed32, scale, shift;
  for (Op& op : e->ops) {
    if (op.kind != OP_CONV) continue;
    // co, ci: the file's dimensions; a pixel-pair op packs them into twice the channels on either side
    const int co = op.px_pair == 1 ? op.cout / 2 : op.cout, ci = op.px_pair ? op.cin / 2 : op.cin, k = op.ksize, K = k * k * ci;
    const bool bn = e->plan[op.layer].batch_norm != 0;
    scale.assign(op.cout_pad, 0.0f); shift.assign(op.cout_pad, 0.0f);
    if (bn) {   // stream order: beta, gamma, moving_mean, moving_variance (net/layers.py:53-63)
      const float *beta = stream + read, *gamma = beta + co, *mean = gamma + co, *var = mean + co;
      for (int o = 0; o < co; ++o) {
        const double inv = (double)gamma[o] / std::sqrt((double)var[o] + 1e-5);
        scale[o] = (float)inv;
        shift[o] = (float)((double)beta[o] - (double)mean[o] * inv);
      }
      read += 4 * (size_t)co;
    } else {
      for (int o = 0; o < co; ++o) { scale[o] = 1.0f; shift[o] = stream[read + o]; }
      read += co;
    }
    if (op.px_pair == 1) for (int o = 0; o < co; ++o) { scale[co + o] = scale[o]; shift[co + o] = shift[o]; }   // both pixels of a pair
    const float* kern = stream + read;   // [O][I][kh][kw] (net/base.py:36-40)
    read += (size_t)co * K;
    cudaFree(op.d_wt); cudaFree(op.d_wt32); cudaFree(op.d_scale); cudaFree(op.d_shift);
    op.d_wt = nullptr; op.d_wt32 = nullptr; op.d_scale = nullptr; op.d_shift = nullptr;
    if (op.path == PATH_DIRECT) {
      packed32.assign((size_t)K * co, 0.0f);
      for (int o = 0; o < co; ++o)
        for (int i = 0; i < ci; ++i)
          for (int t = 0; t < k * k; ++t) packed32[(size_t)(t * ci + i) * co + o] = kern[((size_t)o * ci + i) * k * k + t];
      YB_CUDA(cudaMalloc(&op.d_wt32, packed32.size() * 4));
      YB_CUDA(cudaMemcpy(op.d_wt32, packed32.data(), packed32.size() * 4, cudaMemcpyHostToDevice));
    } else if (op.px_pair == 2) {
      // Input pairs only (stride 2): tap (ky, kp), kp = 0 / 1 the pair left of / at the output pixel, input channel
      // ei * ci + i; output pixel x reads input pixel 2 (x + kp - 1) + ei = 2x + kx - 1, i.e. kx = 2 kp + ei - 1.
      const int K2 = k * 2 * 2 * ci;
      packed.assign((size_t)op.cout_pad * K2, 0);
      for (int o = 0; o < co; ++o)
        for (int ky = 0; ky < k; ++ky)
          for (int kp = 0; kp < 2; ++kp)
            for (int ei = 0; ei < 2; ++ei) {
              const int kx = 2 * kp + ei - 1;
              if (kx < 0 || kx >= k) continue;
              for (int i = 0; i < ci; ++i)
                packed[(size_t)o * K2 + (size_t)(ky * 2 + kp) * 2 * ci + ei * ci + i] =
                    f32_to_bf16_rne(kern[((size_t)o * ci + i) * k * k + ky * k + kx]);
            }
      YB_CUDA(cudaMalloc(&op.d_wt, packed.size() * 2));
      YB_CUDA(cudaMemcpy(op.d_wt, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    } else if (op.px_pair == 1) {
      // Paired problem: output channel eo * co + o (eo = pixel of the output pair), input channel ei * ci + i, tap
      // (ky, kp) with kp the offset in pairs.  Output pixel 2q + eo reads input pixel 2(q + kp - pad) + ei, i.e. the
      // original tap kx = 2 (kp - pad) + ei - eo + pad; the combinations that fall outside the kernel stay zero.
      const int K2 = k * k * 2 * ci, pad = (k - 1) / 2;
      packed.assign((size_t)op.cout_pad * K2, 0);
      for (int eo = 0; eo < 2; ++eo)
        for (int o = 0; o < co; ++o)
          for (int ky = 0; ky < k; ++ky)
            for (int kp = 0; kp < k; ++kp)
              for (int ei = 0; ei < 2; ++ei) {
                const int kx = 2 * (kp - pad) + ei - eo + pad;
                if (kx < 0 || kx >= k) continue;
                for (int i = 0; i < ci; ++i)
                  packed[(size_t)(eo * co + o) * K2 + (size_t)(ky * k + kp) * 2 * ci + ei * ci + i] =
                      f32_to_bf16_rne(kern[((size_t)o * ci + i) * k * k + ky * k + kx]);
              }
      YB_CUDA(cudaMalloc(&op.d_wt, packed.size() * 2));
      YB_CUDA(cudaMemcpy(op.d_wt, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    } else {
      packed.assign((size_t)op.cout_pad * K, 0);
      for (int o = 0; o < co; ++o)
        for (int i = 0; i < ci; ++i)
          for (int t = 0; t < k * k; ++t) packed[(size_t)o * K + t * ci + i] = f32_to_bf16_rne(kern[((size_t)o * ci + i) * k * k + t]);
      YB_CUDA(cudaMalloc(&op.d_wt, packed.size() * 2));
      YB_CUDA(cudaMemcpy(op.d_wt, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    }
    YB_CUDA(cudaMalloc(&op.d_scale, (size_t)op.cout_pad * 4));
    YB_CUDA(cudaMalloc(&op.d_shift, (size_t)op.cout_pad * 4));
    YB_CUDA(cudaMemcpy(op.d_scale, scale.data(), (size_t)op.cout_pad * 4, cudaMemcpyHostToDevice));
    YB_CUDA(cudaMemcpy(op.d_shift, shift.data(), (size_t)op.cout_pad * 4, cudaMemcpyHostToDevice));
    op.h_scale = scale; op.h_shift = shift;
  }
  if (consumed) *consumed = read;
  YB_TRY(build_tensor_maps(e));
  e->weights_loaded = true;
  return YB_OK;
}

// enqueues every op of the plan on the engine's stream; slot >= 0: the input sits in staging buffer `slot`
// Replays (or captures) the CUDA graph of the forward for this batch size and input buffer.  Returns 1 when the forward
// was launched through a graph, 0 when the caller has to enqueue the kernels itself, negative on error.
static int try_graph_forward(yb_engine* e, int n, int slot) {
  const bool want = e->graph_mode == 1 || (e->graph_mode < 0 && n <= YB_GRAPH_AUTO_MAX_N);
  if (!want || e->prof_on || e->ablate || e->dbg_counters || e->conv_impl != 0) return 0;
  yb_engine::FwdGraph* fg = nullptr;
  for (auto& g : e->graphs) if (g.n == n && g.input == e->cur_input && g.dtype == e->cur_input_dtype) { fg = &g; break; }
  if (!fg) {
    if (e->graphs.size() >= 32) clear_graphs(e);
    e->graphs.push_back({n, e->cur_input, e->cur_input_dtype, 0, nullptr});
    fg = &e->graphs.back();
  }
  if (!fg->exec) {
    if (fg->seen < 0) return 0;                    // capture failed before: stay eager
    if (fg->seen++ == 0) return 0;                 // first time: eager
    if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); fg->seen = -1; return 0; }
    int rc = YB_OK;
    for (Op& op : e->ops) { rc = run_op(e, op, n); if (rc != YB_OK) break; }
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    if (rc != YB_OK || ce != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      fg->seen = -1;
      return rc != YB_OK ? rc : 0;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ci != cudaSuccess || !exec) { cudaGetLastError(); fg->seen = -1; return 0; }
    fg->exec = exec;
  }
  YB_CUDA(cudaGraphLaunch(fg->exec, e->stream));
  // the staging buffer is free again once the whole graph has run (coarser than the eager path's event after the last
  // reader of the input; irrelevant at the batch sizes graphs are used for)
  if (slot >= 0) YB_CUDA(cudaEventRecord(e->ev_input_free[slot], e->stream));
  YB_CUDA(cudaEventRecord(e->ev_input_read, e->stream));
  ++e->graph_replays;
  e->fwd_launches = 0;
  for (const Op& op : e->ops) if (!op.skip) ++e->fwd_launches;
  e->last_n = n;
  e->detected = false;
  return 1;
}

static int enqueue_ops(yb_engine* e, int n, int slot, cudaEvent_t* evs, int* n_ev) {
  if (!evs) {
    const int g = try_graph_forward(e, n, slot);
    if (g != 0) return g < 0 ? g : YB_OK;
  }
  e->fwd_launches = 0;
  std::vector<cudaEvent_t> own;
  if (!evs && e->prof_on) {
    own.resize(e->ops.size() + 1);
    for (auto& ev : own) YB_CUDA(cudaEventCreate(&ev));
    e->prof_events.insert(e->prof_events.end(), own.begin(), own.end());
    evs = own.data();
  }
  int k = 0;
  for (Op& op : e->ops) {
    if (evs) YB_CUDA(cudaEventRecord(evs[k], e->stream));
    YB_TRY(run_op(e, op, n));
    if (k == e->last_input_reader) {
      if (slot >= 0) YB_CUDA(cudaEventRecord(e->ev_input_free[slot], e->stream));
      YB_CUDA(cudaEventRecord(e->ev_input_read, e->stream));
    }
    if (!op.skip) ++e->fwd_launches;
    ++k;
  }
  if (evs) { YB_CUDA(cudaEventRecord(evs[k], e->stream)); if (n_ev) *n_ev = k + 1; }
  e->last_n = n;
  e->detected = false;
  return YB_OK;
}

static int forward_impl(yb_engine* e, const void* images, int dtype, int mem, int n, cudaEvent_t* evs, int* n_ev) {
  if (!e || !images) return fail(YB_ERR_INVALID, "yb_engine_forward: bad argument");
  if (n <= 0 || n > e->max_batch) return fail(YB_ERR_INVALID, "batch %d outside [1,%d]", n, e->max_batch);
  if (dtype != YB_F32 && dtype != YB_U8) return fail(YB_ERR_INVALID, "unknown image dtype %d", dtype);
  if (!e->weights_loaded) return fail(YB_ERR_STATE, "yb_engine_forward before yb_engine_load_weights");
  YB_TRY(set_device(e->device));
  const size_t bytes = (size_t)n * e->H * e->W * e->C * (dtype == YB_F32 ? 4 : 1);
  int slot = -1;
  if (mem == YB_MEM_HOST) {
    // staging[slot] may still be read by the first conv of the forward before last: wait for it, copy on the
    // copy stream (overlaps whatever the compute stream is doing), then make the compute stream wait for the copy
    slot = e->stage_toggle;
    e->stage_toggle ^= 1;
    YB_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_input_free[slot], 0));
    YB_CUDA(cudaMemcpyAsync(e->input_dev[slot], images, bytes, cudaMemcpyHostToDevice, e->copy_stream));
    YB_CUDA(cudaEventRecord(e->ev_copied[slot], e->copy_stream));
    YB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_copied[slot], 0));
    e->cur_input = e->input_dev[slot];
  } else if ((reinterpret_cast<uintptr_t>(images) & 15u) != 0) {
    // a device pointer the TMA-fed first layers cannot address (conv_fused.cuh needs 16-byte alignment): one
    // device-to-device copy into the staging buffer, on the compute stream
    slot = e->stage_toggle;
    e->stage_toggle ^= 1;
    YB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_input_free[slot], 0));
    YB_CUDA(cudaMemcpyAsync(e->input_dev[slot], images, bytes, cudaMemcpyDeviceToDevice, e->stream));
    e->cur_input = e->input_dev[slot];
  } else {
    e->cur_input = images;
  }
  e->cur_input_dtype = dtype;
  e->last_input_is_u8_staged = false;
  return enqueue_ops(e, n, slot, evs, n_ev);
}

int yb_engine_forward(yb_engine* e, const void* images, int dtype, int mem, int n) {
  return forward_impl(e, images, dtype, mem, n, nullptr, nullptr);
}

int yb_engine_forward_raw(yb_engine* e, const void* const* images, const int* heights, const int* widths, const int* strides, int n) {
  if (!e || !images || !heights || !widths) return fail(YB_ERR_INVALID, "yb_engine_forward_raw: bad argument");
  if (n <= 0 || n > e->max_batch) return fail(YB_ERR_INVALID, "batch %d outside [1,%d]", n, e->max_batch);
  if (!e->weights_loaded) return fail(YB_ERR_STATE, "yb_engine_forward_raw before yb_engine_load_weights");
  if (e->C != 3) return fail(YB_ERR_INVALID, "raw BGR input needs a 3-channel network input, this one has %d", e->C);
  // The reference passes (input_h, input_w) to cv2.resize as dsize = (width, height) (net/base.py:121): for a non-square
  // network the resized array has shape (input_w, input_h, 3) and its [None, input_h, input_w, 3] placeholder rejects it.
  if (e->H != e->W)
    return fail(YB_ERR_INVALID, "non-square network input %dx%d: the reference's preprocess_image resizes to (%d, %d), which its placeholder rejects",
                e->H, e->W, e->W, e->H);
  YB_TRY(set_device(e->device));
  const int slot = e->stage_toggle;
  e->stage_toggle ^= 1;
  YB_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_input_free[slot], 0));
  YB_TRY(e->pre.run(slot, e->copy_stream, e->stream, e->ev_copied[slot], images, heights, widths, strides, n, e->H, e->W,
                    reinterpret_cast<unsigned char*>(e->input_dev[slot])));
  e->cur_input = e->input_dev[slot];
  e->cur_input_dtype = YB_U8;
  e->last_input_is_u8_staged = true;
  const int r = enqueue_ops(e, n, slot, nullptr, nullptr);
  if (r == YB_OK) ++e->fwd_launches;      // the resize kernel
  return r;
}

int yb_engine_read_input_u8(yb_engine* e, unsigned char* host_out, size_t capacity) {
  if (!e || !host_out) return fail(YB_ERR_INVALID, "yb_engine_read_input_u8: bad argument");
  if (e->last_n <= 0 || !e->last_input_is_u8_staged) return fail(YB_ERR_STATE, "yb_engine_read_input_u8 needs a preceding yb_engine_forward_raw");
  const size_t total = (size_t)e->last_n * e->H * e->W * e->C;
  if (capacity < total) return fail(YB_ERR_CAPACITY, "input needs %zu bytes, buffer holds %zu", total, capacity);
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaMemcpyAsync(host_out, e->cur_input, total, cudaMemcpyDeviceToHost, e->stream));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  return YB_OK;
}

// Stand-alone: cv2.resize(image, (dst_w, dst_h)) [INTER_LINEAR] + BGR->RGB for n host images into host memory
// [n, dst_h, dst_w, 3] (kernel-level parity tests, any destination shape).
int yb_resize_bgr2rgb(const void* const* images, const int* heights, const int* widths, const int* strides, int n, int dst_h, int dst_w,
                      unsigned char* dst_host, int device) {
  if (!images || !heights || !widths || !dst_host || n <= 0 || dst_h <= 0 || dst_w <= 0) return fail(YB_ERR_INVALID, "yb_resize_bgr2rgb: bad argument");
  YB_TRY(set_device(device));
  PreCtx pre;
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr;
  unsigned char* d_out = nullptr;
  const size_t total = (size_t)n * dst_h * dst_w * 3;
  auto go = [&]() -> int {
    YB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    YB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    YB_CUDA(cudaMalloc(&d_out, total));
    YB_TRY(pre.run(0, st, st, ev, images, heights, widths, strides, n, dst_h, dst_w, d_out));
    YB_CUDA(cudaMemcpyAsync(dst_host, d_out, total, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    return YB_OK;
  };
  const int r = go();
  pre.release();
  cudaFree(d_out);
  if (ev) cudaEventDestroy(ev);
  if (st) cudaStreamDestroy(st);
  return r;
}

int yb_engine_output_shape(yb_engine* e, int* rows, int* cols) {
  if (!e || !rows || !cols) return fail(YB_ERR_INVALID, "yb_engine_output_shape: bad argument");
  if (e->decode_mode == YB_DECODE_V2) { *rows = e->heads[0].h * e->heads[0].w; *cols = e->heads[0].na * e->box_len; }
  else { *rows = e->rows; *cols = e->box_len; }
  return YB_OK;
}

int yb_engine_read_output(yb_engine* e, float* host_out, size_t capacity) {
  if (!e || !host_out) return fail(YB_ERR_INVALID, "yb_engine_read_output: bad argument");
  if (e->last_n <= 0) return fail(YB_ERR_STATE, "yb_engine_read_output before yb_engine_forward");
  YB_TRY(set_device(e->device));
  const size_t total = (size_t)e->last_n * e->rows * e->box_len;
  if (capacity < total) return fail(YB_ERR_CAPACITY, "output needs %zu floats, buffer holds %zu", total, capacity);
  YB_TRY(ensure_scratch(e, total));
  // [n, R, 5+C]; for v2 this is bit-for-bit the reference's [n, h, w, A*(5+C)] (rows = (cy*w+cx)*A + a)
  for (const Head& hd : e->heads) {
    const long long cnt = (long long)e->last_n * hd.h * hd.w * hd.na * e->box_len;
    gather_head_kernel<<<ceil_div(cnt, 256), 256, 0, e->stream>>>(reinterpret_cast<const float*>(view_ptr(e, hd.view)), hd.view.ld,
                                                                 hd.h * hd.w, hd.na, e->box_len, hd.row_begin, e->rows, e->scratch_f32, cnt);
    YB_CUDA(cudaGetLastError());
  }
  YB_CUDA(cudaMemcpyAsync(host_out, e->scratch_f32, total * 4, cudaMemcpyDeviceToHost, e->stream));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  return YB_OK;
}

int yb_engine_read_layer(yb_engine* e, int layer, float* host_out, size_t capacity, int shape_hwc[3]) {
  if (!e || !host_out || !shape_hwc) return fail(YB_ERR_INVALID, "yb_engine_read_layer: bad argument");
  if (layer < 0 || layer >= (int)e->plan.size()) return fail(YB_ERR_INVALID, "layer %d out of range", layer);
  if (e->last_n <= 0) return fail(YB_ERR_STATE, "yb_engine_read_layer before yb_engine_forward");
  YB_TRY(set_device(e->device));
  const View& v = e->view[layer];
  if (v.buf == -1) return fail(YB_ERR_INVALID, "layer %d is fused into its consumer and has no tensor of its own", layer);
  if (v.buf >= 0 && !e->keep_all && !e->bufs[v.buf].keep)
    return fail(YB_ERR_STATE, "layer %d's buffer is recycled by the arena; create the engine with YB_KEEP_ALL=1 to read intermediates", layer);
  const size_t total = (size_t)e->last_n * v.h * v.w * v.c;
  if (capacity < total) return fail(YB_ERR_CAPACITY, "layer needs %zu floats, buffer holds %zu", total, capacity);
  shape_hwc[0] = v.h; shape_hwc[1] = v.w; shape_hwc[2] = v.c;
  YB_TRY(ensure_scratch(e, total));
  if (v.buf == -2 && e->cur_input_dtype == YB_U8) return fail(YB_ERR_INVALID, "reading back a uint8 input is not supported");
  read_view_kernel<<<ceil_div((long long)total, 256), 256, 0, e->stream>>>(view_ptr(e, v), v.ld, v.c, v.f32 ? 1 : 0, e->scratch_f32, (long long)total);
  YB_CUDA(cudaGetLastError());
  YB_CUDA(cudaMemcpyAsync(host_out, e->scratch_f32, total * 4, cudaMemcpyDeviceToHost, e->stream));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  return YB_OK;
}

int yb_engine_detect_async(yb_engine* e, double threshold, double iou_threshold, int nms_mode) {
  if (!e) return fail(YB_ERR_INVALID, "yb_engine_detect: engine is NULL");
  if (e->last_n <= 0) return fail(YB_ERR_STATE, "yb_engine_detect before yb_engine_forward");
  if (nms_mode != YB_NMS_REFERENCE && nms_mode != YB_NMS_PER_CLASS) return fail(YB_ERR_INVALID, "unknown nms mode %d", nms_mode);
  YB_TRY(set_device(e->device));
  ScaleDesc sc[POST_MAX_SCALES];
  engine_scales(e, sc);
  // score: numpy >= 2 compares the float32 score with the Python float cast to float32 (NEP 50), net/v3.py:121
  YB_TRY(e->post.decode(e->stream, sc, e->last_n, (float)threshold));
  // IoU: float64 against the Python float itself (net/base.py:203)
  YB_TRY(e->post.nms(e->stream, e->last_n, iou_threshold, nms_mode));
  e->det_launches = 2;
  e->detected = true;
  return YB_OK;
}

static int copy_dets(PostCtx& post, cudaStream_t st, int n, yb_det* out, int* counts, int max_per_image, int* det_launches) {
  if (max_per_image <= 0) return fail(YB_ERR_INVALID, "max_per_image must be positive");
  YB_TRY(post.gather(st, n, max_per_image));
  if (det_launches) ++*det_launches;
  static_assert(sizeof(yb_det) == sizeof(DetOut), "yb_det layout");
  YB_CUDA(cudaMemcpyAsync(counts, post.n_keep, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaMemcpyAsync(out, post.dets, (size_t)n * max_per_image * sizeof(DetOut), cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaStreamSynchronize(st));
  return YB_OK;
}

int yb_engine_detect(yb_engine* e, double threshold, double iou_threshold, int nms_mode, yb_det* out, int* counts, int max_per_image) {
  if (!out || !counts) return fail(YB_ERR_INVALID, "yb_engine_detect: output buffers are NULL");
  YB_TRY(yb_engine_detect_async(e, threshold, iou_threshold, nms_mode));
  return copy_dets(e->post, e->stream, e->last_n, out, counts, max_per_image, &e->det_launches);
}

// ---- stream ordering against the caller's streams (include/yolo_b200.h, "STREAM ORDER") ----
static int order_streams(cudaStream_t first, cudaStream_t then) {
  cudaEvent_t ev = nullptr;
  YB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t r = cudaEventRecord(ev, first);
  if (r == cudaSuccess) r = cudaStreamWaitEvent(then, ev, 0);
  cudaEventDestroy(ev);          // released once the recorded work has completed
  if (r != cudaSuccess) { cudaGetLastError(); return fail(YB_ERR_CUDA, "stream ordering failed: %s", cudaGetErrorString(r)); }
  return YB_OK;
}
int yb_engine_order_after(yb_engine* e, void* cuda_stream) {
  if (!e) return fail(YB_ERR_INVALID, "engine is NULL");
  YB_TRY(set_device(e->device));
  return order_streams(static_cast<cudaStream_t>(cuda_stream), e->stream);
}
int yb_engine_order_before(yb_engine* e, void* cuda_stream) {
  if (!e) return fail(YB_ERR_INVALID, "engine is NULL");
  YB_TRY(set_device(e->device));
  if (!e->ev_input_read) return YB_OK;                   // no forward yet
  YB_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(cuda_stream), e->ev_input_read, 0));
  return YB_OK;
}

int yb_engine_sync(yb_engine* e) {
  if (!e) return fail(YB_ERR_INVALID, "engine is NULL");
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  return YB_OK;
}

int yb_engine_profile(yb_engine* e, const void* images, int dtype, int mem, int n, int* layer_idx, float* ms, int cap, int* n_ops) {
  if (!e || !layer_idx || !ms || !n_ops) return fail(YB_ERR_INVALID, "yb_engine_profile: bad argument");
  YB_TRY(set_device(e->device));
  const int nops = (int)e->ops.size();
  std::vector<cudaEvent_t> evs(nops + 1);
  for (auto& ev : evs) YB_CUDA(cudaEventCreate(&ev));
  int n_ev = 0;
  int r = forward_impl(e, images, dtype, mem, n, evs.data(), &n_ev);
  if (r == YB_OK) {
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) r = fail(YB_ERR_CUDA, "profile sync failed: %s", cudaGetErrorString(ce));
  }
  if (r == YB_OK) {
    *n_ops = nops;
    for (int i = 0; i < nops && i < cap; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, evs[i], evs[i + 1]);
      ms[i] = t;
      layer_idx[i] = e->ops[i].layer;
    }
  }
  for (auto& ev : evs) cudaEventDestroy(ev);
  return r;
}

int yb_engine_set_conv_impl(yb_engine* e, int impl) {
  if (!e || (impl != 0 && impl != 1)) return fail(YB_ERR_INVALID, "yb_engine_set_conv_impl: bad argument");
  e->conv_impl = impl;
  clear_graphs(e);
  return YB_OK;
}

int yb_engine_autotune(yb_engine* e, int n, int reps) {
  if (!e) return fail(YB_ERR_INVALID, "engine is NULL");
  if (n <= 0 || n > e->max_batch) return fail(YB_ERR_INVALID, "batch %d outside [1,%d]", n, e->max_batch);
  if (!e->weights_loaded) return fail(YB_ERR_STATE, "yb_engine_autotune before yb_engine_load_weights");
  YB_TRY(set_device(e->device));
  clear_graphs(e);
  const int split_was = e->split_k;
  e->split_k = 0;                            // candidates are compared as plain kernels; split-K then applies on top
  const int rc = autotune(e, n, reps > 0 ? reps : 5);
  e->split_k = split_was;
  clear_graphs(e);                           // the launch configurations changed
  return rc;
}

int yb_engine_tune_report(yb_engine* e, char* buf, size_t capacity, size_t* needed) {
  if (!e) return fail(YB_ERR_INVALID, "engine is NULL");
  const size_t need = e->tune_report.size() + 1;
  if (needed) *needed = need;
  if (!buf) return YB_OK;
  if (capacity < need) return fail(YB_ERR_CAPACITY, "tune report needs %zu bytes, buffer holds %zu", need, capacity);
  memcpy(buf, e->tune_report.c_str(), need);
  return YB_OK;
}

int yb_engine_set_option(yb_engine* e, const char* name, int value) {
  if (!e || !name) return fail(YB_ERR_INVALID, "yb_engine_set_option: bad argument");
  clear_graphs(e);                           // every option changes what a captured forward would launch
  if (!strcmp(name, "graph")) e->graph_mode = value < 0 ? -1 : (value ? 1 : 0);
  else if (!strcmp(name, "pdl")) e->pdl = value != 0;
  else if (!strcmp(name, "ablate")) e->ablate = value & 15;
  else if (!strcmp(name, "solo_issue")) e->solo_issue = value != 0;
  else if (!strcmp(name, "cycles")) {       // in-kernel cycle counters of the conv roles (read with yb_engine_read_cycles)
    YB_TRY(set_device(e->device));
    if (value && !e->dbg_counters) {
      YB_CUDA(cudaMalloc(&e->dbg_counters, 16 * sizeof(unsigned long long)));
      YB_CUDA(cudaMemset(e->dbg_counters, 0, 16 * sizeof(unsigned long long)));
    } else if (!value && e->dbg_counters) {
      YB_CUDA(cudaStreamSynchronize(e->stream));
      cudaFree(e->dbg_counters);
      e->dbg_counters = nullptr;
    }
  }
  else if (!strcmp(name, "split_k")) {
    YB_TRY(set_device(e->device));
    YB_TRY(set_split_k(e, value != 0));
  }
  else if (!strcmp(name, "bstat")) e->b_stationary = value != 0;
  else if (!strcmp(name, "tma_epi")) e->tma_epilogue = value != 0;
  else if (!strcmp(name, "pairs")) {
    e->cta_pairs = value != 0;
    for (Op& op : e->ops) if (op.kind == OP_CONV && op.path == PATH_TC) op.cfg = default_cfg(op, e->max_batch, e->cta_pairs);
  } else return fail(YB_ERR_INVALID, "unknown option '%s'", name);
  return YB_OK;
}

int yb_engine_set_conv_cfg(yb_engine* e, int op_index, int bn, int pair, int bstat, int tma_epi, int ksub) {
  if (!e || op_index < 0 || op_index >= (int)e->ops.size()) return fail(YB_ERR_INVALID, "yb_engine_set_conv_cfg: bad argument");
  Op& op = e->ops[op_index];
  if (op.kind != OP_CONV || op.path != PATH_TC) return fail(YB_ERR_INVALID, "op %d is not a tcgen05 conv", op_index);
  clear_graphs(e);
  if (bn == 0) { op.cfg = default_cfg(op, e->max_batch, e->cta_pairs); return YB_OK; }
  if ((bn != 32 && bn != 64 && bn != 128 && bn != 256) || bn > op.bn_max) return fail(YB_ERR_INVALID, "op %d: N tile %d not available (max %d)", op_index, bn, op.bn_max);
  if (pair && bn < 64) return fail(YB_ERR_INVALID, "op %d: CTA pairs need an N tile of at least 64", op_index);
  if (ksub < 0 || ksub > 64) return fail(YB_ERR_INVALID, "op %d: bad blocks-per-stage %d", op_index, ksub);
  op.cfg.bn = bn; op.cfg.pair = pair ? 1 : 0; op.cfg.bstat = bstat; op.cfg.tma_epi = tma_epi; op.cfg.ksub = ksub;
  return YB_OK;
}

int yb_engine_read_cycles(yb_engine* e, unsigned long long* out16, int reset) {
  if (!e || !out16) return fail(YB_ERR_INVALID, "yb_engine_read_cycles: bad argument");
  if (!e->dbg_counters) return fail(YB_ERR_STATE, "cycle counters are off: yb_engine_set_option(e, \"cycles\", 1) first");
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  YB_CUDA(cudaMemcpy(out16, e->dbg_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) YB_CUDA(cudaMemset(e->dbg_counters, 0, 16 * sizeof(unsigned long long)));
  return YB_OK;
}

int yb_engine_time_op(yb_engine* e, int op_index, int n, int reps, float* ms) {
  if (!e || !ms || op_index < 0 || op_index >= (int)e->ops.size() || reps <= 0) return fail(YB_ERR_INVALID, "yb_engine_time_op: bad argument");
  if (n <= 0 || n > e->max_batch) return fail(YB_ERR_INVALID, "batch %d outside [1,%d]", n, e->max_batch);
  if (!e->weights_loaded) return fail(YB_ERR_STATE, "yb_engine_time_op before yb_engine_load_weights");
  YB_TRY(set_device(e->device));
  if (e->cur_input == nullptr) YB_TRY(prime_activations(e, n));
  const bool was_prof = e->prof_on;
  e->prof_on = false;
  const int r = time_op(e, e->ops[op_index], n, nullptr, reps, ms);
  e->prof_on = was_prof;
  e->last_n = 0; e->detected = false;
  return r;
}

int yb_engine_launch_count(yb_engine* e, int* forward_launches, int* detect_launches) {
  if (!e) return fail(YB_ERR_INVALID, "engine is NULL");
  if (forward_launches) *forward_launches = e->fwd_launches;
  if (detect_launches) *detect_launches = e->det_launches;
  return YB_OK;
}

int yb_engine_graph_replays(yb_engine* e, int* replays) {
  if (!e || !replays) return fail(YB_ERR_INVALID, "yb_engine_graph_replays: bad argument");
  *replays = e->graph_replays;
  return YB_OK;
}

int yb_engine_mark(yb_engine* e, int idx) {
  if (!e || idx < 0 || idx >= 8) return fail(YB_ERR_INVALID, "yb_engine_mark: bad argument");
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaEventRecord(e->marks[idx], e->stream));
  return YB_OK;
}

int yb_engine_elapsed(yb_engine* e, int from, int to, float* ms) {
  if (!e || !ms || from < 0 || from >= 8 || to < 0 || to >= 8) return fail(YB_ERR_INVALID, "yb_engine_elapsed: bad argument");
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaEventSynchronize(e->marks[to]));
  YB_CUDA(cudaEventElapsedTime(ms, e->marks[from], e->marks[to]));
  return YB_OK;
}

int yb_engine_fetch_async(yb_engine* e, yb_det* out, int* counts, int max_per_image, int slot) {
  if (!e || !out || !counts || slot < 0 || slot > 1) return fail(YB_ERR_INVALID, "yb_engine_fetch_async: bad argument");
  if (!e->detected) return fail(YB_ERR_STATE, "yb_engine_fetch_async before yb_engine_detect_async");
  YB_TRY(set_device(e->device));
  if (max_per_image <= 0) return fail(YB_ERR_INVALID, "max_per_image must be positive");
  YB_TRY(e->post.gather(e->stream, e->last_n, max_per_image));
  ++e->det_launches;
  YB_CUDA(cudaMemcpyAsync(counts, e->post.n_keep, (size_t)e->last_n * 4, cudaMemcpyDeviceToHost, e->stream));
  YB_CUDA(cudaMemcpyAsync(out, e->post.dets, (size_t)e->last_n * max_per_image * sizeof(DetOut), cudaMemcpyDeviceToHost, e->stream));
  YB_CUDA(cudaEventRecord(e->ev_fetch[slot], e->stream));
  return YB_OK;
}

int yb_engine_fetch_wait(yb_engine* e, int slot) {
  if (!e || slot < 0 || slot > 1) return fail(YB_ERR_INVALID, "yb_engine_fetch_wait: bad argument");
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaEventSynchronize(e->ev_fetch[slot]));
  return YB_OK;
}

int yb_engine_profiling(yb_engine* e, int enable) {
  if (!e) return fail(YB_ERR_INVALID, "engine is NULL");
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  e->prof_events.clear();
  e->prof_on = enable != 0;
  return YB_OK;
}

int yb_engine_profile_read(yb_engine* e, int* layer_idx, float* ms_sum, int cap, int* n_ops, int* n_forwards) {
  if (!e || !layer_idx || !ms_sum || !n_ops || !n_forwards) return fail(YB_ERR_INVALID, "yb_engine_profile_read: bad argument");
  YB_TRY(set_device(e->device));
  YB_CUDA(cudaStreamSynchronize(e->stream));
  const int nops = (int)e->ops.size();
  const int per = nops + 1;
  const int nf = (int)e->prof_events.size() / per;
  *n_ops = nops; *n_forwards = nf;
  for (int i = 0; i < nops && i < cap; ++i) { ms_sum[i] = 0.f; layer_idx[i] = e->ops[i].layer; }
  for (int f = 0; f < nf; ++f)
    for (int i = 0; i < nops && i < cap; ++i) {
      float t = 0.f;
      YB_CUDA(cudaEventElapsedTime(&t, e->prof_events[f * per + i], e->prof_events[f * per + i + 1]));
      ms_sum[i] += t;
    }
  return YB_OK;
}

int yb_engine_op_info(yb_engine* e, int op_index, int* layer, int* path, int* bn_tile, int* bk, int* stages, double* flops_per_image) {
  if (!e || op_index < 0 || op_index >= (int)e->ops.size()) return fail(YB_ERR_INVALID, "yb_engine_op_info: bad argument");
  const Op& op = e->ops[op_index];
  if (layer) *layer = op.layer;
  if (path) *path = op.kind == OP_CONV ? op.path : -1 - op.kind;
  if (bn_tile) *bn_tile = op.cfg.bn;
  if (bk) *bk = op.bk;
  if (stages) *stages = op.launched_stages;
  // a pixel-pair op multiplies twice the useful products (half of its packed weights are structural zeros)
  // (px_pair == 2: the 3 x 2 kernel over 64 channels holds the 3 x 3 x 32 useful products)
  auto conv_flops = [](const Op& o) { return 2.0 * o.Ho * o.Wo * o.cout * o.ksize * o.ksize * o.cin / (o.px_pair ? 2.0 : 1.0); };
  if (flops_per_image) {
    *flops_per_image = (op.kind == OP_CONV && !op.skip) ? conv_flops(op) : 0.0;
    if (op.kind == OP_CONV && op.fuse_kind && op.fuse_src >= 0) *flops_per_image += conv_flops(e->ops[op.fuse_src]);     // both layers of a fused pair
  }
  return YB_OK;
}

int yb_engine_op_cfg(yb_engine* e, int op_index, int* bn, int* pair, int* bstat, int* tma_epi, int* stages, int* ksub) {
  if (!e || op_index < 0 || op_index >= (int)e->ops.size()) return fail(YB_ERR_INVALID, "yb_engine_op_cfg: bad argument");
  const Op& op = e->ops[op_index];
  if (bn) *bn = op.cfg.bn;
  if (pair) *pair = op.cfg.pair;
  if (bstat) *bstat = op.launched_bstat;
  if (tma_epi) *tma_epi = op.launched_tma_epi;
  if (stages) *stages = op.launched_stages;
  if (ksub) *ksub = op.launched_ksub;
  return YB_OK;
}

int yb_engine_op_splitk(yb_engine* e, int op_index, int* splitk) {
  if (!e || !splitk || op_index < 0 || op_index >= (int)e->ops.size()) return fail(YB_ERR_INVALID, "yb_engine_op_splitk: bad argument");
  *splitk = e->ops[op_index].launched_splitk;
  return YB_OK;
}

// ---- stand-alone post-processing ----
struct yb_post {
  int device = 0;
  cudaStream_t stream = nullptr;
  PostCtx ctx;
  float* staging = nullptr;   // device copy of host head tensors
  size_t staging_floats = 0;
  int num_classes = 0, decode_mode = 0;
  float decode_ms = 0.f, nms_ms = 0.f;
  bool timed = false;
};

int yb_post_create(const yb_scale* scales, int n_scales, int num_classes, int decode_mode, int max_batch, int device, yb_post** out) {
  if (!scales || !out || n_scales < 1 || n_scales > POST_MAX_SCALES || num_classes <= 0 || max_batch <= 0)
    return fail(YB_ERR_INVALID, "yb_post_create: bad argument");
  YB_TRY(set_device(device));
  yb_post* p = new yb_post();
  p->device = device; p->num_classes = num_classes; p->decode_mode = decode_mode;
  int rows = 0;
  for (int i = 0; i < n_scales; ++i) {
    if (scales[i].n_anchors < 1 || scales[i].n_anchors > YB_MAX_ANCHORS || scales[i].h <= 0 || scales[i].w <= 0) {
      delete p;
      return fail(YB_ERR_INVALID, "yb_post_create: bad scale %d", i);
    }
    ScaleDesc& s = p->ctx.scales[i];
    memset(&s, 0, sizeof(s));
    s.h = scales[i].h; s.w = scales[i].w; s.na = scales[i].n_anchors; s.row_begin = rows;
    memcpy(s.anchors, scales[i].anchors, sizeof(float) * 2 * s.na);
    rows += s.h * s.w * s.na;
  }
  auto go = [&]() -> int {
    YB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    YB_TRY(p->ctx.alloc(max_batch, rows, 5 + num_classes, decode_mode == YB_DECODE_V2));
    p->ctx.n_scales = n_scales;
    return YB_OK;
  };
  int r = go();
  if (r != YB_OK) { yb_post_destroy(p); return r; }
  *out = p;
  return YB_OK;
}

void yb_post_destroy(yb_post* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  p->ctx.release();
  cudaFree(p->staging);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

static int post_stage_and_decode(yb_post* p, const float* head, int mem, int n, double threshold) {
  if (!p || !head) return fail(YB_ERR_INVALID, "yb_post: bad argument");
  if (n <= 0 || n > p->ctx.max_batch) return fail(YB_ERR_INVALID, "batch %d outside [1,%d]", n, p->ctx.max_batch);
  YB_TRY(set_device(p->device));
  const size_t floats = (size_t)n * p->ctx.rows * p->ctx.box_len;
  const float* dev = head;
  if (mem == YB_MEM_HOST) {
    if (floats > p->staging_floats) {
      cudaFree(p->staging); p->staging = nullptr; p->staging_floats = 0;
      YB_CUDA(cudaMalloc(&p->staging, floats * 4));
      p->staging_floats = floats;
    }
    YB_CUDA(cudaMemcpyAsync(p->staging, head, floats * 4, cudaMemcpyHostToDevice, p->stream));
    dev = p->staging;
  }
  ScaleDesc sc[POST_MAX_SCALES];
  for (int i = 0; i < p->ctx.n_scales; ++i) {
    sc[i] = p->ctx.scales[i];
    sc[i].base = dev + (long long)sc[i].row_begin * p->ctx.box_len;
    sc[i].img_stride = (long long)p->ctx.rows * p->ctx.box_len;
    sc[i].cell_stride = sc[i].na * p->ctx.box_len;
  }
  YB_CUDA(cudaEventRecord(p->ctx.ev[0], p->stream));
  YB_TRY(p->ctx.decode(p->stream, sc, n, (float)threshold));      // float32 compare: NEP 50 (see yb_engine_detect_async)
  YB_CUDA(cudaEventRecord(p->ctx.ev[1], p->stream));
  return YB_OK;
}

int yb_post_run(yb_post* p, const float* head, int mem, int n, double threshold, double iou_threshold, int nms_mode,
                yb_det* out, int* counts, int max_per_image, int* cand_counts) {
  if (nms_mode != YB_NMS_REFERENCE && nms_mode != YB_NMS_PER_CLASS) return fail(YB_ERR_INVALID, "unknown nms mode %d", nms_mode);
  YB_TRY(post_stage_and_decode(p, head, mem, n, threshold));
  YB_TRY(p->ctx.nms(p->stream, n, iou_threshold, nms_mode));
  YB_CUDA(cudaEventRecord(p->ctx.ev[2], p->stream));
  p->timed = true;
  if (out && counts) {
    YB_TRY(copy_dets(p->ctx, p->stream, n, out, counts, max_per_image, nullptr));
  }
  if (cand_counts) {
    YB_CUDA(cudaMemcpyAsync(cand_counts, p->ctx.n_cand, (size_t)n * 4, cudaMemcpyDeviceToHost, p->stream));
    YB_CUDA(cudaStreamSynchronize(p->stream));
  }
  return YB_OK;
}

// identity "NMS": order = every candidate row (ascending), used to export the raw decode
__global__ void list_candidates_kernel(int rows, const float* prob, int* order, int* n_keep) {
  __shared__ int s_cnt;
  const int img = blockIdx.x;
  if (threadIdx.x == 0) {
    int c = 0;
    for (int r = 0; r < rows; ++r)
      { const float pr = prob[(long long)img * rows + r]; if (pr == pr) order[(long long)img * rows + c++] = r; }
    s_cnt = c;
    n_keep[img] = c;
  }
}

int yb_post_decode(yb_post* p, const float* head, int mem, int n, double threshold, yb_det* out, int* counts, int max_per_image) {
  if (!out || !counts) return fail(YB_ERR_INVALID, "yb_post_decode: output buffers are NULL");
  YB_TRY(post_stage_and_decode(p, head, mem, n, threshold));
  list_candidates_kernel<<<n, 32, 0, p->stream>>>(p->ctx.rows, p->ctx.prob, p->ctx.order, p->ctx.n_keep);
  YB_CUDA(cudaGetLastError());
  return copy_dets(p->ctx, p->stream, n, out, counts, max_per_image, nullptr);
}

int yb_post_order_after(yb_post* p, void* cuda_stream) {
  if (!p) return fail(YB_ERR_INVALID, "post context is NULL");
  YB_TRY(set_device(p->device));
  return order_streams(static_cast<cudaStream_t>(cuda_stream), p->stream);
}
int yb_post_order_before(yb_post* p, void* cuda_stream) {
  if (!p) return fail(YB_ERR_INVALID, "post context is NULL");
  YB_TRY(set_device(p->device));
  return order_streams(p->stream, static_cast<cudaStream_t>(cuda_stream));
}

int yb_post_sync(yb_post* p) {
  if (!p) return fail(YB_ERR_INVALID, "post is NULL");
  YB_TRY(set_device(p->device));
  YB_CUDA(cudaStreamSynchronize(p->stream));
  return YB_OK;
}

int yb_post_last_ms(yb_post* p, float* decode_ms, float* nms_ms) {
  if (!p || !p->timed) return fail(YB_ERR_STATE, "yb_post_last_ms before yb_post_run");
  YB_TRY(set_device(p->device));
  YB_CUDA(cudaEventSynchronize(p->ctx.ev[2]));
  float a = 0.f, b = 0.f;
  YB_CUDA(cudaEventElapsedTime(&a, p->ctx.ev[0], p->ctx.ev[1]));
  YB_CUDA(cudaEventElapsedTime(&b, p->ctx.ev[1], p->ctx.ev[2]));
  if (decode_ms) *decode_ms = a;
  if (nms_ms) *nms_ms = b;
  return YB_OK;
}

int yb_nms(const void* x, const void* y, const void* w, const void* h, const float* prob, const int32_t* class_idx, int k,
           int f64, double iou_threshold, int nms_mode, int device, int32_t* keep, int* n_keep) {
  if (!keep || !n_keep || k < 0) return fail(YB_ERR_INVALID, "yb_nms: bad argument");
  if (nms_mode != YB_NMS_REFERENCE && nms_mode != YB_NMS_PER_CLASS) return fail(YB_ERR_INVALID, "unknown nms mode %d", nms_mode);
  if (nms_mode == YB_NMS_PER_CLASS && !class_idx) return fail(YB_ERR_INVALID, "per-class NMS needs class_idx");
  if (k == 0) { *n_keep = 0; return YB_OK; }          // net/base.py:196-197
  if (!x || !y || !w || !h || !prob) return fail(YB_ERR_INVALID, "yb_nms: NULL input array");
  YB_TRY(set_device(device));
  PostCtx ctx;
  int r = ctx.alloc(1, k, 5, 0);
  cudaStream_t st = nullptr;
  void *dx = nullptr, *dy = nullptr, *dw = nullptr, *dh = nullptr;
  auto go = [&]() -> int {
    YB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t es = f64 ? 8 : 4;
    YB_CUDA(cudaMalloc(&dx, k * es)); YB_CUDA(cudaMalloc(&dy, k * es)); YB_CUDA(cudaMalloc(&dw, k * es)); YB_CUDA(cudaMalloc(&dh, k * es));
    YB_CUDA(cudaMemcpyAsync(dx, x, k * es, cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaMemcpyAsync(dy, y, k * es, cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaMemcpyAsync(dw, w, k * es, cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaMemcpyAsync(dh, h, k * es, cudaMemcpyHostToDevice, st));
    // a NaN score is the engine's "not a candidate" marker, so it cannot be a score here
    for (int i = 0; i < k; ++i)
      if (prob[i] != prob[i]) return fail(YB_ERR_INVALID, "yb_nms: box %d has a NaN score", i);
    YB_CUDA(cudaMemcpyAsync(ctx.prob, prob, (size_t)k * 4, cudaMemcpyHostToDevice, st));
    if (class_idx) YB_CUDA(cudaMemcpyAsync(ctx.cls, class_idx, (size_t)k * 4, cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaMemsetAsync(ctx.flags, 0, (size_t)k, st));
    NmsArgs a = ctx.nms_args(f64 ? iou_threshold : (double)(float)iou_threshold, nms_mode == YB_NMS_PER_CLASS);
    a.x = dx; a.y = dy; a.w = dw; a.h = dh;
    a.cls = class_idx ? ctx.cls : nullptr;
    if (f64) YB_TRY((ctx.nms_launch<double, true, true>(st, 1, a)));
    else YB_TRY((ctx.nms_launch<float, false, false>(st, 1, a)));
    YB_CUDA(cudaMemcpyAsync(n_keep, ctx.n_keep, 4, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    YB_CUDA(cudaMemcpy(keep, ctx.order, (size_t)(*n_keep) * 4, cudaMemcpyDeviceToHost));
    return YB_OK;
  };
  if (r == YB_OK) r = go();
  cudaFree(dx); cudaFree(dy); cudaFree(dw); cudaFree(dh);
  if (st) cudaStreamDestroy(st);
  ctx.release();
  return r;
}

}  // extern "C"
