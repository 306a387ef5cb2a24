// post.cuh -- YOLO head decode and greedy NMS on the device.
//
// decode_kernel        replaces net/v3.py:109-136 and net/v2.py:93-119 (one warp per box: coalesced
//                      read of its 5+C logits, shuffle argmax, threshold, box maths).
// sort_nms_kernel      replaces net/base.py:195-209 + :180-192 (one CTA per image: bitonic sort by
//                      (score desc, row asc) == Python's stable reverse sort over boxes decoded in
//                      raster order; then a blocked greedy sweep: 64 sorted boxes at a time are
//                      resolved with a 64x64 bit mask, and every newly kept box suppresses all later
//                      boxes in parallel -- the same result as the sequential reference loop).
// The IoU uses the reference's exact operation order in the reference's precision (float64 on
// float32 x,y and float64 w,h -- see DESIGN.md) with non-contracting intrinsics, so the kept set is
// bit-identical to the reference's when both start from the same candidates.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace yb {

constexpr int POST_MAX_SCALES = 4;
constexpr int POST_MAX_ANCHORS = 16;

struct ScaleDesc {
  const float* base;     // first logit of this scale for image 0
  long long img_stride;  // floats between images
  int cell_stride;       // floats between grid cells
  int h, w, na;
  int row_begin;         // first global row of this scale
  float anchors[2 * POST_MAX_ANCHORS];  // grid units, (w,h)
};

struct DecodeArgs {
  ScaleDesc sc[POST_MAX_SCALES];
  int n_scales;
  int rows;         // R = rows per image
  int box_len;      // 5 + C
  int n_images;
  int v2;           // 1: softmax classes, score = obj * max softmax; 0: sigmoid, score = obj
  float thr;
  // dense per-row outputs [n_images * rows]
  float* prob;      // score, or NaN when the row is not a candidate
  float* x;
  float* y;
  double* w;
  double* h;
  int* cls;
};

__device__ __forceinline__ float sigmoid_ref(float v) {   // 1. / (1. + np.exp(-x)) in float32
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
}

// one warp per (image, row)
__global__ void __launch_bounds__(256) decode_kernel(const DecodeArgs a) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)a.n_images * a.rows;
  if (gw >= total) return;
  const int img = (int)(gw / a.rows);
  const int row = (int)(gw - (long long)img * a.rows);
  int s = 0;
#pragma unroll
  for (int i = 1; i < POST_MAX_SCALES; ++i)
    if (i < a.n_scales && row >= a.sc[i].row_begin) s = i;
  const ScaleDesc& sc = a.sc[s];
  const int local = row - sc.row_begin;
  const int cell = local / sc.na;
  const int anc = local - cell * sc.na;
  const float* p = sc.base + (long long)img * sc.img_stride + (long long)cell * sc.cell_stride + anc * a.box_len;

  // lane l holds logits l, l+32, l+64, ... ; logits 0..4 sit in lanes 0..4 of the first round
  const int len = a.box_len;
  float first = (lane < len) ? __ldg(p + lane) : 0.0f;
  const float t_obj = __shfl_sync(0xffffffffu, first, 4);
  const float obj = sigmoid_ref(t_obj);
  const long long o = (long long)img * a.rows + row;
  if (!a.v2 && obj < a.thr) {       // v3: score is the objectness alone (net/v3.py:123-125)
    if (lane == 0) a.prob[o] = __int_as_float(0x7fc00000);
    return;
  }
  // class scores: argmax with ties to the lowest class index (np.argmax)
  float best = -INFINITY;
  int best_k = 0x7fffffff;
  float score;
  if (!a.v2) {
    for (int k = lane; k < len; k += 32) {
      const float t = (k == lane) ? first : __ldg(p + k);
      if (k >= 5) {
        const float pc = sigmoid_ref(t);
        if (pc > best) { best = pc; best_k = k - 5; }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, d);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, d);
      if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    score = obj;
  } else {
    // softmax (net/base.py:175-177): e = exp(x - max); e / sum(e)
    float mx = -INFINITY;
    for (int k = lane; k < len; k += 32) {
      const float t = (k == lane) ? first : __ldg(p + k);
      if (k >= 5) mx = fmaxf(mx, t);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    float sum = 0.0f;
    for (int k = lane; k < len; k += 32) {
      const float t = (k == lane) ? first : __ldg(p + k);
      if (k >= 5) {
        const float e = expf(__fsub_rn(t, mx));
        sum = __fadd_rn(sum, e);
        if (e > best) { best = e; best_k = k - 5; }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, d));
      const float ob = __shfl_xor_sync(0xffffffffu, best, d);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, d);
      if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    const float class_prob = __fdiv_rn(best, sum);
    score = __fmul_rn(obj, class_prob);              // net/v2.py:106
    if (score < a.thr) {
      if (lane == 0) a.prob[o] = __int_as_float(0x7fc00000);
      return;
    }
  }
  const float t0 = __shfl_sync(0xffffffffu, first, 0);
  const float t1 = __shfl_sync(0xffffffffu, first, 1);
  const float t2 = __shfl_sync(0xffffffffu, first, 2);
  const float t3 = __shfl_sync(0xffffffffu, first, 3);
  if (lane == 0) {
    const int cy = cell / sc.w;
    const int cx = cell - cy * sc.w;
    a.prob[o] = score;
    a.x[o] = __fdiv_rn(__fadd_rn(sigmoid_ref(t0), (float)cx), (float)sc.w);   // (sigmoid(tx) + cx) / w
    a.y[o] = __fdiv_rn(__fadd_rn(sigmoid_ref(t1), (float)cy), (float)sc.h);
    // anchors are float64 in the reference: (aw * exp(tw)) / w evaluated in float64
    a.w[o] = __ddiv_rn(__dmul_rn((double)sc.anchors[2 * anc], (double)expf(t2)), (double)sc.w);
    a.h[o] = __ddiv_rn(__dmul_rn((double)sc.anchors[2 * anc + 1], (double)expf(t3)), (double)sc.h);
    a.cls[o] = best_k;
  }
}

// YOLOv3 decode, restructured around the fact that the keep/drop decision needs ONE logit per row (net/v3.py:123-125):
//   phase 1  one lane per row: load t_obj, score = sigmoid(t_obj), write the dense score (NaN = not a candidate);
//   phase 2  the warp walks its candidate rows (ballot) and decodes each cooperatively: coalesced read of the row,
//            class argmax, box maths spread over four lanes.
// Sparse heads (real images: a few candidates per image) therefore touch one 32-byte sector per row instead of the
// whole row.  The class argmax works on d = 1 + exp(-t) instead of sigmoid(t) = 1 / d: the correctly rounded division is
// monotone, so argmax sigmoid == argmin d, and np.argmax's "first of the maxima" only needs the quotient for the rare
// near-ties of d (one division per row instead of one per class).
__global__ void __launch_bounds__(256) decode_v3_kernel(const DecodeArgs a) {
  const int lane = threadIdx.x & 31;
  const long long total = (long long)a.n_images * a.rows;
  const long long warp_first = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5;   // first row of this warp
  if (warp_first >= total) return;
  const long long gr = warp_first + lane;
  const int len = a.box_len;
  // ---- phase 1 ----
  const float* p = nullptr;
  int cell = 0, anc = 0, s = 0;
  bool cand = false;
  if (gr < total) {
    const int img = (int)(gr / a.rows);
    const int row = (int)(gr - (long long)img * a.rows);
#pragma unroll
    for (int i = 1; i < POST_MAX_SCALES; ++i)
      if (i < a.n_scales && row >= a.sc[i].row_begin) s = i;
    const ScaleDesc& sc = a.sc[s];
    const int local = row - sc.row_begin;
    cell = local / sc.na;
    anc = local - cell * sc.na;
    p = sc.base + (long long)img * sc.img_stride + (long long)cell * sc.cell_stride + anc * len;
    const float obj = sigmoid_ref(__ldg(p + 4));
    cand = obj >= a.thr;
    a.prob[gr] = cand ? obj : __int_as_float(0x7fc00000);
  }
  // ---- phase 2a: box maths, one lane per candidate row (the five box logits share the sector t_obj came from) ----
  if (cand) {
    const ScaleDesc& sc = a.sc[s];
    const float t0 = __ldg(p), t1 = __ldg(p + 1), t2 = __ldg(p + 2), t3 = __ldg(p + 3);
    const int cy = cell / sc.w;
    const int cx = cell - cy * sc.w;
    a.x[gr] = __fdiv_rn(__fadd_rn(sigmoid_ref(t0), (float)cx), (float)sc.w);      // (sigmoid(tx) + cx) / w
    a.y[gr] = __fdiv_rn(__fadd_rn(sigmoid_ref(t1), (float)cy), (float)sc.h);
    // anchors are float64 in the reference: (aw * exp(tw)) / w evaluated in float64
    a.w[gr] = __ddiv_rn(__dmul_rn((double)sc.anchors[2 * anc], (double)expf(t2)), (double)sc.w);
    a.h[gr] = __ddiv_rn(__dmul_rn((double)sc.anchors[2 * anc + 1], (double)expf(t3)), (double)sc.h);
  }
  // ---- phase 2b: class argmax, four candidate rows at a time, eight lanes per row ----
  unsigned mask = __ballot_sync(0xffffffffu, cand);
  const int sub = lane >> 3, sl = lane & 7;
  const unsigned gmask = 0xFFu << (sub * 8);
  while (mask) {
    int my_src = -1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int src = mask ? __ffs(mask) - 1 : -1;
      if (mask) mask &= mask - 1;
      if (q == sub) my_src = src;
    }
    const float* rp = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)p, my_src < 0 ? 0 : my_src));
    // per lane: smallest d = 1 + exp(-t) of its classes (lowest index on ties) and the runner-up
    float dmin = INFINITY, d2 = INFINITY;
    int kmin = 0x7fffffff;
    bool saw_nan = false;            // a NaN logit is np.argmax's maximum: it must reach the exact path, not be skipped
    if (my_src >= 0) {
#pragma unroll 5
      for (int k = 5 + sl; k < len; k += 8) {
        // fast exponential (two instructions instead of eight): the ranking only has to be right up to the tie band
        // below, which is far wider than its error (|t| * 1.2e-7 + 2^-22 relative, |t| < 88 where d is finite and > 1)
        const float d = __fadd_rn(1.0f, __expf(-__ldg(rp + k)));
        saw_nan |= (d != d);
        if (d < dmin) { d2 = dmin; dmin = d; kmin = k - 5; }
        else if (d < d2) d2 = d;
      }
    }
    float wd = dmin;
    int wk = kmin;
#pragma unroll
    for (int sft = 4; sft > 0; sft >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, wd, sft);
      const int ok = __shfl_xor_sync(0xffffffffu, wk, sft);
      if (od < wd || (od == wd && ok < wk)) { wd = od; wk = ok; }
    }
    // Near-ties: another class whose d is within the error of the fast exponential (or a few ulps) of the minimum may be
    // the true maximum, or round to the same sigmoid and, having a lower index, be np.argmax's answer; so may any class
    // once 1/d is subnormal.  Rare (band 1e-4 relative): only then are the exact quotients computed, with the accurate
    // exponential (by the eight lanes of that row).
    const float near = wd * 1.0001f;
    const bool tie = my_src >= 0 && (saw_nan || (kmin != wk && dmin <= near) || (d2 <= near) || !(wd < 1e37f));
    const unsigned ties = __ballot_sync(0xffffffffu, tie);
    int best_k = wk;
    if (ties & gmask) {
      // the reference's own arithmetic: first index of the largest 1 / (1 + exp(-t)) in float32
      float smax = -1.0f;
      int cand_k = 0x7fffffff;
      bool any_nan = false;                                        // np.argmax: the first NaN is the maximum
      if (my_src >= 0) {
        for (int k = 5 + sl; k < len; k += 8) {
          const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-__ldg(rp + k))));
          if (sg != sg) { if (!any_nan) { any_nan = true; cand_k = k - 5; } }
          else if (!any_nan && sg > smax) { smax = sg; cand_k = k - 5; }          // ascending k: a later equal value never replaces
        }
      }
#pragma unroll
      for (int sft = 4; sft > 0; sft >>= 1) {
        const float os = __shfl_xor_sync(gmask, smax, sft);
        const int ok = __shfl_xor_sync(gmask, cand_k, sft);
        const bool on = __shfl_xor_sync(gmask, any_nan, sft);
        const bool take = (on != any_nan) ? on : (on ? ok < cand_k : (os > smax || (os == smax && ok < cand_k)));
        if (take) { smax = os; cand_k = ok; any_nan = on; }
      }
      best_k = cand_k;
    }
    if (my_src >= 0 && sl == 0) a.cls[warp_first + my_src] = best_k;
  }
}

// YOLOv3 decode for DENSE candidate sets on a contiguous head tensor [n_images * rows][5 + C] (the stand-alone layout of
// yb_post_run, BASELINE config 5: score threshold 0.001, every row a candidate).  The kernel above reads a row with
// eight lanes and 4-byte loads and is issue-bound at ~75 warp instructions per row; here a warp stages its 32 rows
// (32 * len * 4 bytes, contiguous and 16-byte aligned because the slab starts at a multiple of 32 rows) in shared memory
// with 16-byte loads and then works one lane per row:
//   * the class argmax needs no exponential: sigmoid is monotone, so argmax sigmoid(t) = argmax t unless the two largest
//     logits are so close -- or so saturated -- that their float32 sigmoids could round to the same value, in which case
//     np.argmax's "first of the maxima" is decided by the reference's own arithmetic, computed by the whole warp for that
//     row.  Separation bound: the computed sigmoid (expf <= 2 ulp, one add, one division) is within 2.4e-7 relative of
//     the true one, and sigmoid(a) / sigmoid(b) = 1 + (1 - sigmoid(a)) (e^(a-b) - 1) >= 1 + 9.1e-4 * 1e-3 for
//     a - b >= 1e-3 and a <= 7, four times the rounding noise.  Outside [-80, 7], or with a NaN in the row: exact path.
//   * rows are `len` words apart: an odd len (85) makes the lane-per-row reads bank-conflict free.
constexpr int DECODE_BULK_WARPS = 4;
__global__ void __launch_bounds__(DECODE_BULK_WARPS * 32) decode_v3_bulk_kernel(const DecodeArgs a, const float* __restrict__ head) {
  extern __shared__ __align__(16) float bulk_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int len = a.box_len;
  float* slab = bulk_smem + (size_t)warp * 32 * len;
  const long long total = (long long)a.n_images * a.rows;
  const long long warp_first = ((long long)blockIdx.x * DECODE_BULK_WARPS + warp) << 5;
  if (warp_first >= total) return;
  const int n_rows = (int)((total - warp_first < 32) ? (total - warp_first) : 32);
  const float* src = head + warp_first * len;
  const int words = n_rows * len;
  {
    // asynchronous 16-byte copies straight into shared memory: the whole slab (10.9 KB at len = 85) is in flight at once
    const int vec = words >> 2;                     // warp_first * len * 4 bytes is a multiple of 16: 32 rows per slab
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab);
    for (int i = lane; i < vec; i += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 16u), "l"(src + 4 * i) : "memory");
    for (int i = (vec << 2) + lane; i < words; i += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)i * 4u), "l"(src + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncwarp();
  const long long gr = warp_first + lane;
  const bool live = lane < n_rows;
  const float* my = slab + lane * len;
  bool cand = false, tie = false;
  int best_k = 0;
  if (live) {
    const int img = (int)(gr / a.rows);
    const int row = (int)(gr - (long long)img * a.rows);
    int s = 0;
#pragma unroll
    for (int i = 1; i < POST_MAX_SCALES; ++i)
      if (i < a.n_scales && row >= a.sc[i].row_begin) s = i;
    const ScaleDesc& sc = a.sc[s];
    const int local = row - sc.row_begin;
    const int cell = local / sc.na;
    const int anc = local - cell * sc.na;
    const float obj = sigmoid_ref(my[4]);
    cand = obj >= a.thr;
    a.prob[gr] = cand ? obj : __int_as_float(0x7fc00000);
    if (cand) {
      const int cy = cell / sc.w;
      const int cx = cell - cy * sc.w;
      a.x[gr] = __fdiv_rn(__fadd_rn(sigmoid_ref(my[0]), (float)cx), (float)sc.w);
      a.y[gr] = __fdiv_rn(__fadd_rn(sigmoid_ref(my[1]), (float)cy), (float)sc.h);
      a.w[gr] = __ddiv_rn(__dmul_rn((double)sc.anchors[2 * anc], (double)expf(my[2])), (double)sc.w);
      a.h[gr] = __ddiv_rn(__dmul_rn((double)sc.anchors[2 * anc + 1], (double)expf(my[3])), (double)sc.h);
      // four independent (max, runner-up, index) trackers over k mod 4 keep the compare chains short; merged afterwards
      // with the lower index winning among equal logits (np.argmax: the first of the maxima)
      float m1[4], m2[4];
      int bk[4];
      bool nan = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) { m1[j] = -INFINITY; m2[j] = -INFINITY; bk[j] = 0x7fffffff; }
      int k = 5;
      for (; k + 4 <= len; k += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = my[k + j];
          nan |= (t != t);
          const bool gt = t > m1[j];                           // strict: the first of equal logits stays
          m2[j] = gt ? m1[j] : fmaxf(m2[j], t);
          bk[j] = gt ? k + j - 5 : bk[j];
          m1[j] = gt ? t : m1[j];
        }
      }
      for (int j = 0; k < len; ++k, ++j) {
        const float t = my[k];
        nan |= (t != t);
        const bool gt = t > m1[j];
        m2[j] = gt ? m1[j] : fmaxf(m2[j], t);
        bk[j] = gt ? k - 5 : bk[j];
        m1[j] = gt ? t : m1[j];
      }
      float top = m1[0], second = m2[0];
      best_k = bk[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        const bool win = m1[j] > top || (m1[j] == top && bk[j] < best_k);
        second = fmaxf(fmaxf(second, m2[j]), win ? top : m1[j]);
        best_k = win ? bk[j] : best_k;
        top = win ? m1[j] : top;
      }
      tie = nan || !(top - second >= 1e-3f) || !(top <= 7.0f) || !(top >= -80.0f);
    }
  }
  // exact path for the rare near-tie rows: the reference's float32 sigmoids, first index of the largest, by all 32 lanes
  unsigned ties = __ballot_sync(0xffffffffu, tie);
  while (ties) {
    const int r = __ffs(ties) - 1;
    ties &= ties - 1;
    const float* rp = slab + r * len;
    float smax = -1.0f;
    int cand_k = 0x7fffffff;
    bool any_nan = false;
    for (int k = 5 + lane; k < len; k += 32) {
      const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rp[k])));
      if (sg != sg) { if (!any_nan) { any_nan = true; cand_k = k - 5; } }        // np.argmax: the first NaN wins
      else if (!any_nan && sg > smax) { smax = sg; cand_k = k - 5; }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, smax, sft);
      const int ok = __shfl_xor_sync(0xffffffffu, cand_k, sft);
      const bool on = __shfl_xor_sync(0xffffffffu, any_nan, sft);
      const bool take = (on != any_nan) ? on : (on ? ok < cand_k : (os > smax || (os == smax && ok < cand_k)));
      if (take) { smax = os; cand_k = ok; any_nan = on; }
    }
    if (lane == r) best_k = cand_k;
  }
  if (live && cand) a.cls[gr] = best_k;
}

// ----------------------------------------------------------------------------------------------
// sort + NMS
// ----------------------------------------------------------------------------------------------
template <typename T> struct Arith;
template <> struct Arith<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <> struct Arith<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
// numpy's maximum/minimum propagate NaN
template <typename T> __device__ __forceinline__ T np_max(T a, T b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
template <typename T> __device__ __forceinline__ T np_min(T a, T b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

template <typename T>
struct BoxC {  // corners + area, net/base.py:181-184,266-272
  T x1, y1, x2, y2, area;
};
template <typename T>
__device__ __forceinline__ BoxC<T> make_box(T x, T y, T w, T h) {
  using A = Arith<T>;
  BoxC<T> b;
  const T hw = A::mul(w, (T)0.5), hh = A::mul(h, (T)0.5);   // w / 2. is exact
  b.x1 = A::sub(x, hw); b.y1 = A::sub(y, hh);
  b.x2 = A::add(x, hw); b.y2 = A::add(y, hh);
  b.area = A::mul(w, h);
  return b;
}
// net/base.py:186-192, same operation order.  SAFE selects numpy's NaN-propagating min/max.
template <typename T, bool SAFE>
__device__ __forceinline__ T iou_ref(const BoxC<T>& p, const BoxC<T>& q) {
  using A = Arith<T>;
  T ix1, iy1, ix2, iy2;
  if (SAFE) {
    ix1 = np_max(p.x1, q.x1); iy1 = np_max(p.y1, q.y1);
    ix2 = np_min(p.x2, q.x2); iy2 = np_min(p.y2, q.y2);
  } else {
    ix1 = p.x1 > q.x1 ? p.x1 : q.x1; iy1 = p.y1 > q.y1 ? p.y1 : q.y1;
    ix2 = p.x2 < q.x2 ? p.x2 : q.x2; iy2 = p.y2 < q.y2 ? p.y2 : q.y2;
  }
  T iw = A::sub(ix2, ix1), ih = A::sub(iy2, iy1);
  if (SAFE) { iw = np_max(iw, (T)0); ih = np_max(ih, (T)0); }
  else { iw = iw > (T)0 ? iw : (T)0; ih = ih > (T)0 ? ih : (T)0; }
  const T inter = A::mul(iw, ih);
  T uni = A::sub(A::add(p.area, q.area), inter);
  if (SAFE) uni = np_max(uni, (T)1e-8);
  else uni = uni > (T)1e-8 ? uni : (T)1e-8;
  return A::div(inter, uni);
}
template <typename T>
__device__ __forceinline__ bool box_finite(const BoxC<T>& b) {
  return isfinite(b.x1) && isfinite(b.y1) && isfinite(b.x2) && isfinite(b.y2) && isfinite(b.area);
}

// The reference's decision "iou_score(p, q) >= thr" (strict: ">", the per-class mode) for two FINITE boxes without
// paying for the division in the common case: with u the unit round-off, fl(inter/uni) and fl(thr*uni) are each
// within a factor (1 +- u) of the exact values, so whenever inter differs from fl(thr*uni) by more than a relative
// margin m >> u the comparison of the rounded quotient with thr is decided; only the rare near-ties take the exact
// division, which keeps every decision bit-identical to iou_ref.
template <typename T>
__device__ __forceinline__ bool suppresses_finite(const BoxC<T>& p, const BoxC<T>& q, T thr, bool strict) {
  using A = Arith<T>;
  const T ix1 = p.x1 > q.x1 ? p.x1 : q.x1, iy1 = p.y1 > q.y1 ? p.y1 : q.y1;
  const T ix2 = p.x2 < q.x2 ? p.x2 : q.x2, iy2 = p.y2 < q.y2 ? p.y2 : q.y2;
  T iw = A::sub(ix2, ix1), ih = A::sub(iy2, iy1);
  iw = iw > (T)0 ? iw : (T)0;
  ih = ih > (T)0 ? ih : (T)0;
  const T inter = A::mul(iw, ih);
  T uni = A::sub(A::add(p.area, q.area), inter);
  uni = uni > (T)1e-8 ? uni : (T)1e-8;
  if (thr > (T)0) {
    const T m = sizeof(T) == 8 ? (T)1e-12 : (T)1e-5;
    const T tq = A::mul(thr, uni);
    const T d = A::sub(inter, tq), band = A::mul(m, tq);
    if (d > band) return true;
    if (d < -band) return false;
  }
  const T v = A::div(inter, uni);
  return strict ? (v > thr) : (v >= thr);
}

// float32 interval arithmetic on the float64 box (T = double only).  A box carries outward- and inward-rounded float
// copies of its corners and down-/up-rounded copies of its area, so [lo, hi] float bounds of the REAL-arithmetic
// intersection and union follow with directed-rounding float ops; the float64 result iou_ref computes differs from
// the real-arithmetic IoU by a relative ~1e-15, so a bound that clears thr by a relative 1e-6 decides the comparison
// for certain.  Returns +1 (suppresses), -1 (does not), 0 (too close: take the exact float64 path).
struct BoxF {
  float4 out;     // rd(x1), rd(y1), ru(x2), ru(y2)
  float4 in;      // ru(x1), ru(y1), rd(x2), rd(y2)
  float2 area;    // rd(area), ru(area)
};
__device__ __forceinline__ BoxF make_boxf(double x1, double y1, double x2, double y2, double area) {
  BoxF f;
  f.out = make_float4(__double2float_rd(x1), __double2float_rd(y1), __double2float_ru(x2), __double2float_ru(y2));
  f.in = make_float4(__double2float_ru(x1), __double2float_ru(y1), __double2float_rd(x2), __double2float_rd(y2));
  f.area = make_float2(__double2float_rd(area), __double2float_ru(area));
  return f;
}
__device__ __forceinline__ int decide_f32(const float4& po, const float4& pi, const float2& pa, const float4& qo, const float4& qi,
                                          const float2& qa, float thr_dn, float thr_up) {
  const float iw_hi = fmaxf(__fsub_ru(fminf(po.z, qo.z), fmaxf(po.x, qo.x)), 0.0f);
  const float ih_hi = fmaxf(__fsub_ru(fminf(po.w, qo.w), fmaxf(po.y, qo.y)), 0.0f);
  const float iw_lo = fmaxf(__fsub_rd(fminf(pi.z, qi.z), fmaxf(pi.x, qi.x)), 0.0f);
  const float ih_lo = fmaxf(__fsub_rd(fminf(pi.w, qi.w), fmaxf(pi.y, qi.y)), 0.0f);
  const float inter_hi = __fmul_ru(iw_hi, ih_hi), inter_lo = __fmul_rd(iw_lo, ih_lo);
  const float uni_hi = fmaxf(__fsub_ru(__fadd_ru(pa.y, qa.y), inter_lo), 1.0000001e-8f);
  const float uni_lo = fmaxf(__fsub_rd(__fadd_rd(pa.x, qa.x), inter_hi), 9.9999999e-9f);
  if (inter_lo > __fmul_ru(thr_up, uni_hi)) return 1;      // IoU >= inter_lo / uni_hi > thr (1 + 1e-6)
  if (inter_hi < __fmul_rd(thr_dn, uni_lo)) return -1;     // IoU <= inter_hi / uni_lo < thr (1 - 1e-6)
  return 0;
}

// Outward-rounded float copies of a box's corners: if two such intervals do not overlap, neither do the exact ones.
__device__ __forceinline__ float4 conservative_f4(double x1, double y1, double x2, double y2) {
  return make_float4(__double2float_rd(x1), __double2float_rd(y1), __double2float_ru(x2), __double2float_ru(y2));
}
// true: the boxes certainly share no area (intersection width or height <= 0), i.e. inter == 0 and iou == 0
__device__ __forceinline__ bool surely_disjoint(const float4& a, const float4& b) {
  return a.z <= b.x || b.z <= a.x || a.w <= b.y || b.w <= a.y;
}

// Necessary condition for "iou >= thr" between two REGULAR float64 boxes (finite, positive extent, area consistent with
// the corners, extent not vanishing against the coordinates): with W = x2 - x1,
//   inter <= min(area_p, area_q) (1 + 1e-9)  =>  union >= max(area_p, area_q) (1 - 1.1e-9)
//   fl(inter / union) >= thr                =>  iw * ih >= thr * area_q (1 - 3e-9), and ih <= H_q
//                                           =>  min(x2p, x2q) - max(x1p, x1q) >= thr * W_q (1 - 4e-9)
// i.e. in each dimension either box must reach at least g = min(thr, 1) * (1 - 1e-6) * extent into the other one:
// x2p >= x1q + g_q and x1p <= x2q - g_q (and the same with p and q swapped).  Every box therefore carries, next to
// its outward-rounded corners, the corners "shrunk" by g (lower ones rounded down, upper ones rounded up, so the
// shrunk box of a wide box is an inverted interval: the band the other box has to cover).  The 1e-6 margin covers the
// float64 rounding of x1 + g (<= 1.2e-12 W for a regular box) for every thr >= 1e-2; below that g is 0 and the test is
// plain disjointness.  Irregular boxes are flagged like non-finite ones and always take the exact path.
__device__ __forceinline__ float4 shrunk_f4(double x1, double y1, double x2, double y2, double t) {
  const double gx = __dmul_rn(t, __dsub_rn(x2, x1)), gy = __dmul_rn(t, __dsub_rn(y2, y1));
  return make_float4(__double2float_rd(__dadd_rn(x1, gx)), __double2float_rd(__dadd_rn(y1, gy)),
                     __double2float_ru(__dsub_rn(x2, gx)), __double2float_ru(__dsub_rn(y2, gy)));
}
__device__ __forceinline__ bool box_regular(double x1, double y1, double x2, double y2, double area) {
  const double W = __dsub_rn(x2, x1), H = __dsub_rn(y2, y1), wh = __dmul_rn(W, H);
  return W > 0.0 && H > 0.0 && area > 0.0 && fabs(wh - area) <= 1e-9 * area &&
         W >= 1e-4 * fmax(fabs(x1), fabs(x2)) && H >= 1e-4 * fmax(fabs(y1), fabs(y2));
}
// true: box p (outward corners po, shrunk ps) and box q can certainly not reach iou >= thr
__device__ __forceinline__ bool surely_below(const float4& qo, const float4& qs, const float4& po, const float4& ps) {
  return (qs.z <= po.x) | (po.z <= qs.x) | (qs.w <= po.y) | (po.w <= qs.y) |
         (qo.z <= ps.x) | (ps.z <= qo.x) | (qo.w <= ps.y) | (ps.w <= qo.y);
}
struct NmsArgs {
  int rows;             // R: dense rows per image (stride of every per-row array)
  int per_class;        // 0: reference (class-agnostic, >=); 1: per class, strict >
  int debug;            // measurement only (YB_NMS_DEBUG): bit 0 skips the sweep of the pipelined variant (wrong results)
  double thr;
  // dense inputs [n * rows]; a NaN prob marks a non-candidate
  const float* prob;
  const void* x;        // float or double (XY64)
  const void* y;
  const void* w;        // float or double (WH64)
  const void* h;
  const int* cls;       // may be null when !per_class
  // scratch [n * rows_pow2] / [n * rows]
  unsigned long long* keys;   // global sort scratch, rows_pow2 per image (used when K_pow2 > smem keys)
  int rows_pow2;
  void* sorted_boxes;   // BoxC<T> [n * rows]
  float4* sorted_f4;    // [n * rows]: outward-rounded float corners of the sorted boxes (disjointness prefilter)
  float4* sorted_f4s;   // [n * rows]: shrunk float corners (see surely_below); == sorted_f4 when shrinking is off
  float4* sorted_kx;    // [n * rows]: (o.x, o.z, s.x, s.z) } the values a later box looks up in the kept block's sorted
  float4* sorted_ky;    // [n * rows]: (o.y, o.w, s.y, s.w) } key tables (KeptBlock), conditions 0-3 and 4-7
  float4* sorted_f4i;   // [n * rows]: inward-rounded float corners          } float interval arithmetic,
  float2* sorted_area;  // [n * rows]: area rounded down / up to float       } see decide_f32
  int* sorted_cls;      // [n * rows]
  unsigned long long* sorted_qidx;   // [n * rows]: the eight bucket indices (one byte each) of a box's lookup values in the
                                     // image-wide direct-address tables of the pipelined sweep
  unsigned char* flags; // [n * rows] bit0: removed, bit1: kept, bit2: non-finite or irregular box (exact path only)
  // outputs
  int* order;           // [n * rows]: kept rows in kept order
  int* n_keep;          // [n]
  int* n_cand;          // [n]
};

constexpr int NMS_THREADS = 1024;
constexpr int NMS_SMEM_KEYS = 16384;   // 128 KB of 64-bit keys
constexpr int NMS_BLOCK = 64;
constexpr int NMS_QB = 128;            // buckets per condition of the direct-address prefilter tables

__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <typename KeyPtr>
__device__ __forceinline__ void bitonic_sort_desc(KeyPtr keys, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// #{k : f(sk[k]) < y} with f(k) = (k - lo) * sc, the bucket function of the direct-address tables
__device__ __forceinline__ int nms_rank_f(const float* sk, float lo, float sc, float y) {
  int r = 0;
#pragma unroll
  for (int st = NMS_BLOCK / 2; st >= 1; st >>= 1) r += (__fmul_rn(__fsub_rn(sk[r + st - 1], lo), sc) < y) ? st : 0;
  return r + ((__fmul_rn(__fsub_rn(sk[r], lo), sc) < y) ? 1 : 0);
}

// Generic sweep: the kept boxes of one block of NMS_BLOCK sorted boxes, compacted.  A non-finite or irregular kept box
// carries all-covering prefilter corners, so the prefilter never drops it.
template <typename T>
struct KeptBlock {
  float4 f4[NMS_BLOCK];      // outward-rounded corners
  float4 f4s[NMS_BLOCK];     // shrunk corners
  BoxC<T> box[NMS_BLOCK];
  int cls[NMS_BLOCK];
  int nkept;
  unsigned char nf[NMS_BLOCK];
};

// Pipelined sweep: one block of NMS_BLOCK sorted boxes (not compacted: a kept box is named by its index in the block).
//
// Prefilter tables.  surely_below(q, p) is the negation of eight conditions "key_c(q) > v_c(p)" / "key_c(q) < v_c(p)":
//   c: 0 qs.z > po.x   1 qs.x < po.z   2 qo.z > ps.x   3 qo.x < ps.z   4 qs.w > po.y   5 qs.y < po.w   6 qo.w > ps.y   7 qo.y < ps.w
// For every condition the block's keys are sorted (skey[c] ascending, +inf padding) and tmask[c][r] is the set of boxes
// of the block that satisfy the condition when r keys lie below the looked-up value (r = #{key <= v} for the ">"
// conditions, r = #{key < v} for the "<" ones).  A later box finds r by binary search, seven steps, and the AND of its
// eight masks and the kept set is exactly the set of kept boxes that surely_below does not exclude: O(log) per
// (box, block) instead of 64 pair tests.  The same lookup yields the block's own pair mask.  Everything but `kept`
// depends on the geometry of the block alone and is prepared two blocks ahead, off the critical path.
//
// The sweep itself uses a direct-address form of the same tables: per condition the finite keys span [qlo, qhi], cut
// into NMS_QB uniform buckets; qmask[c][b] is the table mask that holds for EVERY value of bucket b (the mask at the
// bucket's lower edge for ">", at its upper edge for "<", edges widened by 1e-3 of a bucket against the float rounding of
// the bucket index; the first / last bucket also take everything outside the span).  One multiply, one conversion and
// one load per condition; the price is a prefilter that is up to one bucket looser per condition.
template <typename T>
struct SweepBlock {
  unsigned long long qmask[8][NMS_QB];
  float4 qlo[2], qscale[2];                 // per condition: bucket index = floor((v - qlo) * qscale), clamped
  unsigned long long tmask[8][NMS_BLOCK + 1];
  unsigned long long pairmask[NMS_BLOCK];   // bit j of [i]: i < j and box i suppresses box j
  unsigned long long kept;                  // the only field that depends on the earlier blocks
  float skey[8][NMS_BLOCK];
  float4 f4[NMS_BLOCK];      // outward-rounded corners   } all-covering for an exact-only box
  float4 f4s[NMS_BLOCK];     // shrunk corners            }
  float4 f4i[NMS_BLOCK];     // inward-rounded corners
  float4 vx[NMS_BLOCK];      // the boxes' own lookup values (sorted_kx / sorted_ky), for the block's turn as later boxes
  float4 vy[NMS_BLOCK];
  BoxC<T> box[NMS_BLOCK];
  float2 area[NMS_BLOCK];
  int cls[NMS_BLOCK];
  int nb;
  unsigned char nf[NMS_BLOCK];
};

constexpr int NMS_RESOLVER_THREADS = 64;                       // warps 0-1 finalize the next block, warps 2-5 prepare the one after
constexpr int NMS_Q2_ENTRIES = 64;                             // per warp: 32-bit (box, kept box) entries, decided 32 at a time
constexpr int NMS_DYN_SMEM = NMS_SMEM_KEYS * 8;                // sort keys; afterwards the per-warp queues, sort scratch and
                                                               // three SweepBlocks
static_assert((NMS_THREADS / 32) * NMS_Q2_ENTRIES * 4 + 8 * NMS_BLOCK * 8 + 3 * sizeof(SweepBlock<double>) <= NMS_DYN_SMEM, "NMS smem");
__device__ __forceinline__ void nms_resolver_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NMS_RESOLVER_THREADS) : "memory"); }
constexpr int NMS_PREPARER_THREADS = 128;                      // warps 2-5
__device__ __forceinline__ void nms_preparer_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(NMS_PREPARER_THREADS) : "memory"); }
__device__ __forceinline__ float unorderable(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// r = #{k : sk[k] <= v} (INCL) or #{k : sk[k] < v} over the 64 ascending keys sk
template <bool INCL>
__device__ __forceinline__ int nms_rank(const float* sk, float v) {
  int r = 0;
#pragma unroll
  for (int st = NMS_BLOCK / 2; st >= 1; st >>= 1) {
    const float k = sk[r + st - 1];
    r += (INCL ? (k <= v) : (k < v)) ? st : 0;
  }
  const float k = sk[r];
  return r + ((INCL ? (k <= v) : (k < v)) ? 1 : 0);
}

template <typename T, bool XY64, bool WH64>
__global__ void __launch_bounds__(NMS_THREADS, 1) sort_nms_kernel(const NmsArgs a) {
  extern __shared__ unsigned long long s_keys[];           // NMS_SMEM_KEYS
  __shared__ int s_count;
  __shared__ unsigned long long s_mask[NMS_BLOCK];
  __shared__ unsigned long long s_kept_mask;
  __shared__ unsigned s_alive[2];
  __shared__ KeptBlock<T> s_kb;                // generic sweep: the kept boxes of the current block
  __shared__ BoxC<T> s_box[NMS_BLOCK];
  __shared__ float4 s_f4[NMS_BLOCK];
  __shared__ float4 s_f4s[NMS_BLOCK];
  __shared__ int s_cls[NMS_BLOCK];
  __shared__ unsigned char s_nf[NMS_BLOCK];
  __shared__ int s_scan[32];
  __shared__ int s_next[2];                    // next group of 32 later boxes of the sweep (per kept-block buffer)
  __shared__ int s_running;
  // image-wide span of the finite prefilter keys per condition (orderable encoding), then the bucket function
  // b(v) = floor((v - s_qlo) * s_qsc) shared by every block's direct-address tables
  __shared__ unsigned s_glo[8], s_ghi[8];
  __shared__ float s_qlo[8], s_qsc[8];

  const int img = blockIdx.x;
  const int R = a.rows;
  const long long base = (long long)img * R;
  const float* prob = a.prob + base;
  const int tid = threadIdx.x;

  // ---- 1 + 2. one pass over the dense scores: count the candidates and build their keys
  // (orderable score << 32) | (0xFFFFFFFF - row); then sort descending ----
  unsigned long long* gkeys = a.keys + (long long)img * a.rows_pow2;     // only dereferenced beyond NMS_SMEM_KEYS candidates
  if (tid == 0) s_count = 0;
  if (tid < 8) { s_glo[tid] = 0xFFFFFFFFu; s_ghi[tid] = 0u; }
  __syncthreads();
  for (int r0 = 0; r0 < R; r0 += 4 * NMS_THREADS) {
    float pr4[4];                                  // four independent loads in flight before the (ordered) atomics
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * NMS_THREADS + tid;
      pr4[u] = (r < R) ? prob[r] : __int_as_float(0x7fc00000);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // warp-aggregated compaction: one ballot, one shared-memory atomic per warp, the lanes take consecutive slots
      // (the order of the keys does not matter: they are sorted next)
      const bool is_cand = pr4[u] == pr4[u];
      const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
      if (ballot) {
        const int lane = tid & 31;
        int slot0 = 0;
        if (lane == __ffs(ballot) - 1) slot0 = atomicAdd(&s_count, __popc(ballot));
        slot0 = __shfl_sync(0xffffffffu, slot0, __ffs(ballot) - 1);
        if (is_cand) {
          const int r = r0 + u * NMS_THREADS + tid;
          const int pos = slot0 + __popc(ballot & ((1u << lane) - 1u));
          const unsigned long long key = ((unsigned long long)orderable(pr4[u]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)r);
          if (pos < NMS_SMEM_KEYS) s_keys[pos] = key; else gkeys[pos] = key;
        }
      }
    }
  }
  __syncthreads();
  const int K = s_count;
  if (tid == 0) { a.n_cand[img] = K; }
  if (K == 0) {
    if (tid == 0) a.n_keep[img] = 0;
    return;
  }
  int P = 1;
  while (P < K) P <<= 1;
  const bool in_smem = (P <= NMS_SMEM_KEYS);
  if (!in_smem) {          // more candidates than the shared-memory sort holds: everything moves to the global scratch
    for (int i = tid; i < NMS_SMEM_KEYS; i += NMS_THREADS) gkeys[i] = s_keys[i];
  }
  for (int i = K + tid; i < P; i += NMS_THREADS) {
    if (in_smem) s_keys[i] = 0ull; else gkeys[i] = 0ull;
  }
  __syncthreads();
  if (in_smem) bitonic_sort_desc(s_keys, P); else bitonic_sort_desc(gkeys, P);

  const T thr = (T)a.thr;
  const bool per_class = a.per_class != 0;
  // iou == 0 never suppresses when thr > 0 (>=) / thr >= 0 (>): then surely disjoint pairs are skipped outright
  const bool skip_disjoint = per_class ? (thr >= (T)0) : (thr > (T)0);
  // float interval fast path (float64 boxes, positive threshold): bounds of thr (1 -+ 1e-6) rounded outwards
  const bool use_f32 = sizeof(T) == 8 && a.thr > 1e-30 && a.thr < 1e30;
  const float thr_dn = __double2float_rd(a.thr * 0.999999), thr_up = __double2float_ru(a.thr * 1.000001);
  // shrink factor of the prefilter boxes (surely_below): float64 boxes and thr >= 1e-2 only, else plain disjointness
  const double t_shrink = (sizeof(T) == 8 && a.thr >= 1e-2 && a.thr < 1e30) ? fmin(a.thr, 1.0) * 0.999999 : 0.0;

  // ---- 3. gather sorted boxes ----
  BoxC<T>* sb = reinterpret_cast<BoxC<T>*>(a.sorted_boxes) + base;
  float4* sf4 = a.sorted_f4 + base;
  float4* sf4s = a.sorted_f4s + base;
  float4* sky = a.sorted_ky + base;
  float4* skx = a.sorted_kx + base;
  float4* sf4i = a.sorted_f4i + base;
  float2* sarea = a.sorted_area + base;
  int* scls = a.sorted_cls + base;
  unsigned char* flags = a.flags + base;
  int* order = a.order + base;
  unsigned long long* sqidx = a.sorted_qidx + base;
  const bool tables = use_f32 && K > NMS_BLOCK;          // the pipelined sweep below (the only user of the bucket tables)
  float span_lo[8], span_hi[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { span_lo[c] = INFINITY; span_hi[c] = -INFINITY; }
  for (int i = tid; i < K; i += NMS_THREADS) {
    const unsigned long long key = in_smem ? s_keys[i] : gkeys[i];
    const int r = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    const T bx = XY64 ? (T) reinterpret_cast<const double*>(a.x)[base + r] : (T) reinterpret_cast<const float*>(a.x)[base + r];
    const T by = XY64 ? (T) reinterpret_cast<const double*>(a.y)[base + r] : (T) reinterpret_cast<const float*>(a.y)[base + r];
    const T bw = WH64 ? (T) reinterpret_cast<const double*>(a.w)[base + r] : (T) reinterpret_cast<const float*>(a.w)[base + r];
    const T bh = WH64 ? (T) reinterpret_cast<const double*>(a.h)[base + r] : (T) reinterpret_cast<const float*>(a.h)[base + r];
    const BoxC<T> b = make_box<T>(bx, by, bw, bh);
    sb[i] = b;
    const BoxF bf = make_boxf((double)b.x1, (double)b.y1, (double)b.x2, (double)b.y2, (double)b.area);
    sf4[i] = bf.out; sf4i[i] = bf.in; sarea[i] = bf.area;
    scls[i] = a.cls ? a.cls[base + r] : 0;
    bool exact_only = !box_finite(b);
    float4 shr = bf.out;
    if (t_shrink > 0.0 && !exact_only) {
      exact_only = !box_regular((double)b.x1, (double)b.y1, (double)b.x2, (double)b.y2, (double)b.area);
      shr = shrunk_f4((double)b.x1, (double)b.y1, (double)b.x2, (double)b.y2, t_shrink);
    }
    sf4s[i] = shr;
    sky[i] = make_float4(bf.out.y, bf.out.w, shr.y, shr.w);
    skx[i] = make_float4(bf.out.x, bf.out.z, shr.x, shr.z);
    flags[i] = exact_only ? 4 : 0;
    order[i] = r;          // provisional: sorted row ids; compacted to kept rows in step 5
    if (tables && !exact_only) {
      // the keys this box contributes to the sorted tables of its block (prepare_block), in condition order
      const float kc8[8] = {shr.z, shr.x, bf.out.z, bf.out.x, shr.w, shr.y, bf.out.w, bf.out.y};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (kc8[c] > -INFINITY && kc8[c] < INFINITY) { span_lo[c] = fminf(span_lo[c], kc8[c]); span_hi[c] = fmaxf(span_hi[c], kc8[c]); }
      }
    }
  }
  if (tables) {
    // image-wide bucket function: one span per condition over all finite keys; every box's eight bucket indices are
    // computed once here instead of once per (box, block) in the sweep
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float lo = span_lo[c], hi = span_hi[c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
      }
      if ((tid & 31) == 0 && lo <= hi) { atomicMin(&s_glo[c], orderable(lo)); atomicMax(&s_ghi[c], orderable(hi)); }
    }
    __syncthreads();
    if (tid < 8) {
      float lo = 0.f, sc = 0.f;
      if (s_glo[tid] <= s_ghi[tid]) {
        lo = unorderable(s_glo[tid]);
        const float span = __fsub_rn(unorderable(s_ghi[tid]), lo);
        sc = (span > 0.f && span < INFINITY) ? __fdiv_rn((float)NMS_QB, span) : 0.f;
        if (!(sc < INFINITY)) sc = 0.f;
      }
      s_qlo[tid] = lo; s_qsc[tid] = sc;
    }
    __syncthreads();
    for (int i = tid; i < K; i += NMS_THREADS) {
      const float4 vx = skx[i], vy = sky[i];
      const float v8[8] = {vx.x, vx.y, vx.z, vx.w, vy.x, vy.y, vy.z, vy.w};
      unsigned long long packed = 0ull;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const unsigned b = min(__float2uint_rd(__fmul_rn(__fsub_rn(v8[c], s_qlo[c]), s_qsc[c])), (unsigned)(NMS_QB - 1));
        packed |= (unsigned long long)b << (8 * c);
      }
      sqidx[i] = packed;
    }
  }
  __syncthreads();

  // ---- 4. blocked greedy sweep ----
  // Blocks of NMS_BLOCK sorted boxes.  Resolving a block = (a) its 64 x 64 pair mask, (b) the sequential greedy pass over
  // that mask by one thread, (c) a compacted copy of its kept boxes (KeptBlock).  Sweeping a block = every kept box of it
  // suppresses all later boxes.  Equivalent to the reference's loop: a box is kept iff no previously KEPT box reaches
  // iou >= thr with it.
  const int warp = tid >> 5, lane = tid & 31;
  const float4 all_covering = make_float4(-INFINITY, -INFINITY, INFINITY, INFINITY);

  // resolve the block starting at c0 (threads 0..63 only): first apply the kept boxes of the previous block (prev, may
  // be null) to its boxes, then (a)-(c) into out.
  auto resolve_block = [&](const int c0, const KeptBlock<T>* prev, KeptBlock<T>& out) {
    const int nb = min(NMS_BLOCK, K - c0);
    const int i = c0 + tid;
    const bool have = tid < nb;
    unsigned char f = have ? flags[i] : (unsigned char)1;
    BoxC<T> bx;
    bx.x1 = bx.y1 = bx.x2 = bx.y2 = bx.area = (T)0;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), sh = o;
    int c = 0;
    if (have) { bx = sb[i]; o = sf4[i]; sh = sf4s[i]; c = scls[i]; }
    if (prev != nullptr && have && !(f & 1)) {
      const int nk = prev->nkept;
      bool sup = false;
      for (int q = 0; q < nk && !sup; ++q) {
        if (per_class && prev->cls[q] != c) continue;
        if ((f & 4) | prev->nf[q]) {
          const T v = iou_ref<T, true>(prev->box[q], bx);
          sup = per_class ? (v > thr) : (v >= thr);
        } else {
          if (skip_disjoint && surely_below(prev->f4[q], prev->f4s[q], o, sh)) continue;
          sup = suppresses_finite<T>(prev->box[q], bx, thr, per_class);
        }
      }
      if (sup) { f |= 1; flags[i] = f; }
    }
    s_box[tid] = bx; s_f4[tid] = o; s_f4s[tid] = sh; s_cls[tid] = c; s_nf[tid] = f & 4; s_mask[tid] = 0ull;
    const unsigned bal = __ballot_sync(0xffffffffu, have && !(f & 1));
    if (lane == 0) s_alive[warp] = bal;
    nms_resolver_barrier();
    // (a) pair mask: bit j of s_mask[i] set iff i < j and box i suppresses box j.  Rows u and 62 - u hold 64 pairs
    // together, one per thread; row 31 is left to the upper half of the threads.
    for (int u = 0; u < 32; ++u) {
      int pi, pj;
      if (tid > u) { pi = u; pj = tid; } else { pi = 62 - u; pj = 63 - tid; }
      if ((u == 31 && tid <= u) || pj >= nb) continue;
      if (per_class && s_cls[pi] != s_cls[pj]) continue;
      bool sup;
      if (s_nf[pi] | s_nf[pj]) {
        const T v = iou_ref<T, true>(s_box[pi], s_box[pj]);
        sup = per_class ? (v > thr) : (v >= thr);
      } else {
        if (skip_disjoint && surely_below(s_f4[pi], s_f4s[pi], s_f4[pj], s_f4s[pj])) continue;
        sup = suppresses_finite<T>(s_box[pi], s_box[pj], thr, per_class);
      }
      if (sup) atomicOr(&s_mask[pi], 1ull << pj);
    }
    nms_resolver_barrier();
    // (b) sequential pass, branch-free, the masks fetched eight at a time ahead of the dependent chain
    if (tid == 0) {
      unsigned long long alive = (unsigned long long)s_alive[0] | ((unsigned long long)s_alive[1] << 32);
      unsigned long long kept = 0ull;
      for (int i0 = 0; i0 < NMS_BLOCK; i0 += 8) {
        unsigned long long m[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) m[u] = s_mask[i0 + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const unsigned long long bit = (alive >> (i0 + u)) & 1ull;
          kept |= bit << (i0 + u);
          alive &= ~(m[u] & (0ull - bit));
        }
      }
      s_kept_mask = kept;
    }
    nms_resolver_barrier();
    // (c) compacted copy of the kept boxes
    const unsigned long long kept = s_kept_mask;
    if (have && ((kept >> tid) & 1ull)) {
      flags[i] = f | 2;
      const int pos = __popcll(kept & ((1ull << tid) - 1ull));
      const bool nf = (f & 4) != 0;
      out.box[pos] = bx; out.cls[pos] = c; out.nf[pos] = nf ? 1 : 0;
      out.f4[pos] = nf ? all_covering : o; out.f4s[pos] = nf ? all_covering : sh;
    }
    if (tid == 0) out.nkept = __popcll(kept);
  };

  if (use_f32 && K > NMS_BLOCK) {
    // Pipelined variant (float64 boxes, positive threshold: the engine's regime).  While all warps sweep the boxes after
    // block b + 1 with the kept boxes of block b,
    //   warps 0-1 FINALIZE block b + 1: apply the kept boxes of block b to it, then the sequential greedy pass over its
    //             pair mask gives its kept set -- the only steps that depend on the earlier blocks;
    //   warps 2-5 PREPARE block b + 2: staging, sorted keys, tables, pair mask -- geometry only;
    // then both join the sweep, so the sequential chain from block to block is short and hidden.  Groups of 32 later
    // boxes are handed out through a shared counter.  A lane looks its box up in the eight key tables of the block
    // (SweepBlock); the (box, kept box) pairs the prefilter cannot exclude go to a per-warp queue that is carried across
    // groups and decided 32 entries at a time (float interval arithmetic, exact float64 when too close), so every lane
    // works in both phases.
    unsigned char* dyn = reinterpret_cast<unsigned char*>(s_keys);
    unsigned* wq2 = reinterpret_cast<unsigned*>(dyn) + warp * NMS_Q2_ENTRIES;
    unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(dyn + (NMS_THREADS / 32) * NMS_Q2_ENTRIES * 4);
    SweepBlock<T>* blks = reinterpret_cast<SweepBlock<T>*>(dyn + (NMS_THREADS / 32) * NMS_Q2_ENTRIES * 4 + 8 * NMS_BLOCK * 8);
    long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define NMS_TICK(k) { if (a.debug & 2) { const long long now_ = clock64(); tm[k] += now_ - t_last; t_last = now_; } }

    // the boxes of blk that surely_below does not exclude for the later box with lookup values vx, vy
    auto candidates = [&](const SweepBlock<T>& blk, const float4& vx, const float4& vy) -> unsigned long long {
      const int r0 = nms_rank<true>(blk.skey[0], vx.x), r1 = nms_rank<false>(blk.skey[1], vx.y);
      const int r2 = nms_rank<true>(blk.skey[2], vx.z), r3 = nms_rank<false>(blk.skey[3], vx.w);
      const int r4 = nms_rank<true>(blk.skey[4], vy.x), r5 = nms_rank<false>(blk.skey[5], vy.y);
      const int r6 = nms_rank<true>(blk.skey[6], vy.z), r7 = nms_rank<false>(blk.skey[7], vy.w);
      return blk.tmask[0][r0] & blk.tmask[1][r1] & blk.tmask[2][r2] & blk.tmask[3][r3] &
             blk.tmask[4][r4] & blk.tmask[5][r5] & blk.tmask[6][r6] & blk.tmask[7][r7];
    };
    // the direct-address form: a superset of candidates()
    auto bucket_candidates = [&](const SweepBlock<T>& blk, const float4& vx, const float4& vy) -> unsigned long long {
      const float4 la = blk.qlo[0], lb = blk.qlo[1], sa = blk.qscale[0], sb_ = blk.qscale[1];
#define NMS_QIDX(v, l, sc) min(__float2uint_rd(__fmul_rn(__fsub_rn(v, l), sc)), (unsigned)(NMS_QB - 1))
      const unsigned b0_ = NMS_QIDX(vx.x, la.x, sa.x), b1_ = NMS_QIDX(vx.y, la.y, sa.y), b2_ = NMS_QIDX(vx.z, la.z, sa.z), b3_ = NMS_QIDX(vx.w, la.w, sa.w);
      const unsigned b4_ = NMS_QIDX(vy.x, lb.x, sb_.x), b5_ = NMS_QIDX(vy.y, lb.y, sb_.y), b6_ = NMS_QIDX(vy.z, lb.z, sb_.z), b7_ = NMS_QIDX(vy.w, lb.w, sb_.w);
#undef NMS_QIDX
      return blk.qmask[0][b0_] & blk.qmask[1][b1_] & blk.qmask[2][b2_] & blk.qmask[3][b3_] &
             blk.qmask[4][b4_] & blk.qmask[5][b5_] & blk.qmask[6][b6_] & blk.qmask[7][b7_];
    };
    // the reference's decision for kept-role box q of blk against later-role box bj (either may be exact-only)
    auto pair_suppresses = [&](const SweepBlock<T>& blk, const int q, const BoxC<T>& bj, const bool exact_j) -> bool {
      if (exact_j | (blk.nf[q] != 0)) {
        const T v = iou_ref<T, true>(blk.box[q], bj);
        return per_class ? (v > thr) : (v >= thr);
      }
      return suppresses_finite<T>(blk.box[q], bj, thr, per_class);
    };

    // Direct-address tables of a prepared block, one entry per thread (all warps, at the start of the iteration before the
    // block is swept with).  The keys go through the same float function f(k) = (k - qlo) * qscale as the looked-up
    // values; f is monotone, so k > v implies f(k) >= f(v) >= bucket(v) and k < v implies f(k) <= f(v) < bucket(v) + 1:
    // qmask[c][b] = {f(key) >= b} for ">" and {f(key) < b + 1} for "<" can never miss a box, whatever the rounding.
    static_assert(NMS_THREADS == 8 * NMS_QB, "one (condition, bucket) entry per thread");
    auto fill_qmask = [&](SweepBlock<T>& blk) {
      const int cnd = tid / NMS_QB, b = tid % NMS_QB;
      const float lo = reinterpret_cast<const float*>(blk.qlo)[cnd], sc = reinterpret_cast<const float*>(blk.qscale)[cnd];
      const bool flat = sc == 0.f;                              // no span: every bucket takes all
      int r;
      if (cnd & 1) r = (b == NMS_QB - 1 || flat) ? NMS_BLOCK : nms_rank_f(blk.skey[cnd], lo, sc, (float)(b + 1));
      else r = (b == 0 || flat) ? 0 : nms_rank_f(blk.skey[cnd], lo, sc, (float)b);
      blk.qmask[cnd][b] = blk.tmask[cnd][r];
    };

    // PREPARE (threads 64..191; box t = thread 64 + t): everything about the block at c0 that its geometry alone determines
    auto prepare_block = [&](const int c0, SweepBlock<T>& blk) {
      const int t = tid - NMS_RESOLVER_THREADS, w = warp - NMS_RESOLVER_THREADS / 32;
      const int nb = min(NMS_BLOCK, K - c0);
      const int i = c0 + t;
      const bool have = t < nb;
      long long t_last = clock64();
      BoxC<T> bx;
      bx.x1 = bx.y1 = bx.x2 = bx.y2 = bx.area = (T)0;
      float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vy = vx;
      int c = 0;
      bool nf = false;
      if (t < NMS_BLOCK) {
        float4 o = make_float4(INFINITY, INFINITY, INFINITY, INFINITY), sh = o, fi = vx;
        float2 ar = make_float2(0.f, 0.f);
        if (have) {
          bx = sb[i]; c = scls[i]; fi = sf4i[i]; ar = sarea[i]; vx = skx[i]; vy = sky[i];
          nf = (flags[i] & 4) != 0;
          o = nf ? all_covering : sf4[i];
          sh = nf ? all_covering : sf4s[i];
        }
        blk.box[t] = bx; blk.cls[t] = c; blk.nf[t] = nf ? 1 : 0; blk.f4[t] = o; blk.f4s[t] = sh; blk.f4i[t] = fi;
        blk.area[t] = ar; blk.vx[t] = vx; blk.vy[t] = vy; blk.pairmask[t] = 0ull;
        if (t == 0) blk.nb = nb;
        // a padding entry (t >= nb) sorts last (+inf) and never enters a mask
        const float key[8] = {sh.z, have ? sh.x : INFINITY, o.z, have ? o.x : INFINITY, sh.w, have ? sh.y : INFINITY, o.w, have ? o.y : INFINITY};
#pragma unroll
        for (int cnd = 0; cnd < 8; ++cnd) s_sort[cnd * NMS_BLOCK + t] = ((unsigned long long)orderable(key[cnd]) << 32) | (unsigned)t;
      }
      nms_preparer_barrier();
      NMS_TICK(0)
      // sorted keys: each of the four warps sorts two of the eight key arrays (bitonic, 64 keys, two per lane), the two
      // independent sorts interleaved
      unsigned long long* sk = s_sort + w * 2 * NMS_BLOCK;
      for (int k = 2; k <= NMS_BLOCK; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          const int lo_i = ((lane & ~(j - 1)) << 1) | (lane & (j - 1)), hi_i = lo_i | j;
          const bool asc = (lo_i & k) == 0;
          unsigned long long va[2], vb[2];
#pragma unroll
          for (int m = 0; m < 2; ++m) { va[m] = sk[m * NMS_BLOCK + lo_i]; vb[m] = sk[m * NMS_BLOCK + hi_i]; }
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            if ((va[m] > vb[m]) == asc) { sk[m * NMS_BLOCK + lo_i] = vb[m]; sk[m * NMS_BLOCK + hi_i] = va[m]; }
          }
          __syncwarp();
        }
      }
      NMS_TICK(1)
      // tables: suffix OR of the boxes' bits along each sorted order; span of the finite keys (sorted: the -inf keys
      // come first, the +inf keys and the padding last)
      const unsigned long long all = nb < 64 ? (1ull << nb) - 1ull : ~0ull;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int cnd = w * 2 + m;
        const unsigned long long e0 = sk[m * NMS_BLOCK + lane], e1 = sk[m * NMS_BLOCK + lane + 32];
        const float k0 = unorderable((uint32_t)(e0 >> 32)), k1 = unorderable((uint32_t)(e1 >> 32));
        blk.skey[cnd][lane] = k0;
        blk.skey[cnd][lane + 32] = k1;
        unsigned long long s0 = all & (1ull << (int)(e0 & 63ull)), s1 = all & (1ull << (int)(e1 & 63ull));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned long long t0 = __shfl_down_sync(0xffffffffu, s0, d), t1 = __shfl_down_sync(0xffffffffu, s1, d);
          if (lane + d < 32) { s0 |= t0; s1 |= t1; }
        }
        s0 |= __shfl_sync(0xffffffffu, s1, 0);
        // ">" conditions (even): positions >= r satisfy; "<" conditions (odd): positions < r satisfy
        const bool less = (cnd & 1) != 0;
        blk.tmask[cnd][lane] = less ? (all & ~s0) : s0;
        blk.tmask[cnd][lane + 32] = less ? (all & ~s1) : s1;
        if (lane == 0) blk.tmask[cnd][NMS_BLOCK] = less ? all : 0ull;
        const int n_neg = __popc(__ballot_sync(0xffffffffu, k0 == -INFINITY)) + __popc(__ballot_sync(0xffffffffu, k1 == -INFINITY));
        const int n_pos = __popc(__ballot_sync(0xffffffffu, k0 == INFINITY)) + __popc(__ballot_sync(0xffffffffu, k1 == INFINITY));
        __syncwarp();
        (void)n_neg; (void)n_pos;
        if (lane == 0) {       // the image-wide bucket function (the boxes of a block are spread over the whole image anyway)
          reinterpret_cast<float*>(blk.qlo)[cnd] = s_qlo[cnd];
          reinterpret_cast<float*>(blk.qscale)[cnd] = s_qsc[cnd];
        }
      }
      nms_preparer_barrier();
      NMS_TICK(2)
      // pair mask: box t in the later role against the boxes before it
      if (have) {
        unsigned long long todo = (nf ? all : candidates(blk, vx, vy)) & ((1ull << t) - 1ull);
        while (todo) {
          const int q = __ffsll((long long)todo) - 1;
          todo &= todo - 1ull;
          if (per_class && blk.cls[q] != c) continue;
          if (pair_suppresses(blk, q, bx, nf)) atomicOr(&blk.pairmask[q], 1ull << t);
        }
      }
      NMS_TICK(3)
    };

    // FINALIZE (threads 0..63): apply the kept boxes of prev (may be null) to the prepared block at c0, then its kept set
    auto finalize_block = [&](const int c0, SweepBlock<T>& blk, const SweepBlock<T>* prev) {
      const int nb = min(NMS_BLOCK, K - c0);
      const int i = c0 + tid;
      const bool have = tid < nb;
      long long t_last = clock64();
      unsigned char f = have ? flags[i] : (unsigned char)1;
      if (prev != nullptr && have && !(f & 1) && prev->kept != 0ull) {
        const bool exact_j = (f & 4) != 0;
        unsigned long long todo = (exact_j ? ~0ull : candidates(*prev, blk.vx[tid], blk.vy[tid])) & prev->kept;
        bool sup = false;
        while (todo && !sup) {
          const int q = __ffsll((long long)todo) - 1;
          todo &= todo - 1ull;
          if (per_class && prev->cls[q] != blk.cls[tid]) continue;
          sup = pair_suppresses(*prev, q, blk.box[tid], exact_j);
        }
        if (sup) { f |= 1; flags[i] = f; }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, have && !(f & 1));
      if (lane == 0) s_alive[warp] = bal;
      nms_resolver_barrier();
      NMS_TICK(4)
      // sequential pass over the pair mask in 32-bit halves, the masks fetched eight at a time ahead of the dependent
      // chain; a box only ever removes later ones, so what is alive at the end is the kept set
      if (tid == 0) {
        unsigned alo = s_alive[0], ahi = s_alive[1];
        for (int i0 = 0; i0 < 32; i0 += 8) {
          unsigned long long m[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) m[u] = blk.pairmask[i0 + u];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (alo & (1u << (i0 + u))) { alo &= ~(unsigned)m[u]; ahi &= ~(unsigned)(m[u] >> 32); }
          }
        }
        for (int i0 = 32; i0 < 64; i0 += 8) {
          unsigned mh[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) mh[u] = (unsigned)(blk.pairmask[i0 + u] >> 32);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (ahi & (1u << (i0 + u - 32))) ahi &= ~mh[u];
          }
        }
        blk.kept = ((unsigned long long)ahi << 32) | alo;
      }
      nms_resolver_barrier();
      NMS_TICK(5)
      if (have && ((blk.kept >> tid) & 1ull)) flags[i] = f | 2;
    };

    const bool finalizer = tid < NMS_RESOLVER_THREADS, preparer = !finalizer && tid < NMS_RESOLVER_THREADS + NMS_PREPARER_THREADS;
    if (preparer) {
      prepare_block(0, blks[0]);
      if (NMS_BLOCK < K) { nms_preparer_barrier(); prepare_block(NMS_BLOCK, blks[1]); }
    }
    if (tid == 0) { s_next[0] = 0; s_next[1] = 0; }
    __syncthreads();
    if (finalizer) finalize_block(0, blks[0], nullptr);
    fill_qmask(blks[0]);
    __syncthreads();
    int bi = 0;                                   // block index modulo 3
    for (int b0 = 0, par = 0; b0 < K; b0 += NMS_BLOCK, par ^= 1, bi = (bi == 2) ? 0 : bi + 1) {
      const SweepBlock<T>& kc = blks[bi];
      const int b1 = (bi == 2) ? 0 : bi + 1, b2 = (b1 == 2) ? 0 : b1 + 1;
      if (b0 + 2 * NMS_BLOCK < K) fill_qmask(blks[b1]);        // block b + 1 is swept with only if boxes follow it
      if (finalizer) {
        if (b0 + NMS_BLOCK < K) finalize_block(b0 + NMS_BLOCK, blks[b1], &kc);
        if (tid == 0) s_next[par ^ 1] = 0;
      } else if (preparer) {
        if (b0 + 2 * NMS_BLOCK < K) prepare_block(b0 + 2 * NMS_BLOCK, blks[b2]);
      }
      const unsigned long long kept_c = kc.kept;
      if (kept_c != 0ull && !(a.debug & 1)) {
        int n2 = 0;                              // entries in the queue (warp-uniform)
        auto decide_entries = [&](const int n) {
          if (lane < n) {
            const unsigned e = wq2[lane];
            const int jj = (int)(e >> 6), q = (int)(e & 63u);
            const unsigned char fjj = flags[jj];
            if (!(fjj & 1) && !(per_class && kc.cls[q] != scls[jj])) {
              bool hit;
              int d = 0;
              if (!kc.nf[q]) d = decide_f32(kc.f4[q], kc.f4i[q], kc.area[q], sf4[jj], sf4i[jj], sarea[jj], thr_dn, thr_up);
              hit = d > 0;
              if (d == 0) hit = pair_suppresses(kc, q, sb[jj], false);     // too close to the threshold, or an exact-only kept box
              if (hit) flags[jj] = fjj | 1;     // several lanes may store the same byte: same value
            }
          }
        };
        const int first = b0 + 2 * NMS_BLOCK;
        auto grab = [&]() -> int {
          int g = 0;
          if (lane == 0) g = atomicAdd(&s_next[par], 1);
          return __shfl_sync(0xffffffffu, g, 0);
        };
        // software pipeline: the next group is claimed and its flags / lookup values are loaded while this one is worked on
        int g = grab();
        int j0 = first + g * 32;
        unsigned char fj_n = (unsigned char)1;
        unsigned long long qi_n = 0ull;
        if (j0 + lane < K) { fj_n = flags[j0 + lane]; qi_n = sqidx[j0 + lane]; }
        while (j0 < K) {
          const int j = j0 + lane;
          const unsigned char fj = fj_n;
          const unsigned long long qi = qi_n;
          g = grab();
          j0 = first + g * 32;
          fj_n = (unsigned char)1;
          if (j0 + lane < K) { fj_n = flags[j0 + lane]; qi_n = sqidx[j0 + lane]; }
          const bool live = !(fj & 1);
          unsigned long long todo = 0ull;
          if (live && !(fj & 4)) {
            // eight direct-address lookups with the box's precomputed bucket indices: a superset of candidates()
            const unsigned qa = (unsigned)qi, qb = (unsigned)(qi >> 32);
            todo = kc.qmask[0][qa & 0xFFu] & kc.qmask[1][(qa >> 8) & 0xFFu] & kc.qmask[2][(qa >> 16) & 0xFFu] & kc.qmask[3][qa >> 24] &
                   kc.qmask[4][qb & 0xFFu] & kc.qmask[5][(qb >> 8) & 0xFFu] & kc.qmask[6][(qb >> 16) & 0xFFu] & kc.qmask[7][qb >> 24] & kept_c;
          } else if (live) {
            // an exact-only box j (rare): numpy-semantics IoU against every kept box, on this lane alone
            const BoxC<T> bj = sb[j];
            const int cj = per_class ? scls[j] : 0;
            unsigned long long km = kept_c;
            bool sup = false;
            while (km && !sup) {
              const int q = __ffsll((long long)km) - 1;
              km &= km - 1ull;
              if (per_class && kc.cls[q] != cj) continue;
              sup = pair_suppresses(kc, q, bj, true);
            }
            if (sup) flags[j] = fj | 1;
          }
          // surviving pairs -> queue, one per lane and round
          while (__any_sync(0xffffffffu, todo != 0ull)) {
            const bool has = todo != 0ull;
            const int q = has ? __ffsll((long long)todo) - 1 : 0;
            todo &= todo - 1ull;
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if (has) wq2[n2 + __popc(bal & ((1u << lane) - 1u))] = ((unsigned)j << 6) | (unsigned)q;
            n2 += __popc(bal);
            __syncwarp();
            if (n2 >= 32) {
              decide_entries(32);
              const unsigned rest = (32 + lane < n2) ? wq2[32 + lane] : 0u;
              __syncwarp();
              if (32 + lane < n2) wq2[lane] = rest;
              n2 -= 32;
              __syncwarp();
            }
          }
        }
        decide_entries(n2);
      }
      __syncthreads();
    }
    if ((a.debug & 2) && img == 0 && tid == NMS_RESOLVER_THREADS)
      printf("nms prepare (clk, image 0, K=%d): stage %lld  sorts %lld  tables %lld  pairmask %lld\n", K, tm[0], tm[1], tm[2], tm[3]);
    if ((a.debug & 2) && img == 0 && tid == 0)
      printf("nms finalize (clk, image 0, K=%d): apply-prev %lld  serial %lld\n", K, tm[4], tm[5]);
#undef NMS_TICK
  } else {
    // Generic variant (float32 boxes, a threshold that every pair has to be evaluated for, or at most one block of
    // candidates -- the usual case on real images, where the tables of the pipelined variant would never be looked up):
    // block after block.
    for (int b0 = 0; b0 < K; b0 += NMS_BLOCK) {
      if (tid < NMS_RESOLVER_THREADS) resolve_block(b0, nullptr, s_kb);
      __syncthreads();
      const KeptBlock<T>& kc = s_kb;
      const int nkept = kc.nkept;
      if (nkept > 0) {
        for (int j = b0 + NMS_BLOCK + tid; j < K; j += NMS_THREADS) {
          const unsigned char fj = flags[j];
          if (fj & 1) continue;
          const int cj = per_class ? scls[j] : 0;
          unsigned lo = 0xffffffffu, hi = 0xffffffffu;
          if (skip_disjoint && !(fj & 4)) {
            const float4 f4j = sf4[j];
            lo = hi = 0u;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const float4 fa = kc.f4[q], fb = kc.f4[q + 32];      // entries >= nkept are stale but masked off below
              lo |= surely_disjoint(fa, f4j) ? 0u : (1u << q);
              hi |= surely_disjoint(fb, f4j) ? 0u : (1u << q);
            }
          }
          unsigned long long todo = ((unsigned long long)hi << 32) | lo;
          if (nkept < 64) todo &= (1ull << nkept) - 1ull;
          if (todo == 0ull) continue;
          bool sup = false;
          const BoxC<T> bj = sb[j];
          while (todo && !sup) {
            const int q = __ffsll((long long)todo) - 1;
            todo &= todo - 1;
            if (per_class && kc.cls[q] != cj) continue;
            if ((fj & 4) | kc.nf[q]) {
              const T v = iou_ref<T, true>(kc.box[q], bj);
              sup = per_class ? (v > thr) : (v >= thr);
            } else {
              sup = suppresses_finite<T>(kc.box[q], bj, thr, per_class);
            }
          }
          if (sup) flags[j] = fj | 1;
        }
      }
      __syncthreads();
    }
  }

  // ---- 5. ordered compaction of kept rows ----
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int c0 = 0; c0 < K; c0 += NMS_THREADS) {
    const int i = c0 + tid;
    const int keep = (i < K && (flags[i] & 2)) ? 1 : 0;
    const int row = (i < K) ? order[i] : 0;
    // block-wide exclusive scan of keep
    int incl = keep;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if ((tid & 31) >= d) incl += t;
    }
    if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      int wsum = (tid < NMS_THREADS / 32) ? s_scan[tid] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wsum, d);
        if (tid >= d) wsum += t;
      }
      s_scan[tid] = wsum;  // inclusive over warps
    }
    __syncthreads();
    const int warp_off = (tid >> 5) ? s_scan[(tid >> 5) - 1] : 0;
    const int pos = s_running + warp_off + incl - keep;
    const int chunk_total = s_scan[31];       // inclusive sum over all warps (entries beyond the last warp repeat it)
    __syncthreads();        // everyone has read order[i] / s_running / s_scan
    if (keep) order[pos] = row;   // pos <= i, and all reads of this chunk are done
    if (tid == 0) s_running += chunk_total;
    __syncthreads();
  }
  if (tid == 0) a.n_keep[img] = s_running;
}

// kept rows -> yb_det records (kept order)
struct DetOut {
  double w, h;
  float x, y;
  float prob;
  int class_idx;
  int row;
  int pad_;
};
__global__ void gather_dets_kernel(int rows, int max_per_image, const int* order, const int* n_keep, const float* prob,
                                   const float* x, const float* y, const double* w, const double* h, const int* cls,
                                   DetOut* out) {
  const int img = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nk = min(n_keep[img], max_per_image);
  if (i >= nk) return;
  const long long base = (long long)img * rows;
  const int r = order[base + i];
  DetOut d;
  d.w = w[base + r]; d.h = h[base + r]; d.x = x[base + r]; d.y = y[base + r];
  d.prob = prob[base + r]; d.class_idx = cls[base + r]; d.row = r; d.pad_ = 0;
  out[(long long)img * max_per_image + i] = d;
}

}  // namespace yb
