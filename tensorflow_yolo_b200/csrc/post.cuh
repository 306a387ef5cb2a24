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

struct NmsArgs {
  int rows;             // R: dense rows per image (stride of every per-row array)
  int per_class;        // 0: reference (class-agnostic, >=); 1: per class, strict >
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
  int* sorted_cls;      // [n * rows]
  unsigned char* flags; // [n * rows] bit0: removed, bit1: kept, bit2: non-finite box
  // outputs
  int* order;           // [n * rows]: kept rows in kept order
  int* n_keep;          // [n]
  int* n_cand;          // [n]
};

constexpr int NMS_THREADS = 1024;
constexpr int NMS_SMEM_KEYS = 16384;   // 128 KB of 64-bit keys
constexpr int NMS_BLOCK = 64;

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

template <typename T, bool XY64, bool WH64>
__global__ void __launch_bounds__(NMS_THREADS, 1) sort_nms_kernel(const NmsArgs a) {
  extern __shared__ unsigned long long s_keys[];           // NMS_SMEM_KEYS
  __shared__ int s_count;
  __shared__ unsigned long long s_mask[NMS_BLOCK];
  __shared__ unsigned long long s_kept_mask;
  __shared__ BoxC<T> s_box[NMS_BLOCK];
  __shared__ int s_cls[NMS_BLOCK];
  __shared__ unsigned char s_nf[NMS_BLOCK];
  __shared__ int s_scan[32];
  __shared__ int s_running;

  const int img = blockIdx.x;
  const int R = a.rows;
  const long long base = (long long)img * R;
  const float* prob = a.prob + base;
  const int tid = threadIdx.x;

  // ---- 1. count candidates ----
  if (tid == 0) s_count = 0;
  __syncthreads();
  int local = 0;
  for (int r = tid; r < R; r += NMS_THREADS) { const float pr = prob[r]; local += (pr == pr) ? 1 : 0; }
  if (local) atomicAdd(&s_count, local);
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
  unsigned long long* gkeys = a.keys + (long long)img * a.rows_pow2;
  __syncthreads();
  if (tid == 0) s_count = 0;
  __syncthreads();
  // ---- 2. build keys: (orderable score << 32) | (0xFFFFFFFF - row); sort descending ----
  for (int r0 = 0; r0 < R; r0 += NMS_THREADS) {
    const int r = r0 + tid;
    const float pr_ = (r < R) ? prob[r] : __int_as_float(0x7fc00000);
    const bool v = (pr_ == pr_);
    if (v) {
      const int pos = atomicAdd(&s_count, 1);
      const unsigned long long key = ((unsigned long long)orderable(prob[r]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)r);
      if (in_smem) s_keys[pos] = key; else gkeys[pos] = key;
    }
  }
  for (int i = K + tid; i < P; i += NMS_THREADS) {
    if (in_smem) s_keys[i] = 0ull; else gkeys[i] = 0ull;
  }
  __syncthreads();
  if (in_smem) bitonic_sort_desc(s_keys, P); else bitonic_sort_desc(gkeys, P);

  // ---- 3. gather sorted boxes ----
  BoxC<T>* sb = reinterpret_cast<BoxC<T>*>(a.sorted_boxes) + base;
  int* scls = a.sorted_cls + base;
  unsigned char* flags = a.flags + base;
  int* order = a.order + base;
  for (int i = tid; i < K; i += NMS_THREADS) {
    const unsigned long long key = in_smem ? s_keys[i] : gkeys[i];
    const int r = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    const T bx = XY64 ? (T) reinterpret_cast<const double*>(a.x)[base + r] : (T) reinterpret_cast<const float*>(a.x)[base + r];
    const T by = XY64 ? (T) reinterpret_cast<const double*>(a.y)[base + r] : (T) reinterpret_cast<const float*>(a.y)[base + r];
    const T bw = WH64 ? (T) reinterpret_cast<const double*>(a.w)[base + r] : (T) reinterpret_cast<const float*>(a.w)[base + r];
    const T bh = WH64 ? (T) reinterpret_cast<const double*>(a.h)[base + r] : (T) reinterpret_cast<const float*>(a.h)[base + r];
    const BoxC<T> b = make_box<T>(bx, by, bw, bh);
    sb[i] = b;
    scls[i] = a.cls ? a.cls[base + r] : 0;
    flags[i] = box_finite(b) ? 0 : 4;
    order[i] = r;          // provisional: sorted row ids; compacted to kept rows in step 5
  }
  __syncthreads();

  // ---- 4. blocked greedy sweep ----
  const T thr = (T)a.thr;
  const bool per_class = a.per_class != 0;
  for (int b0 = 0; b0 < K; b0 += NMS_BLOCK) {
    const int nb = min(NMS_BLOCK, K - b0);
    if (tid < NMS_BLOCK) {
      s_mask[tid] = 0ull;
      if (tid < nb) { s_box[tid] = sb[b0 + tid]; s_cls[tid] = scls[b0 + tid]; s_nf[tid] = flags[b0 + tid] & 4; }
    }
    __syncthreads();
    // 4a. intra-block pair mask: bit j of s_mask[i] set iff i<j and box i suppresses box j
    for (int pidx = tid; pidx < NMS_BLOCK * NMS_BLOCK; pidx += NMS_THREADS) {
      const int i = pidx >> 6, j = pidx & 63;
      if (i < j && j < nb) {
        bool sup;
        if (s_nf[i] | s_nf[j] | (sizeof(T) == 4)) {
          const T v = iou_ref<T, true>(s_box[i], s_box[j]);
          sup = per_class ? (v > thr) : (v >= thr);
        } else {
          const T v = iou_ref<T, false>(s_box[i], s_box[j]);
          sup = per_class ? (v > thr) : (v >= thr);
        }
        if (per_class && s_cls[i] != s_cls[j]) sup = false;
        if (sup) atomicOr(&s_mask[i], 1ull << j);
      }
    }
    __syncthreads();
    // 4b. sequential resolution of the block by one thread
    if (tid == 0) {
      unsigned long long alive = 0ull;
      for (int i = 0; i < nb; ++i)
        if (!(flags[b0 + i] & 1)) alive |= (1ull << i);
      unsigned long long kept = 0ull;
      for (int i = 0; i < nb; ++i) {
        if (alive & (1ull << i)) {
          kept |= (1ull << i);
          alive &= ~s_mask[i];
        }
      }
      s_kept_mask = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept_mask;
    if (tid < nb && (kept >> tid) & 1ull) flags[b0 + tid] |= 2;
    // 4c. every kept box of this block suppresses all later boxes
    if (kept != 0ull) {
      for (int j = b0 + NMS_BLOCK + tid; j < K; j += NMS_THREADS) {
        const unsigned char fj = flags[j];
        if (fj & 1) continue;
        const BoxC<T> bj = sb[j];
        const int cj = per_class ? scls[j] : 0;
        unsigned long long m = kept;
        bool sup = false;
        while (m && !sup) {
          const int i = __ffsll((long long)m) - 1;
          m &= m - 1;
          if (per_class && s_cls[i] != cj) continue;
          if ((fj & 4) | s_nf[i] | (sizeof(T) == 4)) {
            const T v = iou_ref<T, true>(s_box[i], bj);
            sup = per_class ? (v > thr) : (v >= thr);
          } else {
            const T v = iou_ref<T, false>(s_box[i], bj);
            sup = per_class ? (v > thr) : (v >= thr);
          }
        }
        if (sup) flags[j] = fj | 1;
      }
    }
    __syncthreads();
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
      int wsum = s_scan[tid];
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
    const int chunk_total = s_scan[31];
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
