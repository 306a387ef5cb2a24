/*
 * yolo_b200.h -- C ABI of libyolo_b200.so: the B200-native (sm_100a) YOLO TEST hot path.
 *
 * The reference (wns349/tensorflow-yolo) is pure Python over TensorFlow 1.x and has no FFI of its
 * own; its boundary for this path is the Python protocol of net/yolo.py:41-96.  Each entry point
 * below names the reference call it replaces.  A Python shim with the reference's signatures
 * (tensorflow_yolo_b200/net/*.py) binds these symbols through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success or a negative yb_status; yb_last_error() returns a
 *     thread-local, human-readable message for the last failure.  Nothing aborts or throws.
 *   - the caller owns every host buffer it passes; an engine owns all of its device memory.
 *   - an engine is bound to one CUDA device and is not thread-safe; use one engine (and one host
 *     thread or process) per GPU.  There is no CPU fallback: without a usable CUDA device every
 *     compute entry point fails with YB_ERR_CUDA.
 *   - "host" pointers are ordinary (pageable or pinned) memory; "device" pointers are CUDA device
 *     memory on the engine's device (yb_mem says which).
 *   - STREAM ORDER.  Every engine (and every yb_post) runs on a stream of its own (cudaStreamNonBlocking): its work is
 *     NOT ordered against work the caller enqueues on other streams.  A YB_MEM_DEVICE input must be complete before
 *     the call that reads it and must not be overwritten while the engine may still read it.  Either synchronise the
 *     producing stream, or bracket the call: yb_engine_order_after(e, producer_stream) before yb_engine_forward, and
 *     yb_engine_order_before(e, stream) before `stream` rewrites the buffer (same for yb_post_*).  A YB_MEM_HOST
 *     input of yb_engine_forward is copied asynchronously: keep it alive and unmodified until the forward after
 *     next has been issued, or until yb_engine_sync (two staging slots are in flight at most).
 *   - thresholds are doubles, as the reference's Python floats are.  The score threshold is compared in float32
 *     (numpy >= 2 casts the Python float to the float32 score's dtype: NEP 50 weak scalars), the IoU threshold in
 *     float64 against the float64 IoU (net/base.py:203) -- or in float32 when every box attribute is float32 (yb_nms).
 *     A NaN score never becomes a candidate (the reference would keep the row and then sort NaN keys: undefined order).
 */
#ifndef YOLO_B200_H_
#define YOLO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB_ABI_VERSION 2
#define YB_MAX_SRC 4
#define YB_MAX_ANCHORS 16

typedef enum yb_status {
  YB_OK = 0,
  YB_ERR_INVALID = -1,      /* bad argument / unsupported plan */
  YB_ERR_CUDA = -2,         /* CUDA runtime or driver failure (message has the CUDA error string) */
  YB_ERR_SHORT_WEIGHTS = -3,/* weight stream shorter than the plan needs (reference: ValueError, net/base.py:38) */
  YB_ERR_CAPACITY = -4,     /* caller-provided output buffer too small */
  YB_ERR_STATE = -5         /* call order violated (e.g. detect before forward) */
} yb_status;

/* plan entry kinds: one per class of the reference's net/layers.py */
typedef enum yb_kind {
  YB_INPUT = 0,      /* layers.py:106-109 */
  YB_CONV = 1,       /* layers.py:17-67   conv2d_bn_act */
  YB_MAXPOOL = 2,    /* layers.py:70-81 */
  YB_ROUTE = 3,      /* layers.py:84-87 */
  YB_REORG = 4,      /* layers.py:90-97 */
  YB_SHORTCUT = 5,   /* layers.py:100-103 */
  YB_UPSAMPLE = 6,   /* layers.py:112-116 */
  YB_YOLO = 7,       /* layers.py:126-134 */
  YB_DETECTION = 8   /* layers.py:119-123 */
} yb_kind;

/* One entry of the reference's flat `layers` list (net/v3.py:8-94, net/v2.py:10-60). */
typedef struct yb_layer {
  int kind;                 /* yb_kind */
  int filters;              /* conv: output channels */
  int ksize;                /* conv / maxpool window */
  int stride;               /* conv / maxpool stride, reorg / upsample factor */
  int batch_norm;           /* conv: 1 = batch-norm (eps 1e-5), no bias; 0 = bias */
  int leaky;                /* conv: 1 = leaky-ReLU(0.1); 0 = linear */
  int n_src;
  int src[YB_MAX_SRC];      /* absolute indices into the plan */
  int n_anchors;            /* yolo */
  float anchors[2 * YB_MAX_ANCHORS]; /* yolo: (w,h) pairs in grid units (already divided by the stride) */
} yb_layer;

typedef enum yb_mem { YB_MEM_HOST = 0, YB_MEM_DEVICE = 1 } yb_mem;
typedef enum yb_dtype { YB_F32 = 0, YB_U8 = 1 } yb_dtype;

/* decode variants: net/v3.py:109-136 (sigmoid classes, score = objectness) and
 * net/v2.py:93-119 (softmax classes, score = objectness * max class probability) */
typedef enum yb_decode_mode { YB_DECODE_V3 = 0, YB_DECODE_V2 = 1 } yb_decode_mode;

/* NMS variants.  YB_NMS_REFERENCE = net/base.py:195-209: class-agnostic, suppress iff IoU >= thr.
 * YB_NMS_PER_CLASS = the optional north-star mode: per class, suppress iff IoU > thr. */
typedef enum yb_nms_mode { YB_NMS_REFERENCE = 0, YB_NMS_PER_CLASS = 1 } yb_nms_mode;

/* One detection == one reference BoundingBox (net/base.py:257-272); dtypes follow what the
 * reference's decode produces under numpy>=2: x,y,prob float32; w,h float64. */
typedef struct yb_det {
  double w, h;              /* normalised size */
  float x, y;               /* normalised centre */
  float prob;
  int32_t class_idx;
  int32_t row;              /* row of net[-1].out this box was decoded from */
  int32_t pad_;
} yb_det;

typedef struct yb_engine yb_engine;

/* ---- library ---- */
int yb_abi_version(void);
const char* yb_last_error(void);
/* number of visible CUDA devices (0 and YB_OK when there is no GPU / no driver) */
int yb_device_count(int* count);

/* CRC-32C (Castagnoli) of data[0..n) continuing from `seed` (0 = fresh), the checksum TensorFlow's tensor-bundle
 * checkpoints carry per tensor and per index block (tensorflow/core/lib/hash/crc32c.h); used by the checkpoint
 * reader that replaces tf.train.Saver.restore (net/yolo.py:71-72, net/base.py:55-61).  Host only.
 * impl 0 = SSE4.2 when available, 1 = portable tables. */
int yb_crc32c(const void* data, size_t n, uint32_t seed, int impl, uint32_t* out);

/* ---- engine: replaces tf graph build + tf.Session (net/yolo.py:63,67-68) ---- */
/* plan: the flat layer list.  decode_mode/num_classes describe the head.  For YOLOv2 (whose
 * reference plan ends in the linear conv, net/v2.py:52-59) the caller appends one YB_YOLO entry
 * carrying the anchors so that yb_engine_detect knows the head geometry. */
int yb_engine_create(const yb_layer* plan, int n_layers, int in_h, int in_w, int in_c,
                     int max_batch, int device, int decode_mode, int num_classes, yb_engine** out);
void yb_engine_destroy(yb_engine* e);

/* Replaces base.load_weights + sess.run(ops) (net/base.py:26-46, net/yolo.py:74-75): consumes the
 * darknet float32 stream (header already stripped) in plan order -- per BN conv
 * beta,gamma,moving_mean,moving_variance,kernel[O][I][kh][kw]; per linear conv bias,kernel.
 * *consumed receives the number of floats read (the reference prints read/len, base.py:44);
 * a surplus is not an error, a shortfall is YB_ERR_SHORT_WEIGHTS. */
int yb_engine_load_weights(yb_engine* e, const float* stream, size_t n, size_t* consumed);

/* Replaces sess.run(net[-1].out, {net[0].out: x_batch}) (net/yolo.py:83).  images: NHWC,
 * n*in_h*in_w*in_c elements of dtype (float32 in [0,1], or uint8 which is scaled by 1/255),
 * in host or device memory.  Asynchronous with respect to the host on the engine's stream;
 * results stay on the device.  A device pointer should be 16-byte aligned (the first layers read the
 * images through a TMA tensor map); one that is not is first copied, device to device, into the
 * engine's staging buffer -- same results. */
int yb_engine_forward(yb_engine* e, const void* images, int dtype, int mem, int n);

/* Replaces base.generate_test_batch / preprocess_image (net/base.py:115-168) + the forward: takes the decoded 8-bit
 * BGR images as cv2.imread returns them (host memory, any sizes; strides[i] = bytes per row, NULL = widths[i]*3),
 * and does cv2.resize(image, (input_h, input_w)) [INTER_LINEAR, bit-exact fixed point], BGR->RGB and the /255 scaling
 * on the device before running the conv stack.  The raw images are uploaded on a copy stream into double-buffered
 * staging, so the upload of batch i+1 overlaps the compute of batch i.  Like the reference, only square network
 * inputs work (its dsize = (input_h, input_w) quirk makes its own placeholder reject anything else). */
int yb_engine_forward_raw(yb_engine* e, const void* const* images, const int* heights, const int* widths,
                          const int* strides, int n);
/* The preprocessed uint8 RGB batch [n, H, W, 3] of the last yb_engine_forward_raw (parity tests). */
int yb_engine_read_input_u8(yb_engine* e, unsigned char* host_out, size_t capacity);
/* Stand-alone cv2.resize(image, (dst_w, dst_h)) + BGR->RGB of n host images into host memory [n, dst_h, dst_w, 3]. */
int yb_resize_bgr2rgb(const void* const* images, const int* heights, const int* widths, const int* strides, int n,
                      int dst_h, int dst_w, unsigned char* dst_host, int device);

/* Copies the reference's net[-1].out for the last forward to host float32:
 * [n, R, 5+C] (v3, net/v3.py:90-93) or [n, h, w, A*(5+C)] (v2, net/v2.py:59).  For parity tests;
 * the detect path never materialises it.  capacity in floats. */
int yb_engine_read_output(yb_engine* e, float* host_out, size_t capacity);
int yb_engine_output_shape(yb_engine* e, int* rows, int* cols); /* rows = R (v3) or h*w (v2); cols = 5+C or A*(5+C) */

/* Debug/parity: copies the NHWC float32 value of plan entry `layer` (as the reference's
 * layers[layer].out) for the last forward.  shape_hwc receives h,w,c. */
int yb_engine_read_layer(yb_engine* e, int layer, float* host_out, size_t capacity, int shape_hwc[3]);

/* Replaces find_bounding_boxes (net/v3.py:139-151, net/v2.py:82-90) for the last forward:
 * decode + threshold + NMS on the device.  out: [n][max_per_image] detections in kept
 * (score-descending) order (row stride is always max_per_image); counts[n] receives the number of kept
 * boxes per image, which may exceed max_per_image -- then only the first max_per_image were written.
 * Synchronises the engine's stream. */
int yb_engine_detect(yb_engine* e, double threshold, double iou_threshold, int nms_mode,
                     yb_det* out, int* counts, int max_per_image);
/* Same, but leaves the results on the device (for device-timed benches); returns after enqueueing. */
int yb_engine_detect_async(yb_engine* e, double threshold, double iou_threshold, int nms_mode);
int yb_engine_sync(yb_engine* e);
/* Stream ordering against the caller's CUDA streams (cuda_stream: a cudaStream_t / CUstream handle, NULL = the legacy
 * default stream).  order_after: the engine's stream waits for everything enqueued on cuda_stream so far (call it
 * before yb_engine_forward when that stream produces a YB_MEM_DEVICE input).  order_before: cuda_stream waits until
 * the engine has consumed the input of its last forward (call it before that stream overwrites the input buffer). */
int yb_engine_order_after(yb_engine* e, void* cuda_stream);
int yb_engine_order_before(yb_engine* e, void* cuda_stream);

/* per-op timing of the last forward+detect, measured with CUDA events on the engine's stream:
 * fills up to cap entries of (plan layer index or -1, milliseconds); returns the number of ops in *n_ops */
int yb_engine_profile(yb_engine* e, const void* images, int dtype, int mem, int n,
                      int* layer_idx, float* ms, int cap, int* n_ops);
/* conv implementation: 0 = tcgen05/TMA implicit GEMM (default), 1 = plain CUDA-core kernel
 * (debug cross-check only; never selected implicitly). */
int yb_engine_set_conv_impl(yb_engine* e, int impl);
/* Measures, on the device, every launch configuration of each tcgen05 conv (N tile 32..256, cta_group::2 CTA pairs,
 * weight-stationary B, TMA-store epilogue) for a batch of n images and keeps the fastest per layer; layers with the
 * same shape share one measurement.  Results do not change (the K reduction order is the same in every
 * configuration).  Clobbers the activations: call yb_engine_forward again before reading outputs.  reps <= 0 = 5. */
int yb_engine_autotune(yb_engine* e, int n, int reps);
/* JSON written by the last yb_engine_autotune (every candidate with its time).  buf may be NULL to query *needed. */
int yb_engine_tune_report(yb_engine* e, char* buf, size_t capacity, size_t* needed);
/* Debug / measurement switches: "pdl" (programmatic dependent launch between convs, default 1), "pairs", "bstat",
 * "tma_epi" (allow those kernel features, default 1), "ablate" (bit mask of roofline probes of the conv kernel:
 * 1 = no epilogue work, 2 = no MMAs, 4 = no activation loads, 8 = no weight loads; outputs are wrong while set),
 * "graph" (CUDA graph of the forward: -1 = automatic, used for batches of at most 32 images, where the 75 launches
 * rather than the device work bound sess.run of net/yolo.py:83; 0 = never; 1 = always.  A (batch, input buffer, dtype)
 * combination runs eagerly the first time, is captured the second time and replayed from then on),
 * "split_k" (latency mode, default 0 or YB_SPLIT_K: convs whose grid fills a fraction of the SMs -- small batches, the
 * per-GPU shards of a strong-scaled batch, config/yolo_3.ini:36 batch_size = 1 -- are split along K over 2-4 work
 * units per tile, partial accumulators summed in a fixed order; the fp32 summation order, hence the last bits of the
 * output, then depend on the batch size, which is why it is opt-in). */
int yb_engine_set_option(yb_engine* e, const char* name, int value);
/* Forces the launch configuration of launched op `op_index` (bn = 0 restores the heuristic; bstat / tma_epi: -1 =
 * heuristic, 0 = off, 1 = on when possible; ksub = BK-blocks per pipeline stage, 0 = heuristic) and times one op in isolation (average of reps launches, milliseconds). */
int yb_engine_set_conv_cfg(yb_engine* e, int op_index, int bn, int pair, int bstat, int tma_epi, int ksub);
int yb_engine_time_op(yb_engine* e, int op_index, int n, int reps, float* ms);
/* With option "cycles" = 1 the conv kernel accumulates SM cycle counters per role, summed over CTAs and launches
 * (producer: wait-for-free-stage, total; MMA issuer: wait-for-data, wait-for-accumulator, total; epilogue warps:
 * wait-for-accumulator, wait-for-residual, wait-for-staging-buffer, TMEM load, math+store, total).  out16: 16 values. */
int yb_engine_read_cycles(yb_engine* e, unsigned long long* out16, int reset);
/* number of kernel launches the last forward / detect enqueued */
int yb_engine_launch_count(yb_engine* e, int* forward_launches, int* detect_launches);
/* Number of forwards launched by replaying a captured CUDA graph since the engine was created. */
int yb_engine_graph_replays(yb_engine* e, int* replays);

/* ---- pipelining, timing and introspection (used by bench.py; optional for integrators) ---- */
/* Records device event `idx` (0..7) on the engine's stream / returns the device time between two marks
 * (synchronises on `to`).  This is how the bench times K steps on the stream the kernels run on. */
int yb_engine_mark(yb_engine* e, int idx);
int yb_engine_elapsed(yb_engine* e, int from, int to, float* ms);
/* After yb_engine_detect_async: enqueue the copy of the results into caller (ideally pinned) host buffers
 * without waiting; yb_engine_fetch_wait(slot) blocks until that copy is complete.  Two slots allow the
 * caller to overlap step i+1 (including its H2D image copy, which runs on a separate copy stream) with
 * the consumption of step i. */
int yb_engine_fetch_async(yb_engine* e, yb_det* out, int* counts, int max_per_image, int slot);
int yb_engine_fetch_wait(yb_engine* e, int slot);
/* Per-op CUDA-event timing inside yb_engine_forward: enable, run forwards, then read the summed
 * milliseconds per launched op (layer_idx = plan entry that op implements) over n_forwards forwards. */
int yb_engine_profiling(yb_engine* e, int enable);
int yb_engine_profile_read(yb_engine* e, int* layer_idx, float* ms_sum, int cap, int* n_ops, int* n_forwards);
/* Describes launched op `op_index`: path 0 = tcgen05 conv, 1 = direct first conv, 2 = CUDA-core conv, 3 = fused pair
 * (this conv also computes the conv feeding it, whose own op reports 0 FLOPs and is never launched),
 * negative = non-conv kernel; tile shape; algorithmic FLOPs per image (2*MAC). */
int yb_engine_op_info(yb_engine* e, int op_index, int* layer, int* path, int* bn_tile, int* bk, int* stages,
                      double* flops_per_image);

/* Launch configuration of op `op_index` as resolved by its last launch: N tile, CTA pairs, weight-stationary B,
 * TMA-store epilogue, pipeline stages, BK-blocks per stage. */
int yb_engine_op_cfg(yb_engine* e, int op_index, int* bn, int* pair, int* bstat, int* tma_epi, int* stages, int* ksub);
/* Split-K factor the last launch of op `op_index` used (1 = none; option "split_k"). */
int yb_engine_op_splitk(yb_engine* e, int op_index, int* splitk);

/* ---- stand-alone post-processing on caller tensors ---- */
typedef struct yb_scale {
  int h, w, n_anchors;
  float anchors[2 * YB_MAX_ANCHORS];  /* grid units */
} yb_scale;

/* Opaque post-processing context for stand-alone decode+NMS over head tensors in the reference's
 * net[-1].out layout (BASELINE config 5).  rows_per_image = sum(h*w*A). */
typedef struct yb_post yb_post;
int yb_post_create(const yb_scale* scales, int n_scales, int num_classes, int decode_mode,
                   int max_batch, int device, yb_post** out);
void yb_post_destroy(yb_post* p);
/* head: [n, R, 5+C] float32 (v3) or [n, h, w, A*(5+C)] (v2), host or device memory.
 * Runs decode -> sort -> NMS on the device; out/counts as in yb_engine_detect (may be NULL to
 * leave results on the device).  cand_counts (optional, [n]) receives the number of candidates that
 * passed the threshold before NMS. */
int yb_post_run(yb_post* p, const float* head, int mem, int n, double threshold, double iou_threshold,
                int nms_mode, yb_det* out, int* counts, int max_per_image, int* cand_counts);
/* decode only: candidates of image i in undefined order (sort by row to compare), host output. */
int yb_post_decode(yb_post* p, const float* head, int mem, int n, double threshold,
                   yb_det* out, int* counts, int max_per_image);
int yb_post_sync(yb_post* p);
/* as yb_engine_order_after / yb_engine_order_before, for a YB_MEM_DEVICE head tensor (order_before: cuda_stream waits
 * for everything the context has enqueued so far) */
int yb_post_order_after(yb_post* p, void* cuda_stream);
int yb_post_order_before(yb_post* p, void* cuda_stream);
/* device time of the last yb_post_run split by stage, milliseconds (decode, sort+nms) */
int yb_post_last_ms(yb_post* p, float* decode_ms, float* nms_ms);

/* Replaces base.non_maximum_suppression (net/base.py:195-209) on caller boxes (host memory):
 * x,y,w,h,prob are [k] arrays; coordinates are float64 if f64 != 0 else float32 (the IoU of
 * net/base.py:180-192 is then evaluated in that precision, like numpy does); class_idx may be
 * NULL for YB_NMS_REFERENCE.  iou_threshold is rounded to float32 first when f64 == 0 (numpy's
 * weak-scalar rule).  keep receives the indices of the kept boxes in kept order, *n_keep their
 * number. */
int yb_nms(const void* x, const void* y, const void* w, const void* h, const float* prob,
           const int32_t* class_idx, int k, int f64, double iou_threshold,
           int nms_mode, int device, int32_t* keep, int* n_keep);

#ifdef __cplusplus
}
#endif
#endif /* YOLO_B200_H_ */
