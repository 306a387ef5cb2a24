"""Python handles over the C ABI: Engine (conv stack + detect), PostProcessor (decode+NMS on head
tensors) and nms() (the reference's non_maximum_suppression on caller boxes).  No arithmetic here."""
import ctypes

import numpy as np

from . import _lib
from . import plan as _plan
from ._lib import DET_DTYPE, YB_DECODE_V2, YB_DECODE_V3, YB_NMS_PER_CLASS, YB_NMS_REFERENCE  # noqa: F401


def _buffer(arr, np_dtypes):
    """(pointer, mem, dtype code, keepalive) for a numpy array or a torch tensor (host or CUDA)."""
    if isinstance(arr, np.ndarray):
        if arr.dtype not in np_dtypes:
            raise TypeError("expected dtype in {}, got {}".format(np_dtypes, arr.dtype))
        a = np.ascontiguousarray(arr)
        return a.ctypes.data, _lib.YB_MEM_HOST, a.dtype, a
    if hasattr(arr, "data_ptr") and hasattr(arr, "is_cuda"):      # torch.Tensor without importing torch
        if not arr.is_contiguous():
            raise ValueError("tensor must be contiguous")
        dt = np.dtype(str(arr.dtype).replace("torch.", ""))
        if dt not in np_dtypes:
            raise TypeError("expected dtype in {}, got {}".format(np_dtypes, dt))
        return arr.data_ptr(), (_lib.YB_MEM_DEVICE if arr.is_cuda else _lib.YB_MEM_HOST), dt, arr
    raise TypeError("expected a numpy array or a torch tensor")


def _producer_stream(arr):
    """CUDA stream handle (int) the caller's framework is currently enqueueing on for a device tensor, or None.
    Only torch is recognised (through the already-imported module: this package never imports it itself)."""
    if not (hasattr(arr, "is_cuda") and arr.is_cuda):
        return None
    import sys
    torch = sys.modules.get("torch")
    if torch is None:
        return None
    return int(torch.cuda.current_stream(arr.device).cuda_stream)


def _raw_image_args(images):
    keep = []
    n = len(images)
    ptrs, hs, ws, strides = (ctypes.c_void_p * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)()
    for i, im in enumerate(images):
        im = np.asarray(im)
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise TypeError("image {}: expected a uint8 [h, w, 3] array, got {} {}".format(i, im.dtype, im.shape))
        if im.strides[2] != 1 or im.strides[1] != 3 or im.strides[0] < im.shape[1] * 3:
            im = np.ascontiguousarray(im)
        keep.append(im)
        ptrs[i], hs[i], ws[i], strides[i] = im.ctypes.data, im.shape[0], im.shape[1], im.strides[0]
    return ptrs, hs, ws, strides, keep


def resize_bgr2rgb(images, dst_h, dst_w, device=0):
    """cv2.resize(image, (dst_w, dst_h)) [INTER_LINEAR] + BGR->RGB on the device: [n, dst_h, dst_w, 3] uint8."""
    ptrs, hs, ws, strides, keep = _raw_image_args(images)
    out = np.empty((len(images), dst_h, dst_w, 3), dtype=np.uint8)
    _lib.check(_lib.lib().yb_resize_bgr2rgb(ptrs, hs, ws, strides, len(images), dst_h, dst_w, out.ctypes.data, device))
    return out


class Engine(object):
    """One compiled network on one GPU (yb_engine)."""

    def __init__(self, plan, input_shape, num_classes, decode_mode, max_batch=1, device=0):
        self._h = None
        self.plan = list(plan)
        self.input_shape = tuple(int(s) for s in input_shape)
        self.num_classes, self.decode_mode, self.max_batch, self.device = num_classes, decode_mode, max_batch, device
        arr = _plan.to_c_array(self.plan)
        handle = ctypes.c_void_p()
        h, w, c = self.input_shape
        _lib.check(_lib.lib().yb_engine_create(arr, len(self.plan), h, w, c, max_batch, device, decode_mode,
                                               num_classes, ctypes.byref(handle)))
        self._h = handle
        rows, cols = ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().yb_engine_output_shape(self._h, ctypes.byref(rows), ctypes.byref(cols)))
        self.out_rows, self.out_cols = rows.value, cols.value
        self.last_n = 0

    def close(self):
        if self._h is not None:
            _lib.lib().yb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_weights(self, stream):
        """stream: float32 darknet payload (header stripped).  Returns the number of floats consumed."""
        stream = np.ascontiguousarray(stream, dtype=np.float32)
        consumed = ctypes.c_size_t(0)
        _lib.check(_lib.lib().yb_engine_load_weights(self._h, stream.ctypes.data, stream.size, ctypes.byref(consumed)))
        return consumed.value

    def forward(self, images):
        """images: [n,H,W,C] float32 in [0,1] or uint8; numpy (host) or torch (host/CUDA).  Asynchronous."""
        if isinstance(images, np.ndarray) and images.dtype == np.float64:
            images = images.astype(np.float32)       # the reference feeds float64 into a float32 placeholder
        ptr, mem, dt, keep = _buffer(images, (np.dtype(np.float32), np.dtype(np.uint8)))
        shape = tuple(images.shape)
        if len(shape) != 4 or shape[1:] != self.input_shape:
            raise ValueError("expected images of shape [n,{},{},{}], got {}".format(*(self.input_shape + (shape,))))
        # host inputs are copied asynchronously from the caller's memory: up to two H2D copies are in flight (two staging
        # slots), so the last three inputs are kept alive here; device inputs are ordered against the producing stream
        self._keep = (getattr(self, "_keep", ()) + (keep,))[-3:]
        ps = _producer_stream(images)
        if ps is not None:
            _lib.check(_lib.lib().yb_engine_order_after(self._h, ps))
        _lib.check(_lib.lib().yb_engine_forward(self._h, ptr, _lib.YB_F32 if dt == np.float32 else _lib.YB_U8, mem, shape[0]))
        if ps is not None:      # later writes to the tensor on that stream wait until the first conv has read it
            _lib.check(_lib.lib().yb_engine_order_before(self._h, ps))
        self.last_n = shape[0]

    def forward_raw(self, images):
        """images: list of uint8 BGR arrays [h, w, 3] as cv2.imread returns them (any sizes).  Resize (cv2 INTER_LINEAR,
        bit-exact), BGR->RGB and /255 run on the device, then the conv stack.  Asynchronous."""
        ptrs, hs, ws, strides, keep = _raw_image_args(images)
        self._keep = (getattr(self, "_keep", ()) + (keep,))[-3:]
        _lib.check(_lib.lib().yb_engine_forward_raw(self._h, ptrs, hs, ws, strides, len(images)))
        self.last_n = len(images)

    def read_input_u8(self):
        """The preprocessed uint8 RGB batch [n, H, W, 3] of the last forward_raw."""
        out = np.empty((self.last_n,) + self.input_shape, dtype=np.uint8)
        _lib.check(_lib.lib().yb_engine_read_input_u8(self._h, out.ctypes.data, out.size))
        return out

    def read_output(self):
        """The reference's net[-1].out for the last forward, as float32 numpy."""
        n = self.last_n
        out = np.empty((n, self.out_rows, self.out_cols), dtype=np.float32)
        _lib.check(_lib.lib().yb_engine_read_output(self._h, out.ctypes.data, out.size))
        return out

    def read_layer(self, index):
        """NHWC float32 value of plan entry `index` (needs YB_KEEP_ALL=1 for recycled intermediates)."""
        spec = self.plan[index]
        n = self.last_n
        cap = n * int(np.prod(spec.shape)) if spec.kind not in (_plan.KIND_YOLO, _plan.KIND_DETECTION) else 0
        if cap == 0:
            raise ValueError("layer {} has no NHWC tensor".format(index))
        out = np.empty(cap, dtype=np.float32)
        hwc = (ctypes.c_int * 3)()
        _lib.check(_lib.lib().yb_engine_read_layer(self._h, index, out.ctypes.data, out.size, ctypes.byref(hwc)))
        return out.reshape(n, hwc[0], hwc[1], hwc[2])

    def detect(self, threshold, iou_threshold, nms_mode=YB_NMS_REFERENCE, max_per_image=None):
        """decode + NMS of the last forward.  Returns a list (per image) of structured arrays (DET_DTYPE)
        in kept order."""
        n = self.last_n
        cap = int(max_per_image or 256)
        while True:
            dets = np.zeros((n, cap), dtype=DET_DTYPE)
            counts = np.zeros(n, dtype=np.int32)
            _lib.check(_lib.lib().yb_engine_detect(self._h, threshold, iou_threshold, nms_mode,
                                                   dets.ctypes.data, counts.ctypes.data, cap))
            if max_per_image is not None or counts.max(initial=0) <= cap:
                break
            cap = int(counts.max())          # rare: more detections than the first guess; fetch again
        return [dets[i, :min(int(counts[i]), cap)] for i in range(n)]

    def detect_async(self, threshold, iou_threshold, nms_mode=YB_NMS_REFERENCE):
        _lib.check(_lib.lib().yb_engine_detect_async(self._h, threshold, iou_threshold, nms_mode))

    def sync(self):
        _lib.check(_lib.lib().yb_engine_sync(self._h))

    def mark(self, idx):
        _lib.check(_lib.lib().yb_engine_mark(self._h, idx))

    def elapsed_ms(self, a, b):
        ms = ctypes.c_float()
        _lib.check(_lib.lib().yb_engine_elapsed(self._h, a, b, ctypes.byref(ms)))
        return ms.value

    def fetch_async(self, dets, counts, slot):
        """dets: [n, cap] DET_DTYPE array, counts: [n] int32 (ideally views of pinned memory)."""
        _lib.check(_lib.lib().yb_engine_fetch_async(self._h, dets.ctypes.data, counts.ctypes.data, dets.shape[1], slot))

    def fetch_wait(self, slot):
        _lib.check(_lib.lib().yb_engine_fetch_wait(self._h, slot))

    def profiling(self, enable):
        _lib.check(_lib.lib().yb_engine_profiling(self._h, int(enable)))

    def profile_read(self):
        """[(plan layer, summed ms)] per launched op and the number of forwards they were summed over."""
        cap = 4 * len(self.plan)
        idx = np.zeros(cap, dtype=np.int32)
        ms = np.zeros(cap, dtype=np.float32)
        n_ops, n_fwd = ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().yb_engine_profile_read(self._h, idx.ctypes.data, ms.ctypes.data, cap,
                                                     ctypes.byref(n_ops), ctypes.byref(n_fwd)))
        return [(int(idx[i]), float(ms[i])) for i in range(min(n_ops.value, cap))], n_fwd.value

    def op_info(self, op_index):
        layer, path, bn, bk, st = (ctypes.c_int() for _ in range(5))
        fl = ctypes.c_double()
        _lib.check(_lib.lib().yb_engine_op_info(self._h, op_index, ctypes.byref(layer), ctypes.byref(path), ctypes.byref(bn),
                                                ctypes.byref(bk), ctypes.byref(st), ctypes.byref(fl)))
        return {"layer": layer.value, "path": path.value, "bn": bn.value, "bk": bk.value, "stages": st.value,
                "flops_per_image": fl.value}

    def op_cfg(self, op_index):
        v = [ctypes.c_int() for _ in range(6)]
        _lib.check(_lib.lib().yb_engine_op_cfg(self._h, op_index, *[ctypes.byref(x) for x in v]))
        cfg = dict(zip(("bn", "pair", "bstat", "tma_epi", "stages", "ksub"), [x.value for x in v]))
        sk = ctypes.c_int(1)
        _lib.check(_lib.lib().yb_engine_op_splitk(self._h, op_index, ctypes.byref(sk)))
        cfg["splitk"] = sk.value
        return cfg

    def set_conv_impl(self, impl):
        _lib.check(_lib.lib().yb_engine_set_conv_impl(self._h, impl))

    def autotune(self, n=None, reps=5):
        """Times every launch configuration of each tcgen05 conv on the device for a batch of n and keeps the fastest.
        Returns the report (dict).  Outputs are unchanged; the activations are clobbered until the next forward."""
        import json
        _lib.check(_lib.lib().yb_engine_autotune(self._h, int(n or self.max_batch), int(reps)))
        self.last_n = 0
        return json.loads(self.tune_report())

    def tune_report(self):
        need = ctypes.c_size_t(0)
        _lib.check(_lib.lib().yb_engine_tune_report(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _lib.check(_lib.lib().yb_engine_tune_report(self._h, buf, need.value, None))
        return buf.value.decode("utf-8") or "{}"

    def set_option(self, name, value):
        _lib.check(_lib.lib().yb_engine_set_option(self._h, name.encode("ascii"), int(value)))

    def set_conv_cfg(self, op_index, bn, pair=0, bstat=-1, tma_epi=-1, ksub=0):
        _lib.check(_lib.lib().yb_engine_set_conv_cfg(self._h, op_index, bn, pair, bstat, tma_epi, ksub))

    def time_op(self, op_index, n=None, reps=10):
        ms = ctypes.c_float()
        _lib.check(_lib.lib().yb_engine_time_op(self._h, op_index, int(n or self.max_batch), reps, ctypes.byref(ms)))
        self.last_n = 0
        return ms.value

    CYCLE_NAMES = ("prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tmem", "mma_total", "epi_wait_acc",
                   "epi_wait_res", "epi_wait_buf", "epi_tmem_ld", "epi_math", "epi_total", "epi_fence_store")

    def read_cycles(self, reset=True):
        """In-kernel cycle counters of the conv roles (set_option("cycles", 1) first), summed over CTAs and launches."""
        out = np.zeros(16, dtype=np.uint64)
        _lib.check(_lib.lib().yb_engine_read_cycles(self._h, out.ctypes.data, int(reset)))
        return dict(zip(self.CYCLE_NAMES, [int(v) for v in out]))

    def launch_count(self):
        a, b = ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().yb_engine_launch_count(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def graph_replays(self):
        """Forwards launched by replaying a captured CUDA graph (set_option("graph", -1 | 0 | 1))."""
        r = ctypes.c_int()
        _lib.check(_lib.lib().yb_engine_graph_replays(self._h, ctypes.byref(r)))
        return r.value

    def profile(self, images):
        """[(plan layer index, milliseconds)] per launched op for one forward (CUDA events)."""
        ptr, mem, dt, keep = _buffer(images, (np.dtype(np.float32), np.dtype(np.uint8)))
        cap = 4 * len(self.plan)
        idx = np.zeros(cap, dtype=np.int32)
        ms = np.zeros(cap, dtype=np.float32)
        n_ops = ctypes.c_int()
        _lib.check(_lib.lib().yb_engine_profile(self._h, ptr, _lib.YB_F32 if dt == np.float32 else _lib.YB_U8, mem,
                                                images.shape[0], idx.ctypes.data, ms.ctypes.data, cap, ctypes.byref(n_ops)))
        self.last_n = images.shape[0]
        return [(int(idx[i]), float(ms[i])) for i in range(min(n_ops.value, cap))]


class PostProcessor(object):
    """Stand-alone decode + NMS over head tensors in the reference's net[-1].out layout (yb_post)."""

    def __init__(self, scales, num_classes, decode_mode, max_batch=1, device=0):
        """scales: [(h, w, [(aw, ah), ...])] with anchors in grid units, in detection order."""
        self._h = None
        arr = (_lib.yb_scale * len(scales))()
        self.rows = 0
        for i, (h, w, anchors) in enumerate(scales):
            arr[i].h, arr[i].w, arr[i].n_anchors = int(h), int(w), len(anchors)
            for j, (aw, ah) in enumerate(anchors):
                arr[i].anchors[2 * j], arr[i].anchors[2 * j + 1] = float(aw), float(ah)
            self.rows += int(h) * int(w) * len(anchors)
        self.box_len = 5 + num_classes
        self.max_batch = max_batch
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().yb_post_create(arr, len(scales), num_classes, decode_mode, max_batch, device,
                                             ctypes.byref(handle)))
        self._h = handle

    def close(self):
        if self._h is not None:
            _lib.lib().yb_post_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _head(self, head):
        ptr, mem, _, keep = _buffer(head, (np.dtype(np.float32),))
        n = int(head.shape[0])
        if int(np.prod(head.shape[1:])) != self.rows * self.box_len:
            raise ValueError("head tensor has {} values per image, expected {}".format(
                int(np.prod(head.shape[1:])), self.rows * self.box_len))
        return ptr, mem, n, keep

    def run(self, head, threshold, iou_threshold, nms_mode=YB_NMS_REFERENCE, max_per_image=None, fetch=True):
        ptr, mem, n, keep = self._head(head)
        ps = _producer_stream(head)
        if ps is not None:
            _lib.check(_lib.lib().yb_post_order_after(self._h, ps))
        if not fetch:
            _lib.check(_lib.lib().yb_post_run(self._h, ptr, mem, n, threshold, iou_threshold, nms_mode, None, None, 0, None))
            if ps is not None:
                _lib.check(_lib.lib().yb_post_order_before(self._h, ps))
            self._keep = keep
            return None
        cap = int(max_per_image or self.rows)
        dets = np.zeros((n, cap), dtype=DET_DTYPE)
        counts = np.zeros(n, dtype=np.int32)
        cands = np.zeros(n, dtype=np.int32)
        _lib.check(_lib.lib().yb_post_run(self._h, ptr, mem, n, threshold, iou_threshold, nms_mode,
                                          dets.ctypes.data, counts.ctypes.data, cap, cands.ctypes.data))
        self.last_candidates = cands
        return [dets[i, :min(int(counts[i]), cap)] for i in range(n)]

    def decode(self, head, threshold):
        """All candidates (before NMS) per image, sorted by row."""
        ptr, mem, n, keep = self._head(head)
        cap = self.rows
        dets = np.zeros((n, cap), dtype=DET_DTYPE)
        counts = np.zeros(n, dtype=np.int32)
        _lib.check(_lib.lib().yb_post_decode(self._h, ptr, mem, n, threshold, dets.ctypes.data, counts.ctypes.data, cap))
        return [dets[i, :int(counts[i])] for i in range(n)]

    def sync(self):
        _lib.check(_lib.lib().yb_post_sync(self._h))

    def last_ms(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        _lib.check(_lib.lib().yb_post_last_ms(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value


def nms(x, y, w, h, prob, iou_threshold, class_idx=None, nms_mode=YB_NMS_REFERENCE, device=0):
    """Greedy NMS on the device with the reference's semantics (net/base.py:195-209).  If any coordinate
    array is float64 the IoU is evaluated in float64 (as numpy would), otherwise in float32.
    Returns kept indices in kept order."""
    arrs = [np.asarray(a) for a in (x, y, w, h)]
    f64 = any(a.dtype == np.float64 for a in arrs)
    dt = np.float64 if f64 else np.float32
    arrs = [np.ascontiguousarray(a, dtype=dt) for a in arrs]
    prob = np.ascontiguousarray(prob, dtype=np.float32)
    k = prob.size
    cls = None if class_idx is None else np.ascontiguousarray(class_idx, dtype=np.int32)
    keep = np.zeros(max(k, 1), dtype=np.int32)
    n_keep = ctypes.c_int(0)
    _lib.check(_lib.lib().yb_nms(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data,
                                 prob.ctypes.data, None if cls is None else cls.ctypes.data, k, int(f64),
                                 float(iou_threshold), nms_mode, device, keep.ctypes.data, ctypes.byref(n_keep)))
    return keep[:n_keep.value].astype(np.int64)
