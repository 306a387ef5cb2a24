"""ctypes binding of libyolo_b200.so (include/yolo_b200.h).  Fails loudly when the library is missing."""
import ctypes
import os

from . import plan as _plan

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libyolo_b200.so")

YB_OK, YB_ERR_INVALID, YB_ERR_CUDA, YB_ERR_SHORT_WEIGHTS, YB_ERR_CAPACITY, YB_ERR_STATE = 0, -1, -2, -3, -4, -5
YB_MEM_HOST, YB_MEM_DEVICE = 0, 1
YB_F32, YB_U8 = 0, 1
YB_DECODE_V3, YB_DECODE_V2 = 0, 1
YB_NMS_REFERENCE, YB_NMS_PER_CLASS = 0, 1


class yb_det(ctypes.Structure):
    _fields_ = [("w", ctypes.c_double), ("h", ctypes.c_double), ("x", ctypes.c_float), ("y", ctypes.c_float),
                ("prob", ctypes.c_float), ("class_idx", ctypes.c_int32), ("row", ctypes.c_int32),
                ("pad_", ctypes.c_int32)]


class yb_scale(ctypes.Structure):
    _fields_ = [("h", ctypes.c_int), ("w", ctypes.c_int), ("n_anchors", ctypes.c_int),
                ("anchors", ctypes.c_float * (2 * _plan.YB_MAX_ANCHORS))]


# numpy view of yb_det
DET_DTYPE = [("w", "<f8"), ("h", "<f8"), ("x", "<f4"), ("y", "<f4"), ("prob", "<f4"),
             ("class_idx", "<i4"), ("row", "<i4"), ("pad_", "<i4")]


class YoloB200Error(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libyolo_b200 error {}: {}".format(code, message))
        self.code = code


_P = ctypes.POINTER
_SIGNATURES = {
    "yb_abi_version": (ctypes.c_int, []),
    "yb_last_error": (ctypes.c_char_p, []),
    "yb_device_count": (ctypes.c_int, [_P(ctypes.c_int)]),
    "yb_crc32c": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, _P(ctypes.c_uint32)]),
    "yb_engine_create": (ctypes.c_int, [_P(_plan.yb_layer), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P(ctypes.c_void_p)]),
    "yb_engine_destroy": (None, [ctypes.c_void_p]),
    "yb_engine_load_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, _P(ctypes.c_size_t)]),
    "yb_engine_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "yb_engine_forward_raw": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_void_p), _P(ctypes.c_int), _P(ctypes.c_int),
                                             _P(ctypes.c_int), ctypes.c_int]),
    "yb_engine_read_input_u8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "yb_resize_bgr2rgb": (ctypes.c_int, [_P(ctypes.c_void_p), _P(ctypes.c_int), _P(ctypes.c_int), _P(ctypes.c_int), ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "yb_engine_read_output": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "yb_engine_output_shape": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_int), _P(ctypes.c_int)]),
    "yb_engine_read_layer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                            _P(ctypes.c_int * 3)]),
    "yb_engine_detect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "yb_engine_detect_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    "yb_engine_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_engine_order_after": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "yb_engine_order_before": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "yb_post_order_after": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "yb_post_order_before": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "yb_engine_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _P(ctypes.c_int)]),
    "yb_engine_set_conv_impl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_engine_autotune": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "yb_engine_tune_report": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, _P(ctypes.c_size_t)]),
    "yb_engine_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "yb_engine_set_conv_cfg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int]),
    "yb_engine_time_op": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P(ctypes.c_float)]),
    "yb_engine_read_cycles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "yb_engine_launch_count": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_int), _P(ctypes.c_int)]),
    "yb_engine_graph_replays": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_int)]),
    "yb_engine_mark": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_engine_elapsed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _P(ctypes.c_float)]),
    "yb_engine_fetch_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "yb_engine_fetch_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_engine_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_engine_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                              _P(ctypes.c_int), _P(ctypes.c_int)]),
    "yb_engine_op_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _P(ctypes.c_int), _P(ctypes.c_int), _P(ctypes.c_int),
                                         _P(ctypes.c_int), _P(ctypes.c_int), _P(ctypes.c_double)]),
    "yb_engine_op_cfg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int] + [_P(ctypes.c_int)] * 6),
    "yb_engine_op_splitk": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _P(ctypes.c_int)]),
    "yb_post_create": (ctypes.c_int, [_P(yb_scale), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, _P(ctypes.c_void_p)]),
    "yb_post_destroy": (None, [ctypes.c_void_p]),
    "yb_post_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_void_p]),
    "yb_post_decode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "yb_post_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_post_last_ms": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_float), _P(ctypes.c_float)]),
    "yb_nms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p, _P(ctypes.c_int)]),
}

EXPORTS = sorted(_SIGNATURES)
_lib = None


def lib():
    """Loads libyolo_b200.so (once).  There is no Python or CPU substitute for it."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "{} is missing: build it with `python -m tensorflow_yolo_b200.build` (needs nvcc). "
            "tensorflow_yolo_b200 has no fallback implementation.".format(LIB_PATH))
    handle = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(handle, name)     # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = restype, argtypes
    if handle.yb_abi_version() != 2:
        raise ImportError("libyolo_b200.so ABI version mismatch")
    _lib = handle
    return _lib


def check(code):
    if code != YB_OK:
        raise YoloB200Error(code, lib().yb_last_error().decode("utf-8", "replace"))


def crc32c(data, seed=0, impl=0):
    """CRC-32C of a bytes-like object / contiguous numpy array (native, SSE4.2 when available)."""
    import numpy as np
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    out = ctypes.c_uint32(0)
    check(lib().yb_crc32c(buf.ctypes.data if buf.size else None, buf.size, seed, impl, ctypes.byref(out)))
    return out.value


def device_count():
    c = ctypes.c_int(0)
    check(lib().yb_device_count(ctypes.byref(c)))
    return c.value
