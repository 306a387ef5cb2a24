"""Flat layer plan: the data-only form of a YOLO network handed to the C-ABI engine.

A plan is a list of :class:`LayerSpec`, one per entry of the reference's ``layers`` list
(net/v3.py:8-94 -> 109 entries, net/v2.py:10-60 -> 32 entries); indices are load-bearing because
routes and shortcuts address other entries by position.  The layer classes in
``tensorflow_yolo_b200.net.layers`` append to a plan instead of building a TensorFlow graph.
"""
import ctypes

KIND_INPUT, KIND_CONV, KIND_MAXPOOL, KIND_ROUTE, KIND_REORG = 0, 1, 2, 3, 4
KIND_SHORTCUT, KIND_UPSAMPLE, KIND_YOLO, KIND_DETECTION = 5, 6, 7, 8
KIND_NAMES = ["input", "conv", "maxpool", "route", "reorg", "shortcut", "upsample", "yolo", "detection"]

YB_MAX_SRC = 4
YB_MAX_ANCHORS = 16


class yb_layer(ctypes.Structure):
    """Mirror of ``struct yb_layer`` in include/yolo_b200.h."""
    _fields_ = [
        ("kind", ctypes.c_int),
        ("filters", ctypes.c_int),       # conv: output channels
        ("ksize", ctypes.c_int),         # conv / maxpool window
        ("stride", ctypes.c_int),        # conv / maxpool / reorg / upsample factor
        ("batch_norm", ctypes.c_int),    # conv: 1 = BN (no bias), 0 = bias
        ("leaky", ctypes.c_int),         # conv: 1 = leaky(0.1), 0 = linear
        ("n_src", ctypes.c_int),
        ("src", ctypes.c_int * YB_MAX_SRC),
        ("n_anchors", ctypes.c_int),     # yolo
        ("anchors", ctypes.c_float * (2 * YB_MAX_ANCHORS)),   # yolo: (w,h) pairs in grid units
    ]


class LayerSpec(object):
    __slots__ = ("kind", "filters", "ksize", "stride", "batch_norm", "leaky", "src", "anchors", "shape")

    def __init__(self, kind, shape, src=(), filters=0, ksize=0, stride=1, batch_norm=False, leaky=False,
                 anchors=()):
        self.kind = kind
        self.shape = tuple(int(s) for s in shape)          # (h, w, c) of the output
        self.src = [int(s) for s in src]
        self.filters = int(filters)
        self.ksize = int(ksize)
        self.stride = int(stride)
        self.batch_norm = bool(batch_norm)
        self.leaky = bool(leaky)
        self.anchors = [(float(a[0]), float(a[1])) for a in anchors]

    def to_c(self):
        c = yb_layer()
        c.kind, c.filters, c.ksize, c.stride = self.kind, self.filters, self.ksize, self.stride
        c.batch_norm, c.leaky = int(self.batch_norm), int(self.leaky)
        if len(self.src) > YB_MAX_SRC:
            raise ValueError("too many sources for one layer")
        c.n_src = len(self.src)
        for i, s in enumerate(self.src):
            c.src[i] = s
        if len(self.anchors) > YB_MAX_ANCHORS:
            raise ValueError("too many anchors for one yolo layer")
        c.n_anchors = len(self.anchors)
        for i, (aw, ah) in enumerate(self.anchors):
            c.anchors[2 * i], c.anchors[2 * i + 1] = aw, ah
        return c

    def as_dict(self):
        d = {"kind": KIND_NAMES[self.kind], "src": list(self.src), "shape": list(self.shape)}
        if self.kind == KIND_CONV:
            d.update(filters=self.filters, ksize=self.ksize, stride=self.stride,
                     bn=self.batch_norm, leaky=self.leaky)
        elif self.kind in (KIND_MAXPOOL, KIND_REORG, KIND_UPSAMPLE):
            d.update(stride=self.stride)
        if self.kind == KIND_MAXPOOL:
            d.update(size=self.ksize)
        if self.kind == KIND_YOLO:
            d.update(anchors=[list(a) for a in self.anchors])
        return d


def to_c_array(plan):
    arr = (yb_layer * len(plan))()
    for i, spec in enumerate(plan):
        arr[i] = spec.to_c()
    return arr


def conv_specs(plan):
    """[(layer index, cin, cout, ksize, batch_norm, feeds_shortcut, is_head)] in weight-stream order."""
    consumers = {}
    for i, spec in enumerate(plan):
        for s in spec.src:
            consumers.setdefault(s, []).append(i)
    out = []
    for i, spec in enumerate(plan):
        if spec.kind != KIND_CONV:
            continue
        cin = plan[spec.src[0]].shape[2]
        feeds = any(plan[c].kind == KIND_SHORTCUT and plan[c].src[0] == i for c in consumers.get(i, []))
        out.append((i, cin, spec.filters, spec.ksize, spec.batch_norm, feeds, not spec.batch_norm))
    return out


def weight_count(plan):
    """Number of float32 values the darknet stream must hold (net/base.py:26-46 walk)."""
    n = 0
    for _, cin, cout, k, bn, _, _ in conv_specs(plan):
        n += (4 * cout if bn else cout) + cout * cin * k * k
    return n


def conv_flops(plan):
    """2*MAC over all convs, per image."""
    f = 0
    for spec in plan:
        if spec.kind == KIND_CONV:
            h, w, c = spec.shape
            cin = plan[spec.src[0]].shape[2]
            f += 2 * h * w * c * spec.ksize * spec.ksize * cin
    return f
