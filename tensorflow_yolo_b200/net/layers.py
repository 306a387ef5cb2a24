"""Layer wrappers with the reference's constructor signatures, recording a flat plan.

Drop-in for the reference's ``net/layers.py`` on the TEST path: every class keeps the name,
positional arguments and the two attributes its callers rely on -- ``.out`` (a handle fed to the
next layer) and ``.variable_names`` (the darknet weight-stream order) -- but instead of emitting
TensorFlow ops each constructor appends one :class:`~tensorflow_yolo_b200.plan.LayerSpec` to the
current :class:`Graph`.  The plan is later compiled by the CUDA engine; nothing here computes.

Reference semantics each spec stands for (all in /root/reference/net/layers.py):
  conv2d_bn_act :17-67   pad (k-1)//2 before / rest after only when stride>1, then VALID; SAME when
                         stride==1; bias iff no batch-norm; BN eps 1e-5; leaky alpha 0.1
  max_pool2d    :70-81   zero pad (0,1) then 2x2/2 VALID
  route         :84-87   channel concat in argument order
  reorg         :90-97   extract_image_patches == space-to-depth, channel order (dy,dx,c)
  shortcut      :100-103 prev + shortcut_out
  input_layer   :106-109 float32 placeholder [None,H,W,C]
  upsample      :112-116 nearest neighbour, in[y//s, x//s]
  detection_layer :119-123 concat of the yolo layers on axis 1
  yolo_layer    :126-134 rows (cy*w+cx)*b+a; anchors divided by the stride
"""
from .. import plan as _plan

_BATCH_NORM_EPSILON = 1e-5
_LEAKY_RELU = 0.1


class Graph(object):
    """The plan under construction (stands in for TF's default graph)."""

    def __init__(self):
        self.specs = []

    def add(self, spec):
        self.specs.append(spec)
        return len(self.specs) - 1


_default_graph = Graph()


def reset_default_graph():
    global _default_graph
    _default_graph = Graph()
    return _default_graph


def get_default_graph():
    return _default_graph


class _StaticShape(object):
    def __init__(self, dims):
        self._dims = list(dims)

    def as_list(self):
        return list(self._dims)


class SymbolicTensor(object):
    """What ``layer.out`` is: (graph, producing plan index, static NHWC shape)."""

    def __init__(self, graph, index, hwc, leading=(None,), name=None):
        self.graph, self.index, self.name = graph, index, name
        self._dims = list(leading) + list(hwc)

    def get_shape(self):
        return _StaticShape(self._dims)

    @property
    def shape(self):
        return _StaticShape(self._dims)

    def __add__(self, other):
        # the reference's shortcut is literally ``prev + shortcut_out``
        return shortcut(self, other).out


def _hwc(t):
    return tuple(t.get_shape().as_list()[1:4])


def _emit(graph, spec, leading=(None,)):
    idx = graph.add(spec)
    return SymbolicTensor(graph, idx, spec.shape, leading)


class conv2d_bn_act(object):
    name_count = 0

    def __init__(self, prev, filter_size, kernel_size, stride=1, use_batch_normalization=True,
                 activation_fn="leaky", is_training=False, scope="yolo"):
        if is_training:
            raise NotImplementedError("tensorflow_yolo_b200 implements the TEST path only")
        name = "conv2d_bn_act_{}".format(conv2d_bn_act.name_count)
        conv2d_bn_act.name_count += 1
        h, w, _ = _hwc(prev)
        if stride > 1:
            # explicit (k-1)-pixel zero pad then VALID  ->  floor((h + k-1 - k)/s) + 1
            ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        else:
            ho, wo = h, w
        spec = _plan.LayerSpec(_plan.KIND_CONV, (ho, wo, filter_size), src=[prev.index],
                               filters=filter_size, ksize=kernel_size, stride=stride,
                               batch_norm=use_batch_normalization, leaky=(activation_fn == "leaky"))
        self.out = _emit(prev.graph, spec)
        stem = "{}/{}/".format(scope, name)
        per_channel = ["beta", "gamma", "moving_mean", "moving_variance"] if use_batch_normalization else ["bias"]
        self.variable_names = [stem + v for v in per_channel + ["kernel"]]

    @staticmethod
    def reset():
        conv2d_bn_act.name_count = 0


class max_pool2d(object):
    def __init__(self, prev, kernel_size, stride=2):
        h, w, c = _hwc(prev)
        if stride > 1:
            ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        else:
            ho, wo = h, w
        if kernel_size != 2 or stride != 2 or h % 2 or w % 2:
            raise NotImplementedError("only the 2x2/2 max pool on even maps used by YOLOv2 is implemented")
        spec = _plan.LayerSpec(_plan.KIND_MAXPOOL, (ho, wo, c), src=[prev.index], ksize=kernel_size, stride=stride)
        self.out = _emit(prev.graph, spec)
        self.variable_names = []


class route(object):
    def __init__(self, prevs):
        h, w, _ = _hwc(prevs[0])
        for p in prevs:
            if _hwc(p)[:2] != (h, w):
                raise ValueError("route inputs must share their spatial size")
        c = sum(_hwc(p)[2] for p in prevs)
        spec = _plan.LayerSpec(_plan.KIND_ROUTE, (h, w, c), src=[p.index for p in prevs])
        self.out = _emit(prevs[0].graph, spec)
        self.variable_names = []


class reorg(object):
    def __init__(self, prev, stride):
        h, w, c = _hwc(prev)
        spec = _plan.LayerSpec(_plan.KIND_REORG, (h // stride, w // stride, c * stride * stride),
                               src=[prev.index], stride=stride)
        self.out = _emit(prev.graph, spec)
        self.variable_names = []


class shortcut(object):
    def __init__(self, prev, shortcut_out):
        if _hwc(prev) != _hwc(shortcut_out):
            raise ValueError("shortcut operands differ in shape")
        spec = _plan.LayerSpec(_plan.KIND_SHORTCUT, _hwc(prev), src=[prev.index, shortcut_out.index])
        self.out = _emit(prev.graph, spec)
        self.variable_names = []


class input_layer(object):
    def __init__(self, shape, name="input"):
        graph = get_default_graph()
        spec = _plan.LayerSpec(_plan.KIND_INPUT, tuple(shape[1:4]))
        self.out = _emit(graph, spec, leading=(shape[0],))
        self.out.name = name
        self.variable_names = []


class upsample(object):
    def __init__(self, prev, stride):
        h, w, c = _hwc(prev)
        spec = _plan.LayerSpec(_plan.KIND_UPSAMPLE, (h * stride, w * stride, c), src=[prev.index], stride=stride)
        self.out = _emit(prev.graph, spec)
        self.variable_names = []


class detection_layer(object):
    def __init__(self, yolos):
        self.yolos = yolos
        rows = sum(l.out.get_shape().as_list()[1] for l in yolos)
        cols = yolos[0].out.get_shape().as_list()[2]
        graph = yolos[0].out.graph
        spec = _plan.LayerSpec(_plan.KIND_DETECTION, (1, rows, cols), src=[l.out.index for l in yolos])
        idx = graph.add(spec)
        self.out = SymbolicTensor(graph, idx, (rows, cols))
        self.variable_names = []


class yolo_layer(object):
    def __init__(self, prev, sub_anchors, no_c, input_shape):
        self.h, self.w, channels = _hwc(prev)
        stride = (input_shape[0] / self.h, input_shape[1] / self.w)
        self.anchors = [(a[0] / stride[0], a[1] / stride[1]) for a in sub_anchors]  # anchors are (w, h)
        self.b = len(self.anchors)
        if channels != self.b * (5 + no_c):
            raise ValueError("head has {} channels, expected {}".format(channels, self.b * (5 + no_c)))
        spec = _plan.LayerSpec(_plan.KIND_YOLO, (1, self.h * self.w * self.b, 5 + no_c),
                               src=[prev.index], anchors=self.anchors)
        idx = prev.graph.add(spec)
        self.out = SymbolicTensor(prev.graph, idx, (self.h * self.w * self.b, 5 + no_c))
        self.variable_names = []
