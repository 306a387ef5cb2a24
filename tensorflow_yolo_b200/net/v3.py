"""YOLOv3 builder, weight loader and post-processing with the reference's signatures (net/v3.py).

``create_network`` returns the same 109-entry layer list as the reference (indices are load-bearing:
entry 62 and entry 37 are the route sources "61 + 1" and "36 + 1"), built from a stage table instead of
a TensorFlow graph.  ``find_bounding_boxes`` runs decode + NMS on the GPU (yb_post / yb_engine_detect).
"""
import numpy as np

from . import base
from . import layers as L
from .. import engine as _engine

# Darknet-53 trunk: (filters of the stride-2 conv, residual blocks that follow it) -- net/v3.py:25-40
_TRUNK = [(64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)]
# the three detection branches: bottleneck width and the trunk entry concatenated after upsampling
# (net/v3.py:44-85; None = first branch, no lateral input)
_BRANCHES = [(512, None), (256, 61 + 1), (128, 36 + 1)]


@staticmethod
def create_network(anchors, class_names, is_training, scope="yolo", input_shape=(416, 416, 3)):
    num_classes = len(class_names)
    per_scale = np.reshape(anchors, [3, -1, 2])[::-1, :, :]      # largest anchors go to the coarsest grid
    L.conv2d_bn_act.reset()
    graph = L.reset_default_graph()
    net = []

    def conv(filters, k, stride=1, **kw):
        net.append(L.conv2d_bn_act(net[-1].out, filters, k, stride, is_training=is_training, scope=scope, **kw))

    net.append(L.input_layer([None, input_shape[0], input_shape[1], input_shape[2]], "input"))
    conv(32, 3)
    for filters, blocks in _TRUNK:
        conv(filters, 3, 2)
        for _ in range(blocks):
            conv(filters // 2, 1)
            conv(filters, 3)
            net.append(L.shortcut(net[-1].out, net[-3].out))

    yolos = []
    for width, lateral in _BRANCHES:
        if lateral is not None:
            net.append(L.route([net[-4].out]))
            conv(width, 1)
            net.append(L.upsample(net[-1].out, 2))
            net.append(L.route([net[-1].out, net[lateral].out]))
        for _ in range(3):
            conv(width, 1)
            conv(width * 2, 3)
        sub_anchors = per_scale[len(yolos)]
        conv(len(sub_anchors) * (5 + num_classes), 1, 1, use_batch_normalization=False, activation_fn="linear")
        net.append(L.yolo_layer(net[-1].out, sub_anchors, num_classes, input_shape))
        yolos.append(net[-1])

    net.append(L.detection_layer(yolos))
    net[-1].out.name = "output"
    state = base.NetworkState(graph, "v3", num_classes, input_shape=input_shape)
    graph._yb_state = state
    net[0]._yb_state = state
    return net


@staticmethod
def load_weights(layers, weights_path):
    print("Reading pre-trained weights from {}".format(weights_path))
    with open(weights_path, "rb") as f:
        header = np.fromfile(f, count=5, dtype=np.int32)       # major, minor, revision, subversion, n
        print(" ".join(str(v) for v in header))
        weights = np.fromfile(f, dtype=np.float32)
    print("Found {} weight values.".format(len(weights)))
    return base.load_weights(layers, weights)


def _scales(net):
    return [(l.h, l.w, l.anchors) for l in net[-1].yolos]


@staticmethod
def find_bounding_boxes(net_out, net, threshold, iou_threshold, anchors, class_names):
    """net_out: [n, R, 5+C] host array (what Session.run returned).  Decode + NMS run on the GPU."""
    net_out = np.ascontiguousarray(net_out, dtype=np.float32)
    state = base.state_of(net)
    post = getattr(state, "post", None)
    if post is None or post.max_batch < net_out.shape[0]:
        post = _engine.PostProcessor(_scales(net), len(class_names), _engine.YB_DECODE_V3,
                                     max_batch=net_out.shape[0], device=state.device)
        state.post = post
    return [base.boxes_from_dets(d) for d in post.run(net_out, threshold, iou_threshold)]
