"""YOLOv2 builder, weight loader and post-processing with the reference's signatures (net/v2.py:10-119).

Only the inference half of the reference's v2 module is in scope: the loss, optimiser,
batch maker and anchor k-means belong to TRAIN/ANCHOR modes.
"""
import numpy as np

from . import base
from . import layers as L
from .. import engine as _engine

# Darknet-19 body as (filters, kernel) runs separated by max pools -- net/v2.py:20-42
_BODY = [
    [(32, 3)], "pool", [(64, 3)], "pool",
    [(128, 3), (64, 1), (128, 3)], "pool",
    [(256, 3), (128, 1), (256, 3)], "pool",
    [(512, 3), (256, 1), (512, 3), (256, 1), (512, 3)], "pool",
    [(1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3)],
]


@staticmethod
def create_full_network(anchors, class_names, is_training, scope="yolo", input_shape=(416, 416, 3)):
    num_anchors, num_classes = len(anchors), len(class_names)
    L.conv2d_bn_act.reset()
    graph = L.reset_default_graph()
    net = []

    def conv(filters, k, **kw):
        net.append(L.conv2d_bn_act(net[-1].out, filters, k, stride=1, is_training=is_training, scope=scope, **kw))

    net.append(L.input_layer([None, input_shape[0], input_shape[1], input_shape[2]], "input"))
    for item in _BODY:
        if item == "pool":
            net.append(L.max_pool2d(net[-1].out, 2, stride=2))
        else:
            for filters, k in item:
                conv(filters, k)
    conv(1024, 3)
    conv(1024, 3)
    net.append(L.route([net[-9].out]))              # passthrough source: the last 26x26x512 conv
    conv(64, 1)
    net.append(L.reorg(net[-1].out, 2))
    net.append(L.route([net[-1].out, net[-4].out]))
    conv(1024, 3)
    conv(num_anchors * (5 + num_classes), 1, use_batch_normalization=False, activation_fn="linear")
    net[-1].out.name = "output"
    state = base.NetworkState(graph, "v2", num_classes,
                              anchors_v2=[(float(a[0]), float(a[1])) for a in np.reshape(anchors, [-1, 2])],
                              input_shape=input_shape)
    graph._yb_state = state
    net[0]._yb_state = state
    return net


def _read_darknet_v2_header(f):
    """16 bytes: int32 major, minor, revision, then a 4-byte `seen` counter -- float32 for format >= 0.2 (and sane
    version numbers), int32 before (net/v2.py:68-77; the reference keeps it at 4 bytes in both cases)."""
    major, minor, revision = (int(v) for v in np.fromfile(f, count=3, dtype=np.int32))
    new_format = major * 10 + minor >= 2 and major < 1000 and minor < 1000
    seen = np.fromfile(f, count=1, dtype=np.float32 if new_format else np.int32)
    return (major, minor, revision), seen


@staticmethod
def load_weights(layers, weights_path):
    print("Reading pre-trained weights from {}".format(weights_path))
    with open(weights_path, "rb") as f:
        version, seen = _read_darknet_v2_header(f)
        print("major, minor, revision: {}, {}, {}".format(*version))
        print("SEEN: ", seen)
        weights = np.fromfile(f, dtype=np.float32)
    print("Found {} weight values.".format(len(weights)))
    return base.load_weights(layers, weights)


@staticmethod
def find_bounding_boxes(net_out, net, threshold, iou_threshold, anchors, class_names):
    """net_out: [n, h, w, A*(5+C)] host array.  Decode (softmax variant) + NMS run on the GPU."""
    net_out = np.ascontiguousarray(net_out, dtype=np.float32)
    anchors = [(float(a[0]), float(a[1])) for a in np.reshape(anchors, [-1, 2])]
    state = base.state_of(net)
    post = getattr(state, "post", None)
    if post is None or post.max_batch < net_out.shape[0]:
        post = _engine.PostProcessor([(net_out.shape[1], net_out.shape[2], anchors)], len(class_names),
                                     _engine.YB_DECODE_V2, max_batch=net_out.shape[0], device=state.device)
        state.post = post
    return [base.boxes_from_dets(d) for d in post.run(net_out, threshold, iou_threshold)]


@staticmethod
def generate_anchors(params):
    """ANCHOR mode (net/v2.py:298-323): k-means over the normalised (width, height) of every annotated box; the
    centres are scaled to grid units of the network input (input_w / stride by input_h / stride).
    Returns (flat anchor array [w0, h0, w1, h1, ...], set of class names)."""
    grid_w = int(params["input_w"]) / int(params["stride"])
    grid_h = int(params["input_h"]) / int(params["stride"])
    annotations = base.parse_annotations(params["annotation_dir"], params["image_dir"], normalize=True)
    print("{} annotations found.".format(len(annotations)))
    boxes = [box for _, image_boxes in annotations for box in image_boxes]
    sizes = [[float(x2 - x1), float(y2 - y1)] for x1, y1, x2, y2, _ in boxes]
    class_names = {name for _, _, _, _, name in boxes}
    centres = base.run_kmeans(sizes, int(params["num_anchors"]), float(params["tolerate"]))
    return np.reshape([[cw * grid_w, ch * grid_h] for cw, ch in centres], [-1]), class_names
