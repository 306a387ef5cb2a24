"""Shared test helpers: network construction on both sides (product plan / oracle topology), goldens."""
import os

import numpy as np

from oracle import convstack
from tensorflow_yolo_b200 import synth
from tensorflow_yolo_b200.net import v2 as pv2, v3 as pv3

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
V3_ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]
V2_ANCHORS_VOC = [1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071]
V2_ANCHORS_COCO = [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828]


def names(n):
    return ["c%d" % i for i in range(n)]


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def build_v3(shape, num_classes=80, seed=2, **kw):
    """(product layer list, oracle topology, weight stream)"""
    anchors = np.reshape(V3_ANCHORS, [-1, 2])
    net = pv3.create_network(anchors, names(num_classes), False, input_shape=shape)
    specs = net[0]._yb_state.graph.specs
    stream = synth.weight_stream(specs, seed=seed, num_classes=num_classes, **kw)
    topo = convstack.topology_v3(num_classes, anchors, shape)
    return net, topo, stream


def build_v2(shape, num_classes=20, seed=3, anchors=None, **kw):
    anchors = np.reshape(V2_ANCHORS_VOC if anchors is None else anchors, [-1, 2])
    net = pv2.create_full_network(anchors, names(num_classes), False, input_shape=shape)
    specs = net[0]._yb_state.graph.specs
    stream = synth.weight_stream(specs, seed=seed, num_classes=num_classes, **kw)
    topo = convstack.topology_v2(num_classes, len(anchors), shape)
    return net, topo, stream


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cand_from_dets(d):
    """structured yb_det array -> oracle-style candidate dict"""
    return {"x": d["x"].astype(np.float32), "y": d["y"].astype(np.float32), "w": d["w"].astype(np.float64),
            "h": d["h"].astype(np.float64), "prob": d["prob"].astype(np.float32),
            "class_idx": d["class_idx"].astype(np.int64), "row": d["row"].astype(np.int64)}
