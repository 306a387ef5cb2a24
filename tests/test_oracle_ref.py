"""CPU, build container only: re-derives the goldens' facts live from /root/reference (skipped elsewhere)."""
import numpy as np
import pytest

import helpers
from oracle import convstack, postprocess, refimport

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference checkout not present")


def test_reference_dtypes_under_numpy2():
    ref = refimport.load()
    out = np.random.RandomState(0).standard_normal((3, 3, 3, 85)).astype(np.float32)
    anchors = [(np.float64(3.6), np.float64(2.8))] * 3
    b = ref.v3._find_bounding_boxes(out, anchors, 0.0)[0]
    assert (type(b.x), type(b.w), type(b.prob)) == (np.float32, np.float64, np.float32)
    assert type(ref.base.iou_score(b, b)) is np.float64


def test_topology_matches_reference_builders():
    ref = refimport.load()
    for shape in ((64, 64, 3), (416, 416, 3)):
        anchors = np.reshape(helpers.V3_ANCHORS, [-1, 2])
        rnet = ref.v3.create_network(anchors, helpers.names(80), False, input_shape=shape)
        pnet, topo, _ = helpers.build_v3(shape, 80) if shape[0] == 64 else (None, convstack.topology_v3(80, anchors, shape), None)
        assert len(rnet) == len(topo) == 109
        kinds = {"conv2d_bn_act": "conv", "shortcut": "shortcut", "route": "route", "upsample": "upsample",
                 "yolo_layer": "yolo", "detection_layer": "detection", "input_layer": "input"}
        assert [kinds[type(l).__name__] for l in rnet] == [r["kind"] for r in topo]
        shapes = [l.out.get_shape().as_list()[1:] for l in rnet]
        geo = convstack.yolo_geometry(topo, shape)
        assert [(l.h, l.w, l.b, l.anchors) for l in rnet[-1].yolos] == geo
        if pnet is not None:
            assert [l.out.get_shape().as_list()[1:] for l in pnet] == shapes
    anchors2 = np.reshape(helpers.V2_ANCHORS_VOC, [-1, 2])
    rnet = ref.v2.create_full_network(anchors2, helpers.names(20), False, input_shape=(416, 416, 3))
    pnet, topo, _ = helpers.build_v2((416, 416, 3), 20)
    assert len(rnet) == len(topo) == len(pnet) == 32
    assert [l.out.get_shape().as_list()[1:] for l in pnet] == [l.out.get_shape().as_list()[1:] for l in rnet]
    assert [l.variable_names for l in pnet] == [l.variable_names for l in rnet]


def test_decode_and_nms_match_reference_live():
    ref = refimport.load()
    rs = np.random.RandomState(21)
    out = (rs.standard_normal((7, 5, 3, 85)) * 2).astype(np.float32)
    anchors = [(np.float64(1.25), np.float64(2.5)), (np.float64(3.0), np.float64(0.7)), (np.float64(5.5), np.float64(6.1))]
    boxes = ref.v3._find_bounding_boxes(out, anchors, 0.4)
    c = postprocess.decode_v3_scale(out, anchors, 0.4)
    assert len(boxes) == len(c["row"])
    for k in ("x", "y", "w", "h", "prob"):
        assert np.array_equal(np.asarray([getattr(b, k) for b in boxes]), c[k])      # same machine: bit-exact
    assert np.array_equal(np.asarray([b.class_idx for b in boxes]), c["class_idx"])
    for j, b in enumerate(boxes):
        b.j = j
    kept = ref.base.non_maximum_suppression(list(boxes), 0.45)
    assert np.array_equal(np.asarray([b.j for b in kept]), postprocess.nms(c, 0.45))
