"""CPU: the oracle restatements against the committed golden vectors (made by the unmodified reference,
oracle/make_golden.py).  NMS must be bit-identical; decode values to 1e-6 (numpy's exp is not guaranteed
bit-stable across CPU SIMD levels); the conv stack to fp32 rounding."""
import numpy as np

import helpers
from oracle import convstack, make_golden, postprocess
from tensorflow_yolo_b200 import synth


def _check_cands(c, g, i):
    assert np.array_equal(c["row"], g["cand%d_row" % i])
    assert np.array_equal(c["class_idx"], g["cand%d_class_idx" % i])
    for k in ("x", "y", "w", "h", "prob"):
        assert c[k].dtype == g["cand%d_%s" % (i, k)].dtype
        np.testing.assert_allclose(c[k], g["cand%d_%s" % (i, k)], rtol=1e-6, atol=0)


def test_decode_nms_v3_matches_reference_golden():
    g = helpers.golden("post_v3.npz")
    head = make_golden.post_v3_head()
    topo = convstack.topology_v3(80, np.reshape(helpers.V3_ANCHORS, [-1, 2]), make_golden.POST_V3_SHAPE)
    geo = convstack.yolo_geometry(topo, make_golden.POST_V3_SHAPE)
    for i, img in enumerate(head):
        c = postprocess.decode_v3_image(img, geo, float(g["threshold"]))
        _check_cands(c, g, i)
        # NMS on the reference's own decoded boxes: bit-identical kept list (order included)
        ref_c = {k: g["cand%d_%s" % (i, k)] for k in ("x", "y", "w", "h", "prob", "class_idx", "row")}
        kept = postprocess.nms(ref_c, float(g["iou_threshold"]))
        assert np.array_equal(ref_c["row"][kept], g["kept%d_row" % i])


def test_decode_nms_v2_matches_reference_golden():
    g = helpers.golden("post_v2.npz")
    head = make_golden.post_v2_head()
    anchors = np.reshape(make_golden.V2_ANCHORS, [-1, 2])
    h5 = np.reshape(head, [-1, 4, 6, 5, 25])
    for i, img in enumerate(h5):
        c = postprocess.decode_v2(img, anchors, float(g["threshold"]))
        _check_cands(c, g, i)
        ref_c = {k: g["cand%d_%s" % (i, k)] for k in ("x", "y", "w", "h", "prob", "class_idx", "row")}
        kept = postprocess.nms(ref_c, float(g["iou_threshold"]))
        assert np.array_equal(ref_c["row"][kept], g["kept%d_row" % i])


def test_nms_adversarial_cases_bit_identical():
    g = helpers.golden("nms_cases.npz")
    for ci in range(int(g["n_cases"])):
        for regime, (xy_t, wh_t) in (("f64", (np.float32, np.float64)), ("f32", (np.float32, np.float32)),
                                      ("d64", (np.float64, np.float64))):
            c = {"x": g["case%d_in_x" % ci].astype(xy_t), "y": g["case%d_in_y" % ci].astype(xy_t),
                 "w": g["case%d_in_w" % ci].astype(wh_t), "h": g["case%d_in_h" % ci].astype(wh_t),
                 "prob": g["case%d_in_prob" % ci].astype(np.float32)}
            with np.errstate(all="ignore"):
                kept = postprocess.nms(c, 0.6)
            assert np.array_equal(kept, g["case%d_%s" % (ci, regime)]), (ci, regime)


def test_nms_empty():
    c = {k: np.zeros(0, np.float32) for k in ("x", "y", "w", "h", "prob")}
    assert len(postprocess.nms(c, 0.6)) == 0


def test_convstack_v3_matches_reference_over_stub():
    g = helpers.golden("conv_v3.npz")
    shape = make_golden.CONV_V3_SHAPE
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x = synth.images(2, shape[0], shape[1], seed=1)
    y = convstack.forward(topo, stream, x)
    assert y.shape == g["net_out"].shape
    assert helpers.rel_err(y, g["net_out"]) < 1e-5
    assert list(g["variable_names"]) == sum([l.variable_names for l in net], [])


def test_convstack_v2_matches_reference_over_stub():
    g = helpers.golden("conv_v2.npz")
    shape = make_golden.CONV_V2_SHAPE
    net, topo, stream = helpers.build_v2(shape, 20, seed=3)
    x = synth.images(2, shape[0], shape[1], seed=4)
    y = convstack.forward(topo, stream, x)
    assert y.shape == g["net_out"].shape
    assert helpers.rel_err(y, g["net_out"]) < 1e-5
    assert list(g["variable_names"]) == sum([l.variable_names for l in net], [])
