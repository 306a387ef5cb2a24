"""Generates tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The reference's own functions produce every expected value stored here:
  post_v3.npz / post_v2.npz  net.v3/_v2._find_bounding_boxes + net.base.non_maximum_suppression
  nms_cases.npz              net.base.non_maximum_suppression on adversarial box sets
  conv_v3.npz / conv_v2.npz  net.v3.create_network / net.v2.create_full_network + net.base.load_weights
                             executed over oracle.tfstub (TensorFlow itself is not installable: the
                             op arithmetic is the stub's, the topology/wiring/weight order the reference's)
Inputs are regenerated from seeds by the tests (tensorflow_yolo_b200.synth); only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refimport  # noqa: E402
from tensorflow_yolo_b200 import synth  # noqa: E402
from tensorflow_yolo_b200.net import v2 as pv2, v3 as pv3  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
V3_ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]
V2_ANCHORS = [1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071]
POST_V3_SHAPE = (64, 64, 3)
CONV_V3_SHAPE = (96, 64, 3)
CONV_V2_SHAPE = (64, 96, 3)


def boxes_to_arrays(boxes):
    return {"x": np.asarray([b.x for b in boxes], dtype=np.float32), "y": np.asarray([b.y for b in boxes], dtype=np.float32),
            "w": np.asarray([b.w for b in boxes], dtype=np.float64), "h": np.asarray([b.h for b in boxes], dtype=np.float64),
            "prob": np.asarray([b.prob for b in boxes], dtype=np.float32),
            "class_idx": np.asarray([b.class_idx for b in boxes], dtype=np.int64),
            "row": np.asarray([b.row for b in boxes], dtype=np.int64)}


def post_v3_head(n=2, seed=11):
    rows = (2 * 2 + 4 * 4 + 8 * 8) * 3
    t = (np.random.RandomState(seed).standard_normal((n, rows, 85)) * 2.0).astype(np.float32)
    t[..., 4] -= np.float32(0.5)
    return t


def post_v2_head(n=2, seed=12):
    t = (np.random.RandomState(seed).standard_normal((n, 4, 6, 5 * 25)) * 2.0).astype(np.float32)
    return t


def nms_cases():
    """Adversarial candidate sets: list of dicts with x,y,w,h,prob (float64 master copies)."""
    rs = np.random.RandomState(5)
    cases = []
    # 0: exact ties in score, heavy overlap -> first decoded wins
    cases.append(dict(x=[.5, .5, .5, .52], y=[.5, .5, .5, .5], w=[.2, .2, .2, .2], h=[.2, .2, .2, .2], prob=[.9, .9, .9, .9]))
    # 1: IoU exactly at the threshold (two unit-aligned boxes: inter 0.6*1, union 1.0+...): built so iou == 0.6 in float64
    cases.append(dict(x=[0.5, 0.75], y=[0.5, 0.5], w=[1.0, 1.0], h=[1.0, 1.0], prob=[.8, .7]))     # inter .75 union 1.25 -> 0.6
    # 2: zero-area boxes (IoU 0 via union clamp) and identical zero-area boxes
    cases.append(dict(x=[.3, .3, .6], y=[.3, .3, .6], w=[0., 0., .1], h=[.1, .1, 0.], prob=[.5, .6, .7]))
    # 3: chain a>b>c where a suppresses b, b would suppress c but is gone -> c kept
    cases.append(dict(x=[.50, .56, .62], y=[.5, .5, .5], w=[.2, .2, .2], h=[.2, .2, .2], prob=[.9, .8, .7]))
    # 4: single box
    cases.append(dict(x=[.1], y=[.2], w=[.3], h=[.4], prob=[.55]))
    # 5: 70 near-duplicates (crosses the 64-box block of the device sweep) + noise
    n = 70
    cases.append(dict(x=(.5 + rs.uniform(-.01, .01, n)).tolist(), y=(.5 + rs.uniform(-.01, .01, n)).tolist(),
                      w=(.3 + rs.uniform(-.01, .01, n)).tolist(), h=(.3 + rs.uniform(-.01, .01, n)).tolist(),
                      prob=rs.uniform(.5, 1., n).tolist()))
    # 6: 300 random boxes, moderate overlap, many ties in score (quantised)
    n = 300
    cases.append(dict(x=rs.uniform(0, 1, n).tolist(), y=rs.uniform(0, 1, n).tolist(), w=rs.uniform(.05, .4, n).tolist(),
                      h=rs.uniform(.05, .4, n).tolist(), prob=(np.round(rs.uniform(.5, 1., n) * 20) / 20).tolist()))
    # 7: huge / infinite boxes (exp overflow of the decode) and NaN propagation through min/max
    cases.append(dict(x=[.5, .5, .4, .6], y=[.5, .5, .4, .6], w=[np.inf, np.inf, .2, 1e30], h=[np.inf, .5, .2, 1e30], prob=[.9, .8, .7, .6]))
    # 8: negative sizes (never produced by the decode, but the function accepts them)
    cases.append(dict(x=[.5, .5], y=[.5, .5], w=[-.2, .2], h=[.2, .2], prob=[.9, .8]))
    return cases


def main():
    ref = refimport.load()
    os.makedirs(GOLDEN, exist_ok=True)
    base, tf = ref.base, ref.tf

    # ---------------- decode + NMS, v3 ----------------
    names = ["c%d" % i for i in range(80)]
    rnet = ref.v3.create_network(np.reshape(V3_ANCHORS, [-1, 2]), names, False, input_shape=POST_V3_SHAPE)
    head = post_v3_head()
    thr, iou_thr = 0.5, 0.6
    out = {}
    for i, img in enumerate(head):
        idx, cand = 0, []
        for l in rnet[-1].yolos:
            dim = l.h * l.w * l.b
            l_out = np.reshape(img[idx:idx + dim, ...], [l.h, l.w, l.b, -1])
            boxes = ref.v3._find_bounding_boxes(l_out, l.anchors, thr)
            # recover the row of every box from the loop order: replay the threshold test
            rows = [idx + (cy * l.w + cx) * l.b + b for cy in range(l.h) for cx in range(l.w) for b in range(l.b)
                    if not (base.sigmoid(l_out[cy, cx, b, 4]) < thr)]
            assert len(rows) == len(boxes)
            for bx, r in zip(boxes, rows):
                bx.row = r
            cand.extend(boxes)
            idx += dim
        for k, v in boxes_to_arrays(cand).items():
            out["cand%d_%s" % (i, k)] = v
        kept = base.non_maximum_suppression(list(cand), iou_thr)
        out["kept%d_row" % i] = np.asarray([b.row for b in kept], dtype=np.int64)
    # the public entry point must agree with the pieces above
    res = ref.yolo.YoloV3().find_bounding_boxes(head, rnet, thr, iou_thr, None, names)
    for i, boxes in enumerate(res):
        assert len(boxes) == len(out["kept%d_row" % i])
    out["threshold"], out["iou_threshold"] = np.float64(thr), np.float64(iou_thr)
    np.savez_compressed(os.path.join(GOLDEN, "post_v3.npz"), **out)
    print("post_v3:", [len(out["cand%d_row" % i]) for i in range(2)], [len(out["kept%d_row" % i]) for i in range(2)])

    # ---------------- decode + NMS, v2 ----------------
    head2 = post_v2_head()
    anchors2 = np.reshape(V2_ANCHORS, [-1, 2])
    thr2 = 0.3
    out = {}
    h5 = np.reshape(head2, [-1, 4, 6, 5, 25])
    for i, img in enumerate(h5):
        boxes = ref.v2._find_bounding_boxes(img, anchors2, thr2)
        rows = [(cy * 6 + cx) * 5 + b for cy in range(4) for cx in range(6) for b in range(5)
                if not (base.sigmoid(img[cy, cx, b, 4]) * np.max(base.softmax(img[cy, cx, b, 5:])) < thr2)]
        assert len(rows) == len(boxes)
        for bx, r in zip(boxes, rows):
            bx.row = r
        for k, v in boxes_to_arrays(boxes).items():
            out["cand%d_%s" % (i, k)] = v
        kept = base.non_maximum_suppression(list(boxes), iou_thr)
        out["kept%d_row" % i] = np.asarray([b.row for b in kept], dtype=np.int64)
    res = ref.yolo.YoloV2().find_bounding_boxes(head2, None, thr2, iou_thr, anchors2, ["c"] * 20)
    for i, boxes in enumerate(res):
        assert len(boxes) == len(out["kept%d_row" % i])
    out["threshold"], out["iou_threshold"] = np.float64(thr2), np.float64(iou_thr)
    np.savez_compressed(os.path.join(GOLDEN, "post_v2.npz"), **out)
    print("post_v2:", [len(out["cand%d_row" % i]) for i in range(2)], [len(out["kept%d_row" % i]) for i in range(2)])

    # ---------------- adversarial NMS ----------------
    out = {}
    for ci, case in enumerate(nms_cases()):
        for regime, (xy_t, wh_t) in (("f64", (np.float32, np.float64)), ("f32", (np.float32, np.float32)), ("d64", (np.float64, np.float64))):
            boxes = []
            for j in range(len(case["prob"])):
                b = base.BoundingBox(x=xy_t(case["x"][j]), y=xy_t(case["y"][j]), w=wh_t(case["w"][j]), h=wh_t(case["h"][j]),
                                     prob=np.float32(case["prob"][j]))
                b.row = j
                boxes.append(b)
            with np.errstate(all="ignore"):
                kept = base.non_maximum_suppression(boxes, 0.6)
            out["case%d_%s" % (ci, regime)] = np.asarray([b.row for b in kept], dtype=np.int64)
        for k in ("x", "y", "w", "h", "prob"):
            out["case%d_in_%s" % (ci, k)] = np.asarray(case[k], dtype=np.float64)
    assert len(base.non_maximum_suppression([], 0.6)) == 0
    out["n_cases"] = np.int64(len(nms_cases()))
    np.savez_compressed(os.path.join(GOLDEN, "nms_cases.npz"), **out)
    print("nms_cases:", {k: v.tolist() for k, v in out.items() if k.endswith("_f64") and len(v) < 8})

    # ---------------- conv stacks over the TF stub ----------------
    pnet = pv3.create_network(np.reshape(V3_ANCHORS, [-1, 2]), names, False, input_shape=CONV_V3_SHAPE)
    stream = synth.weight_stream(pnet[0]._yb_state.graph.specs, seed=2, num_classes=80)
    rnet = ref.v3.create_network(np.reshape(V3_ANCHORS, [-1, 2]), names, False, input_shape=CONV_V3_SHAPE)
    tf.run(base.load_weights(rnet, stream))
    x = synth.images(2, CONV_V3_SHAPE[0], CONV_V3_SHAPE[1], seed=1)
    y = tf.run(rnet[-1].out, {rnet[0].out: x})
    np.savez_compressed(os.path.join(GOLDEN, "conv_v3.npz"), net_out=y.astype(np.float32),
                        variable_names=np.asarray(sum([l.variable_names for l in rnet], [])))
    print("conv_v3:", y.shape, float(np.abs(y).max()))

    names20 = ["c%d" % i for i in range(20)]
    pnet = pv2.create_full_network(anchors2, names20, False, input_shape=CONV_V2_SHAPE)
    stream = synth.weight_stream(pnet[0]._yb_state.graph.specs, seed=3, num_classes=20)
    rnet = ref.v2.create_full_network(anchors2, names20, False, input_shape=CONV_V2_SHAPE)
    tf.run(base.load_weights(rnet, stream))
    x = synth.images(2, CONV_V2_SHAPE[0], CONV_V2_SHAPE[1], seed=4)
    y = tf.run(rnet[-1].out, {rnet[0].out: x})
    np.savez_compressed(os.path.join(GOLDEN, "conv_v2.npz"), net_out=y.astype(np.float32),
                        variable_names=np.asarray(sum([l.variable_names for l in rnet], [])))
    print("conv_v2:", y.shape, float(np.abs(y).max()))


if __name__ == "__main__":
    main()
