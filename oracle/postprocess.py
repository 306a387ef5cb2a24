"""CPU restatement of the reference's head decode and greedy NMS (numpy) -- TEST INFRASTRUCTURE.

Pinned against the reference's own functions executed unmodified in the build container
(tests/golden/postprocess_*.npz made by oracle/make_golden.py; tests/test_oracle_ref.py re-runs
the comparison whenever /root/reference is present).

dtype regime = the reference's code under numpy >= 2 (NEP 50), probed in the build container:
  x, y, prob   numpy.float32   (base.sigmoid on float32 scalars stays float32, net/base.py:171-172)
  w, h         numpy.float64   (anchors are numpy.float64 scalars: net/layers.py:130-131 for v3,
                                np.reshape(params["anchors"]) for v2 -> float64 * float32 = float64)
  class_idx    numpy.int64     (np.argmax)
  IoU          float64         (x -+ w/2 promotes to float64, net/base.py:180-192, :266-272)
  thresholds   `p < threshold`: p float32, threshold python float -> compared in float32;
               `iou >= iou_threshold`: float64 compare.
"""
import numpy as np


def _sigmoid(x):                                    # net/base.py:171-172
    return 1. / (1. + np.exp(-x))


def decode_v3_scale(out, anchors, threshold):
    """Vectorised net/v3.py:109-136.  ``out``: [h,w,b,5+C] float32; ``anchors``: b pairs (grid units).

    Returns dict of arrays in the reference's raster order (cy, cx, b):
    x,y float32; w,h float64; prob float32; class_idx int64; row int64 (= (cy*w+cx)*b + a).
    """
    out = np.asarray(out, dtype=np.float32)
    h, w, nb = out.shape[0:3]
    prob_obj = _sigmoid(out[..., 4])                                   # v3.py:119
    prob_classes = _sigmoid(out[..., 5:])                              # :120
    class_idx = np.argmax(prob_classes, axis=-1)                       # :121 (first max)
    keep = ~(prob_obj < np.float32(threshold))                         # :124 (float32 compare)
    cy, cx, b = np.nonzero(keep)                                       # C order == loop order :115-117
    t = out[cy, cx, b, :]
    aw = np.asarray([a[0] for a in anchors], dtype=np.float64)[b]
    ah = np.asarray([a[1] for a in anchors], dtype=np.float64)[b]
    x = (_sigmoid(t[:, 0]) + cx.astype(np.float32)) / np.float32(w)    # :129  (f32 + int -> f32) / int
    y = (_sigmoid(t[:, 1]) + cy.astype(np.float32)) / np.float32(h)    # :130
    bw = (aw * np.exp(t[:, 2]).astype(np.float64)) / w                 # :131  float64
    bh = (ah * np.exp(t[:, 3]).astype(np.float64)) / h                 # :132
    return {"x": x.astype(np.float32), "y": y.astype(np.float32), "w": bw, "h": bh,
            "prob": prob_obj[cy, cx, b].astype(np.float32),
            "class_idx": class_idx[cy, cx, b].astype(np.int64),
            "row": ((cy * w + cx) * nb + b).astype(np.int64)}


def _softmax_rows(x):
    """net/base.py:175-177 applied per row.  The sum uses numpy's own float32 reduction on each
    contiguous row so that its pairwise order matches the reference's ``e_x.sum()``."""
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    s = np.empty(e.shape[:-1], dtype=e.dtype)
    flat_e, flat_s = e.reshape(-1, e.shape[-1]), s.reshape(-1)
    for i in range(flat_e.shape[0]):
        flat_s[i] = flat_e[i].sum()
    return e / s[..., None]


def decode_v2(out, anchors, threshold):
    """Vectorised net/v2.py:93-119.  ``out``: [h,w,b,5+C] float32; anchors in grid units."""
    out = np.asarray(out, dtype=np.float32)
    h, w, nb = out.shape[0:3]
    prob_obj = _sigmoid(out[..., 4])                                   # v2.py:102
    prob_classes = _softmax_rows(out[..., 5:])                         # :103
    class_idx = np.argmax(prob_classes, axis=-1)                       # :104
    class_prob = np.take_along_axis(prob_classes, class_idx[..., None], axis=-1)[..., 0]
    p = prob_obj * class_prob                                          # :106
    keep = ~(p < np.float32(threshold))                                # :107
    cy, cx, b = np.nonzero(keep)
    t = out[cy, cx, b, :]
    anchors = np.asarray(anchors, dtype=np.float64).reshape(-1, 2)
    aw, ah = anchors[b, 0], anchors[b, 1]
    x = (_sigmoid(t[:, 0]) + cx.astype(np.float32)) / np.float32(w)
    y = (_sigmoid(t[:, 1]) + cy.astype(np.float32)) / np.float32(h)
    bw = (aw * np.exp(t[:, 2]).astype(np.float64)) / w
    bh = (ah * np.exp(t[:, 3]).astype(np.float64)) / h
    return {"x": x.astype(np.float32), "y": y.astype(np.float32), "w": bw, "h": bh,
            "prob": p[cy, cx, b].astype(np.float32),
            "class_idx": class_idx[cy, cx, b].astype(np.int64),
            "row": ((cy * w + cx) * nb + b).astype(np.int64)}


def _concat(parts):
    keys = ["x", "y", "w", "h", "prob", "class_idx", "row"]
    if not parts:
        return {"x": np.zeros(0, np.float32), "y": np.zeros(0, np.float32), "w": np.zeros(0, np.float64),
                "h": np.zeros(0, np.float64), "prob": np.zeros(0, np.float32),
                "class_idx": np.zeros(0, np.int64), "row": np.zeros(0, np.int64)}
    return {k: np.concatenate([p[k] for p in parts]) for k in keys}


def decode_v3_image(rows, yolo_geometry, threshold):
    """net/v3.py:139-149 for one image.  ``rows``: [R,5+C]; ``yolo_geometry``: [(h,w,b,anchors)]
    in detection order.  ``row`` in the result is the global row index in [0,R)."""
    parts, idx = [], 0
    for (h, w, b, anchors) in yolo_geometry:
        dim = h * w * b
        l_out = np.reshape(rows[idx:idx + dim, ...], [h, w, b, -1])
        d = decode_v3_scale(l_out, anchors, threshold)
        d["row"] = d["row"] + idx
        parts.append(d)
        idx += dim
    return _concat(parts)


def iou_matrix_row(c, i, js):
    """net/base.py:180-192 between box ``i`` and boxes ``js`` of candidate dict ``c`` (same op order)."""
    x, y, w, h = c["x"], c["y"], c["w"], c["h"]
    two = 2.
    min1 = (x[i] - w[i] / two, y[i] - h[i] / two)
    max1 = (x[i] + w[i] / two, y[i] + h[i] / two)
    area1 = w[i] * h[i]
    min2 = (x[js] - w[js] / two, y[js] - h[js] / two)
    max2 = (x[js] + w[js] / two, y[js] + h[js] / two)
    area2 = w[js] * h[js]
    iw = np.maximum(np.minimum(max1[0], max2[0]) - np.maximum(min1[0], min2[0]), 0)
    ih = np.maximum(np.minimum(max1[1], max2[1]) - np.maximum(min1[1], min2[1]), 0)
    inter = iw * ih
    dt = inter.dtype.type
    union = np.maximum(area1 + area2 - inter, dt(1e-8))
    return inter / union


def nms(c, iou_threshold):
    """net/base.py:195-209: stable sort by prob descending, greedy, class-agnostic, suppress iff
    IoU >= threshold.  ``c``: candidate dict (arrays).  Returns indices into ``c`` of the kept boxes,
    in kept (score-descending) order.

    Equivalent reformulation of the reference loop: walking the sorted list, a box is kept iff no
    previously *kept* box overlaps it; so each newly kept box suppresses all later boxes at once.
    """
    n = len(c["prob"])
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    # list.sort(key, reverse=True) is stable: equal keys keep their original relative order
    order = np.argsort(-c["prob"].astype(np.float64), kind="stable")
    s = {k: np.asarray(c[k])[order] for k in ("x", "y", "w", "h")}
    # thr compare in the IoU's dtype (python-float thr is weak under NEP 50)
    iou_dt = np.result_type(s["x"].dtype, s["w"].dtype)
    thr = iou_dt.type(iou_threshold)
    removed = np.zeros(n, dtype=bool)
    kept = []
    for i in range(n):
        if removed[i]:
            continue
        kept.append(i)
        if i + 1 < n:
            js = np.arange(i + 1, n)
            js = js[~removed[i + 1:]]
            if js.size:
                iou = iou_matrix_row(s, i, js)
                removed[js[iou >= thr]] = True
    return order[np.asarray(kept, dtype=np.int64)]


def find_bounding_boxes_v3(net_out, yolo_geometry, threshold, iou_threshold):
    """net/v3.py:139-151 -> per image (candidates dict restricted to kept boxes, kept order)."""
    results = []
    for out in net_out:
        c = decode_v3_image(out, yolo_geometry, threshold)
        k = nms(c, iou_threshold)
        results.append({key: val[k] for key, val in c.items()})
    return results


def find_bounding_boxes_v2(net_out, anchors, num_classes, threshold, iou_threshold):
    """net/v2.py:82-90."""
    anchors = np.asarray(anchors, dtype=np.float64).reshape(-1, 2)
    net_out = np.reshape(net_out, [-1, net_out.shape[1], net_out.shape[2], len(anchors), 5 + num_classes])
    results = []
    for out in net_out:
        c = decode_v2(out, anchors, threshold)
        k = nms(c, iou_threshold)
        results.append({key: val[k] for key, val in c.items()})
    return results
