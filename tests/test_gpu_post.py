"""GPU: decode + NMS kernels through the C ABI against the oracle and the reference-made goldens.
Bars: decode values rtol 1e-6 (expf vs numpy exp), identical candidate rows away from the threshold,
classes identical; NMS kept lists BIT-IDENTICAL (order included)."""
import numpy as np
import pytest

import helpers
from oracle import convstack, make_golden, postprocess
from tensorflow_yolo_b200 import engine, synth
from tensorflow_yolo_b200.net import base as pbase, v2 as pv2, v3 as pv3

pytestmark = pytest.mark.gpu


def _geo(shape):
    topo = convstack.topology_v3(80, np.reshape(helpers.V3_ANCHORS, [-1, 2]), shape)
    return convstack.yolo_geometry(topo, shape)


def _post_v3(shape, max_batch):
    return engine.PostProcessor([(h, w, a) for h, w, b, a in _geo(shape)], 80, engine.YB_DECODE_V3, max_batch=max_batch)


def _assert_decode_close(c, ref):
    assert np.array_equal(c["row"], ref["row"])
    for k in ("x", "y", "w", "h", "prob"):
        np.testing.assert_allclose(c[k], ref[k], rtol=1e-6, atol=0)
    assert np.array_equal(c["class_idx"], ref["class_idx"])


def test_decode_v3_golden():
    g = helpers.golden("post_v3.npz")
    post = _post_v3(make_golden.POST_V3_SHAPE, 2)
    cands = post.decode(make_golden.post_v3_head(), float(g["threshold"]))
    for i, d in enumerate(cands):
        ref = {k: g["cand%d_%s" % (i, k)] for k in ("x", "y", "w", "h", "prob", "class_idx", "row")}
        _assert_decode_close(helpers.cand_from_dets(d), ref)


def test_decode_v2_golden():
    g = helpers.golden("post_v2.npz")
    anchors = [tuple(a) for a in np.reshape(make_golden.V2_ANCHORS, [-1, 2])]
    post = engine.PostProcessor([(4, 6, anchors)], 20, engine.YB_DECODE_V2, max_batch=2)
    cands = post.decode(make_golden.post_v2_head(), float(g["threshold"]))
    for i, d in enumerate(cands):
        ref = {k: g["cand%d_%s" % (i, k)] for k in ("x", "y", "w", "h", "prob", "class_idx", "row")}
        c = helpers.cand_from_dets(d)
        assert np.array_equal(c["row"], ref["row"])
        for k in ("x", "y", "w", "h"):
            np.testing.assert_allclose(c[k], ref[k], rtol=1e-6)
        np.testing.assert_allclose(c["prob"], ref["prob"], rtol=2e-6)      # softmax sum order differs
        assert np.array_equal(c["class_idx"], ref["class_idx"])


def test_nms_on_reference_decoded_boxes_bit_identical():
    """north_star: identical index sets when fed the reference's own decoded boxes and scores."""
    for name in ("post_v3.npz", "post_v2.npz"):
        g = helpers.golden(name)
        for i in range(2):
            k = engine.nms(g["cand%d_x" % i], g["cand%d_y" % i], g["cand%d_w" % i], g["cand%d_h" % i],
                           g["cand%d_prob" % i], float(g["iou_threshold"]))
            assert np.array_equal(g["cand%d_row" % i][k], g["kept%d_row" % i])


def test_nms_adversarial_cases_bit_identical():
    g = helpers.golden("nms_cases.npz")
    for ci in range(int(g["n_cases"])):
        for regime, (xy_t, wh_t) in (("f64", (np.float32, np.float64)), ("f32", (np.float32, np.float32)),
                                      ("d64", (np.float64, np.float64))):
            k = engine.nms(g["case%d_in_x" % ci].astype(xy_t), g["case%d_in_y" % ci].astype(xy_t),
                           g["case%d_in_w" % ci].astype(wh_t), g["case%d_in_h" % ci].astype(wh_t),
                           g["case%d_in_prob" % ci].astype(np.float32), 0.6)
            assert np.array_equal(k, g["case%d_%s" % (ci, regime)]), (ci, regime)
    z = np.zeros(0, np.float32)
    assert len(engine.nms(z, z, z, z, z, 0.6)) == 0
    assert pbase.non_maximum_suppression([], 0.6) == []


@pytest.mark.parametrize("k,seed", [(1, 0), (63, 1), (64, 2), (65, 3), (1000, 4), (5000, 5)])
def test_nms_random_sets_vs_oracle(k, seed):
    rs = np.random.RandomState(seed)
    c = {"x": rs.uniform(0, 1, k).astype(np.float32), "y": rs.uniform(0, 1, k).astype(np.float32),
         "w": rs.uniform(.02, .5, k), "h": rs.uniform(.02, .5, k),
         "prob": (np.round(rs.uniform(.3, 1, k) * 50) / 50).astype(np.float32)}     # many exact score ties
    for thr in (0.3, 0.6):
        got = engine.nms(c["x"], c["y"], c["w"], c["h"], c["prob"], thr)
        assert np.array_equal(got, postprocess.nms(c, thr))


def test_nms_via_reference_named_function():
    rs = np.random.RandomState(9)
    boxes = [pbase.BoundingBox(x=np.float32(rs.uniform()), y=np.float32(rs.uniform()), w=np.float64(rs.uniform(.1, .5)),
                               h=np.float64(rs.uniform(.1, .5)), class_idx=int(rs.randint(3)), prob=np.float32(rs.uniform()))
             for _ in range(200)]
    c = {k: np.asarray([getattr(b, k) for b in boxes]) for k in ("x", "y", "w", "h", "prob")}
    kept = pbase.non_maximum_suppression(list(boxes), 0.5)
    assert [boxes.index(b) for b in kept] == postprocess.nms(c, 0.5).tolist()


def test_find_bounding_boxes_signature_v3_and_v2():
    """The reference's module functions, same arguments, GPU underneath."""
    shape = (64, 64, 3)
    net = pv3.create_network(np.reshape(helpers.V3_ANCHORS, [-1, 2]), helpers.names(80), False, input_shape=shape)
    head = make_golden.post_v3_head(n=3, seed=31)
    res = pv3.find_bounding_boxes(head, net, 0.5, 0.6, None, helpers.names(80))
    ref = postprocess.find_bounding_boxes_v3(head, _geo(shape), 0.5, 0.6)
    assert len(res) == 3
    for boxes, r in zip(res, ref):
        assert len(boxes) == len(r["row"])
        assert type(boxes[0].x) is np.float32 and type(boxes[0].w) is np.float64 and type(boxes[0].class_idx) is np.int64
        np.testing.assert_allclose([b.prob for b in boxes], r["prob"], rtol=1e-6)
        np.testing.assert_allclose([b.w for b in boxes], r["w"], rtol=1e-6)
        assert [int(b.class_idx) for b in boxes] == r["class_idx"].tolist()
    anchors = np.reshape(helpers.V2_ANCHORS_VOC, [-1, 2])
    net2 = pv2.create_full_network(anchors, helpers.names(20), False, input_shape=(128, 192, 3))
    head2 = make_golden.post_v2_head(n=2, seed=32)
    res2 = pv2.find_bounding_boxes(head2, net2, 0.3, 0.6, anchors, helpers.names(20))
    ref2 = postprocess.find_bounding_boxes_v2(head2, anchors, 20, 0.3, 0.6)
    for boxes, r in zip(res2, ref2):
        assert len(boxes) == len(r["row"])
        np.testing.assert_allclose([b.prob for b in boxes], r["prob"], rtol=2e-6)


def test_pipeline_nms_bit_exact_on_gpu_candidates_416():
    """Full 416 geometry, sparse and denser thresholds: the kept list must equal the oracle NMS run on the
    GPU's own decoded candidates (isolates NMS from the 1-ulp decode differences)."""
    shape = (416, 416, 3)
    post = _post_v3(shape, 2)
    head = synth.head_tensor(2, 10647, 85, seed=3, obj_shift=-2.0)
    for thr in (0.5, 0.05):
        kept = post.run(head, thr, 0.6)
        cand = post.decode(head, thr)
        ref_dec = postprocess.decode_v3_image(head[0], _geo(shape), thr)
        # candidate sets agree except inside a 1e-6 band around the threshold
        gpu_rows, ref_rows = set(cand[0]["row"].tolist()), set(ref_dec["row"].tolist())
        for r in gpu_rows ^ ref_rows:
            assert abs(1.0 / (1.0 + np.exp(-np.float64(head[0, r, 4]))) - thr) < 1e-6
        for i in range(2):
            c = helpers.cand_from_dets(cand[i])
            k = postprocess.nms(c, 0.6)
            assert np.array_equal(c["row"][k], kept[i]["row"])
            assert np.all(np.diff(kept[i]["prob"]) <= 0)                     # kept order = score descending


def test_dense_config5_properties_full_rows():
    """BASELINE config 5 geometry (10647 x 85, thr 0.001, every row a candidate).  One image is checked
    bit-for-bit against the oracle; a batch through size-independent properties: kept sorted by score,
    idempotence (NMS of the kept set keeps everything), no kept pair overlaps >= thr on a sample,
    per-image independence (same image twice -> same result)."""
    shape = (416, 416, 3)
    n = 16
    head = synth.head_tensor(n, 10647, 85, seed=0)
    head[n - 1] = head[0]
    post = _post_v3(shape, n)
    kept = post.run(head, 0.001, 0.6)
    assert np.all(post.last_candidates == 10647)
    c0 = helpers.cand_from_dets(post.decode(head[:1], 0.001)[0])
    assert np.array_equal(c0["row"][postprocess.nms(c0, 0.6)], kept[0]["row"])
    assert np.array_equal(kept[0]["row"], kept[n - 1]["row"])
    for i in (1, n // 2):
        d = kept[i]
        assert np.all(np.diff(d["prob"]) <= 0)
        again = engine.nms(d["x"], d["y"], d["w"], d["h"], d["prob"], 0.6)
        assert np.array_equal(again, np.arange(len(d)))
        c = helpers.cand_from_dets(d)
        for j in range(0, min(len(d), 200)):
            iou = postprocess.iou_matrix_row(c, j, np.arange(j + 1, len(d)))
            assert not np.any(iou >= 0.6)


def test_608_rows_use_global_sort_scratch():
    """22743 rows per image (608 input) exceed the 16384-key shared-memory sort and take the global-memory path."""
    shape = (608, 608, 3)
    post = _post_v3(shape, 1)
    head = synth.head_tensor(1, 22743, 85, seed=6)
    kept = post.run(head, 0.001, 0.6)
    c = helpers.cand_from_dets(post.decode(head, 0.001)[0])
    assert len(c["row"]) == 22743
    assert np.array_equal(c["row"][postprocess.nms(c, 0.6)], kept[0]["row"])


def test_per_class_mode():
    """Optional north-star mode (not in the reference): per-class, suppress iff IoU > thr."""
    rs = np.random.RandomState(2)
    k = 400
    c = {"x": rs.uniform(0, 1, k).astype(np.float32), "y": rs.uniform(0, 1, k).astype(np.float32),
         "w": rs.uniform(.05, .5, k), "h": rs.uniform(.05, .5, k), "prob": rs.uniform(.3, 1, k).astype(np.float32)}
    cls = rs.randint(0, 4, k).astype(np.int32)
    got = engine.nms(c["x"], c["y"], c["w"], c["h"], c["prob"], 0.45, class_idx=cls, nms_mode=engine.YB_NMS_PER_CLASS)
    order = np.argsort(-c["prob"].astype(np.float64), kind="stable")
    kept = []
    for i in order:
        ok = True
        for j in kept:
            if cls[j] == cls[i] and postprocess.iou_matrix_row(c, j, np.asarray([i]))[0] > 0.45:
                ok = False
                break
        if ok:
            kept.append(i)
    assert got.tolist() == kept


def test_device_resident_head_and_batch_limits():
    import torch
    shape = (64, 64, 3)
    post = _post_v3(shape, 2)
    head = make_golden.post_v3_head()
    a = post.run(head, 0.5, 0.6)
    b = post.run(torch.from_numpy(head).cuda(), 0.5, 0.6)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    small = post.run(head, 0.5, 0.6, max_per_image=7)             # truncation keeps the row stride of the caller
    for x, y in zip(a, small):
        assert np.array_equal(x[:7], y)
    with pytest.raises(Exception):
        post.run(np.concatenate([head, head]), 0.5, 0.6)          # batch 4 > max_batch 2


def test_no_candidates_and_single_candidate():
    """Empty inputs of the domain: an image without any candidate (the reference returns [] for it, net/base.py:196-197)
    next to images with one and with many candidates, in one batch."""
    topo = convstack.topology_v3(80, np.reshape(helpers.V3_ANCHORS, [-1, 2]), (416, 416, 3))
    geo = convstack.yolo_geometry(topo, (416, 416, 3))
    rows = sum(h * w * b for h, w, b, a in geo)
    head = np.random.RandomState(5).standard_normal((3, rows, 85)).astype(np.float32)
    head[0, :, 4] = -20.0                    # no row reaches the threshold
    head[1, :, 4] = -20.0
    head[1, 1234, 4] = 3.0                   # exactly one candidate
    post = engine.PostProcessor([(h, w, a) for h, w, b, a in geo], 80, engine.YB_DECODE_V3, max_batch=3)
    kept = post.run(head, 0.5, 0.6)
    assert len(kept[0]) == 0 and post.last_candidates[0] == 0
    assert len(kept[1]) == 1 and kept[1][0]["row"] == 1234 and post.last_candidates[1] == 1
    ref = postprocess.find_bounding_boxes_v3(head[:2], geo, 0.5, 0.6)
    for i in range(2):
        assert np.array_equal(kept[i]["row"], ref[i]["row"])
    # the dense image: the candidate set is exact (sigmoid(t) >= 0.5 <=> t >= 0); its kept list is covered by the
    # bit-identical NMS tests above, which start from the reference's own decoded boxes
    assert post.last_candidates[2] == int((head[2, :, 4] >= 0).sum()) and 0 < len(kept[2]) <= post.last_candidates[2]
    # a threshold nothing can reach empties the whole batch
    assert all(len(k) == 0 for k in post.run(head, 1.5, 0.6))


def _clustered_boxes(rs, k, irregular):
    """Jittered copies of a few base boxes (IoU values all over (0, 1), many near any threshold) with log-normal sizes;
    optionally mixed with boxes the shrunk-corner prefilter must hand to the exact path: zero / negative extents,
    extents vanishing against the coordinates, huge boxes."""
    nb = max(1, k // 40)
    bx, by = rs.uniform(0, 1, nb), rs.uniform(0, 1, nb)
    bw, bh = 0.05 * np.exp(rs.normal(0, 1, nb)), 0.05 * np.exp(rs.normal(0, 1, nb))
    pick = rs.randint(nb, size=k)
    jit = rs.choice([0.02, 0.1, 0.4], size=k)
    x = (bx[pick] + bw[pick] * jit * rs.normal(0, 1, k)).astype(np.float32)
    y = (by[pick] + bh[pick] * jit * rs.normal(0, 1, k)).astype(np.float32)
    w = bw[pick] * np.exp(jit * rs.normal(0, 1, k))
    h = bh[pick] * np.exp(jit * rs.normal(0, 1, k))
    if irregular:
        idx = rs.permutation(k)
        n = max(1, k // 25)
        w[idx[:n]] = 0.0
        h[idx[n:2 * n]] = -h[idx[n:2 * n]]
        w[idx[2 * n:3 * n]] = -w[idx[2 * n:3 * n]]; h[idx[2 * n:3 * n]] = -h[idx[2 * n:3 * n]]
        x[idx[3 * n:4 * n]] += np.float32(4096.0); w[idx[3 * n:4 * n]] = 1e-3      # W / |x| < 1e-4
        w[idx[4 * n:5 * n]] = 1e6; h[idx[4 * n:5 * n]] = 1e6
        w[idx[5 * n:6 * n]] = 1e-200; h[idx[5 * n:6 * n]] = 1e-200                # area underflows to 0
    return {"x": x, "y": y, "w": w, "h": h, "prob": rs.uniform(.01, 1, k).astype(np.float32)}


@pytest.mark.parametrize("k,seed,irregular", [(700, 0, False), (700, 1, True), (3000, 2, False), (3000, 3, True)])
def test_nms_near_threshold_clusters_and_irregular_boxes(k, seed, irregular):
    """The prefilter (float corners shrunk by thr * extent) may only drop pairs the exact float64 rule rejects:
    kept lists stay bit-identical for thresholds below / at / above its switch-on point (1e-2) and beyond 1."""
    c = _clustered_boxes(np.random.RandomState(100 + seed), k, irregular)
    for thr in (0.003, 0.01, 0.0101, 0.3, 0.6, 0.9, 1.0, 1.25):
        got = engine.nms(c["x"], c["y"], c["w"], c["h"], c["prob"], thr)
        assert np.array_equal(got, postprocess.nms(c, thr)), thr


def test_nms_dense_config5_image_vs_oracle():
    """One image of BASELINE config 5 (10,647 candidates, N(0,1) logits, thresholds 0.001 / 0.6)."""
    shape = (416, 416, 3)
    post = _post_v3(shape, 1)
    head = np.random.RandomState(0).standard_normal((1, 10647, 85)).astype(np.float32)
    kept = post.run(head, 0.001, 0.6)
    cand = post.decode(head, 0.001)
    c = helpers.cand_from_dets(cand[0])
    assert len(c["row"]) == 10647
    assert np.array_equal(c["row"][postprocess.nms(c, 0.6)], kept[0]["row"])


@pytest.mark.parametrize("bulk", ["0", "1"])
def test_decode_class_argmax_near_ties(bulk, monkeypatch):
    monkeypatch.setenv("YB_DECODE_BULK", bulk)        # 1: the dense kernel (rows staged in shared memory, argmax on raw logits)
    _near_ties_case()


def test_bulk_decode_equals_row_kernel_bit_for_bit(monkeypatch):
    """The two v3 decode kernels (sparse: one sector per row + cooperative candidates; bulk: dense heads, config 5) must
    produce identical candidates -- every field bit for bit -- on random heads, with NaN / inf logits, at a row count
    that is not a multiple of the 32-row slabs, for thresholds on both sides of the automatic switch."""
    shape = (96, 96, 3)                                   # R = 3 * (9 + 36 + 144) = 567 rows per image
    rs = np.random.RandomState(5)
    head = rs.normal(0, 2.0, (3, 567, 85)).astype(np.float32)
    head[0, 10, 5:] = np.nan
    head[0, 11, 17] = np.nan
    head[1, 3, 5:9] = np.inf
    head[1, 4, 5:] = -np.inf
    head[2, 5, 40] = 7.5
    head[2, 6, 5:] = -85.0
    post = _post_v3(shape, 3)
    for thr in (0.001, 0.3):
        out = {}
        for bulk in ("0", "1"):
            monkeypatch.setenv("YB_DECODE_BULK", bulk)
            out[bulk] = post.decode(head, thr)
        for i, (a, b) in enumerate(zip(out["0"], out["1"])):
            assert len(a) == len(b) and len(a) > 100
            for f in ("row", "class_idx", "prob", "x", "y", "w", "h"):
                assert np.array_equal(a[f], b[f], equal_nan=True), f
            with np.errstate(all="ignore"):              # both against numpy's argmax (first NaN / first of equal maxima)
                ref = postprocess.decode_v3_image(head[i], _geo(shape), thr)
            assert np.array_equal(a["row"], ref["row"]) and np.array_equal(a["class_idx"], ref["class_idx"])
        monkeypatch.delenv("YB_DECODE_BULK")
        auto = post.decode(head, thr)
        for a, b in zip(out["0"], auto):
            assert np.array_equal(a["class_idx"], b["class_idx"]) and np.array_equal(a["row"], b["row"])


def _near_ties_case():
    """np.argmax(sigmoid(t)) with classes a few float32 ulps apart, equal, saturated (sigmoid == 1 for many classes) and
    tiny (subnormal sigmoids): the device ranks classes on 1 + exp(-t) with the fast exponential and has to fall back to the
    exact quotients whenever that ranking cannot be trusted; the first of the maxima wins, as in numpy."""
    shape = (64, 64, 3)
    post = _post_v3(shape, 2)
    rs = np.random.RandomState(11)
    R = 252
    head = np.zeros((2, R, 85), np.float32)
    head[..., :4] = rs.normal(0, 1, (2, R, 4))
    head[..., 4] = 3.0                                               # every row is a candidate
    base = rs.choice([-95.0, -60.0, -20.0, -3.0, 0.0, 2.5, 9.0, 17.0, 40.0, 90.0], size=(2, R, 1)).astype(np.float32)
    cls = np.broadcast_to(base, (2, R, 80)).copy()
    kind = rs.randint(0, 4, size=(2, R))
    for i in range(2):
        for r in range(R):
            row = cls[i, r]
            if kind[i, r] == 0:                                      # a handful of classes within a few ulps of the top
                top = rs.choice(80, 5, replace=False)
                row -= np.float32(abs(row[0]) * 1e-3 + 1e-3)
                for j, k in enumerate(top):
                    v = np.float32(base[i, r, 0])
                    for _ in range(int(rs.randint(0, 4))):
                        v = np.nextafter(v, np.float32(-np.inf))
                    row[k] = v
            elif kind[i, r] == 1:                                    # all equal: class 0
                pass
            elif kind[i, r] == 2:                                    # relative differences around the tie band
                row += (rs.choice([1e-7, 1e-6, 1e-5, 1e-4, 1e-3], 80) * rs.normal(0, 1, 80) * max(1.0, abs(float(row[0])))).astype(np.float32)
            else:                                                    # ordinary logits
                row[:] = rs.normal(0, 3, 80)
    head[..., 5:] = cls
    cands = post.decode(head, 0.5)
    for i in range(2):
        ref = postprocess.decode_v3_image(head[i], _geo(shape), 0.5)
        c = helpers.cand_from_dets(cands[i])
        assert np.array_equal(c["row"], ref["row"]) and len(c["row"]) == R
        bad = np.nonzero(c["class_idx"] != ref["class_idx"])[0]
        assert bad.size == 0, [(int(r), int(kind[i, r]), float(base[i, r, 0]), int(c["class_idx"][r]), int(ref["class_idx"][r]))
                               for r in bad[:8]]
