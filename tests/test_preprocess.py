"""Preprocessing.  CPU: the oracle's restatement of cv2.resize(INTER_LINEAR) + BGR->RGB is pinned
bit for bit against cv2 itself and against the reference's own preprocess_image.  GPU: the CUDA kernel behind
yb_resize_bgr2rgb / yb_engine_forward_raw against the oracle (and cv2) on the same images."""
import os

import numpy as np
import pytest

import helpers
from oracle import preprocess as opre

cv2 = pytest.importorskip("cv2")

SIZES = [(90, 120, 416, 416), (100, 120, 416, 416), (480, 640, 416, 416), (832, 832, 416, 416), (416, 416, 416, 416),
         (1080, 1920, 416, 416), (375, 500, 608, 608), (50, 37, 416, 416), (1, 1, 8, 8), (2, 3, 7, 5), (833, 831, 416, 416),
         (600, 600, 300, 300), (600, 601, 300, 300), (7, 1000, 64, 96), (1000, 7, 96, 64)]


def _images(seed, sizes):
    rs = np.random.RandomState(seed)
    return [rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w, _, _ in sizes]


def test_oracle_resize_is_bit_exact_with_cv2():
    rs = np.random.RandomState(1)
    sizes = SIZES + [(rs.randint(1, 500), rs.randint(1, 500), rs.randint(1, 450), rs.randint(1, 450)) for _ in range(25)]
    for img, (sh, sw, dh, dw) in zip(_images(0, sizes), sizes):
        assert np.array_equal(opre.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh))), (sh, sw, dh, dw)
    # smooth images (neighbouring pixels correlated) exercise the rounding differently from noise
    yy, xx = np.mgrid[0:333, 0:517]
    smooth = np.stack([(xx * 255 // 516), (yy * 255 // 332), ((xx + yy) % 256)], 2).astype(np.uint8)
    assert np.array_equal(opre.resize_linear_u8(smooth, 416, 416), cv2.resize(smooth, (416, 416)))


def test_oracle_matches_the_reference_preprocess_image(tmp_path):
    """The unmodified reference function (net/base.py:115-155) on real files, including its dsize quirk."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip("/root/reference is not present")
    ref_base = refimport.load().base
    rs = np.random.RandomState(5)
    for i, (h, w, shape) in enumerate([(90, 120, (416, 416, 3)), (300, 200, (96, 96, 3)), (64, 80, (64, 32, 3))]):
        path = str(tmp_path / ("im%d.png" % i))
        img = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        cv2.imwrite(path, img)
        want, _ = ref_base.preprocess_image(path, shape)
        got = opre.preprocess_image(cv2.imread(path), shape)
        assert want.dtype == got.dtype == np.float64 and want.shape == got.shape == (shape[1], shape[0], 3)
        assert np.array_equal(want, got)


def test_oracle_matches_committed_reference_fixture():
    """tests/golden/preprocess.npz: outputs of the reference's preprocess_image made by oracle/make_golden_pre.py."""
    g = helpers.golden("preprocess.npz")
    for i in range(int(g["n_cases"])):
        shape = tuple(int(v) for v in g["shape%d" % i])
        got = opre.preprocess_u8(g["src%d" % i], shape)
        assert got.shape == (shape[1], shape[0], 3) and np.array_equal(got, g["rgb%d" % i]), i
        assert np.array_equal(opre.preprocess_image(g["src%d" % i], shape), g["rgb%d" % i] / 255.)


@pytest.mark.gpu
def test_gpu_resize_matches_committed_reference_fixture():
    from tensorflow_yolo_b200 import engine
    g = helpers.golden("preprocess.npz")
    for i in range(int(g["n_cases"])):
        shape = tuple(int(v) for v in g["shape%d" % i])
        got = engine.resize_bgr2rgb([g["src%d" % i]], shape[1], shape[0])[0]        # dsize = (input_h, input_w): rows = input_w
        assert np.array_equal(got, g["rgb%d" % i]), i


@pytest.mark.gpu
def test_gpu_resize_bit_exact_any_shape():
    from tensorflow_yolo_b200 import engine
    imgs = _images(2, SIZES)
    # one call per destination shape; mixed source sizes inside a call share nothing but the kernel
    by_dst = {}
    for img, (sh, sw, dh, dw) in zip(imgs, SIZES):
        by_dst.setdefault((dh, dw), []).append(img)
    for (dh, dw), group in by_dst.items():
        got = engine.resize_bgr2rgb(group, dh, dw)
        for g, img in zip(got, group):
            want = opre.resize_linear_u8(img, dw, dh)[:, :, ::-1]
            assert np.array_equal(g, want), (img.shape, dh, dw)
            assert np.array_equal(g, cv2.resize(img, (dw, dh))[:, :, ::-1])
    # strided views (a crop of a bigger image) are taken as they are
    big = _images(3, [(200, 300, 0, 0)])[0]
    crop = big[10:150, 20:220]
    assert np.array_equal(engine.resize_bgr2rgb([crop], 64, 64)[0], cv2.resize(np.ascontiguousarray(crop), (64, 64))[:, :, ::-1])


@pytest.mark.gpu
def test_forward_raw_equals_host_preprocessing_bit_for_bit(tmp_path):
    """Feeding decoded images (device resize) gives the very same network output as the reference protocol: host
    cv2 preprocessing -> float64/255 -> float32 feed."""
    from tensorflow_yolo_b200 import engine
    from tensorflow_yolo_b200.net import base as pbase
    shape = (96, 96, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2, obj_bias=-1.0)
    eng = engine.Engine(net[0]._yb_state.plan(), shape, 80, engine.YB_DECODE_V3, max_batch=4)
    eng.load_weights(stream)
    rs = np.random.RandomState(9)
    raw, paths = [], []
    for i, (h, w) in enumerate([(90, 120), (192, 192), (300, 111), (96, 96)]):
        img = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        p = str(tmp_path / ("r%d.png" % i))
        cv2.imwrite(p, img)
        raw.append(cv2.imread(p))
        paths.append(p)
    for rep in range(3):                      # both staging slots, and reuse
        eng.forward_raw(raw)
        y_raw = eng.read_output()
        u8 = eng.read_input_u8()
        dets_raw = eng.detect(0.5, 0.6)
    for i, img in enumerate(raw):
        assert np.array_equal(u8[i], opre.preprocess_u8(img, shape))
    x_host = np.concatenate([pbase.preprocess_image(p, shape)[0][None] for p in paths], 0)     # float64, reference protocol
    eng.forward(x_host)
    assert np.array_equal(eng.read_output(), y_raw)
    dets_host = eng.detect(0.5, 0.6)
    for a, b in zip(dets_raw, dets_host):
        assert np.array_equal(a, b)
    # smaller batch, other sizes
    eng.forward_raw(raw[:2])
    assert np.array_equal(eng.read_output(), y_raw[:2])
    with pytest.raises(Exception):
        eng.forward_raw([np.zeros((4, 4), np.uint8)])
    eng.close()
    # the reference's dsize quirk: a non-square network cannot be fed (its own placeholder rejects the resized array)
    net2, _, stream2 = helpers.build_v3((64, 96, 3), 80, seed=2)
    eng2 = engine.Engine(net2[0]._yb_state.plan(), (64, 96, 3), 80, engine.YB_DECODE_V3, max_batch=1)
    eng2.load_weights(stream2)
    with pytest.raises(Exception, match="non-square"):
        eng2.forward_raw(raw[:1])
    eng2.close()
