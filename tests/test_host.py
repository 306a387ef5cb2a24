"""CPU: host logic -- plan builder, weight files, config/launcher surface, and the C-ABI library
(loads, exports every declared symbol, refuses to compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from tensorflow_yolo_b200 import _lib, engine, plan as P, synth
from tensorflow_yolo_b200.net import base, layers, v2, v3, yolo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_v3_plan_counts_match_survey():
    net, topo, _ = helpers.build_v3((416, 416, 3), 80)
    specs = net[0]._yb_state.graph.specs
    assert len(net) == 109 and len(specs) == 109
    assert P.weight_count(specs) == 62001757
    assert P.conv_flops(specs) == 65864075264
    assert net[-1].out.get_shape().as_list() == [None, 10647, 85]
    assert [(l.h, l.w, l.b) for l in net[-1].yolos] == [(13, 13, 3), (26, 26, 3), (52, 52, 3)]
    assert net[-1].yolos[0].anchors[0] == (116 / 32.0, 90 / 32.0)
    # load-bearing indices: route sources "61 + 1" and "36 + 1" (net/v3.py:59,75)
    assert specs[87].src == [86, 62] and specs[99].src == [98, 37]
    assert specs[62].shape == (26, 26, 512) and specs[37].shape == (52, 52, 256)
    # product plan == oracle topology
    strip = lambda d: {k: v for k, v in d.items() if k not in ("shape", "anchors")}
    assert [strip(s.as_dict()) for s in specs[1:]] == [strip(t) for t in topo[1:]]
    from oracle import convstack
    geo = convstack.yolo_geometry(topo, (416, 416, 3))
    assert [(l.h, l.w, l.b, l.anchors) for l in net[-1].yolos] == geo


def test_v3_608_rows():
    net = v3.create_network(np.reshape(helpers.V3_ANCHORS, [-1, 2]), helpers.names(80), False, input_shape=(608, 608, 3))
    assert net[-1].out.get_shape().as_list() == [None, 22743, 85]


def test_v2_plan_counts_match_survey():
    for nc, anchors, count in ((80, helpers.V2_ANCHORS_COCO, 50983561), (20, helpers.V2_ANCHORS_VOC, 50676061)):
        net, topo, _ = helpers.build_v2((416, 416, 3), nc, anchors=anchors)
        specs = net[0]._yb_state.graph.specs
        assert len(net) == 32
        assert P.weight_count(specs) == count
        assert net[-1].out.get_shape().as_list() == [None, 13, 13, 5 * (5 + nc)]
        assert specs[26].src == [17] and specs[29].src == [28, 25]
        strip = lambda d: {k: v for k, v in d.items() if k not in ("shape", "anchors")}
        assert [strip(s.as_dict()) for s in specs[1:]] == [strip(t) for t in topo[1:]]
        # the engine plan carries one extra YOLO entry with the grid-unit anchors
        eng_plan = net[0]._yb_state.plan()
        assert len(eng_plan) == 33 and eng_plan[-1].kind == P.KIND_YOLO and len(eng_plan[-1].anchors) == 5


def test_variable_names_order():
    net, _, _ = helpers.build_v3((64, 64, 3), 3)
    assert net[1].variable_names == ["yolo/conv2d_bn_act_0/beta", "yolo/conv2d_bn_act_0/gamma",
                                     "yolo/conv2d_bn_act_0/moving_mean", "yolo/conv2d_bn_act_0/moving_variance",
                                     "yolo/conv2d_bn_act_0/kernel"]
    head = [l for l in net if isinstance(l, layers.conv2d_bn_act)][58]
    assert head.variable_names == ["yolo/conv2d_bn_act_58/bias", "yolo/conv2d_bn_act_58/kernel"]
    assert sum(isinstance(l, layers.conv2d_bn_act) for l in net) == 75


def test_builder_rejects_training_and_bad_heads():
    with pytest.raises(NotImplementedError):
        v3.create_network(np.reshape(helpers.V3_ANCHORS, [-1, 2]), helpers.names(2), True)
    layers.reset_default_graph()
    inp = layers.input_layer([None, 32, 32, 3])
    c = layers.conv2d_bn_act(inp.out, 30, 1, use_batch_normalization=False, activation_fn="linear")
    with pytest.raises(ValueError):
        layers.yolo_layer(c.out, [(1, 1)] * 3, 4, (32, 32, 3))      # 30 != 3*(5+4)


def test_weight_file_roundtrip_and_short_stream(tmp_path, capsys):
    net, _, stream = helpers.build_v3((64, 64, 3), 2, seed=5)
    path = str(tmp_path / "w.weights")
    synth.write_weights_v3(path, stream)
    ops = v3.load_weights(net, path)
    assert len(ops) == 1 and np.array_equal(ops[0].stream, stream)
    assert "Weights ready ({}/{} read)".format(stream.size, stream.size) in capsys.readouterr().out
    synth.write_weights_v3(path, stream[:-10])
    with pytest.raises(ValueError):
        v3.load_weights(net, path)
    with pytest.raises(FileNotFoundError):
        v3.load_weights(net, str(tmp_path / "missing.weights"))
    net2, _, stream2 = helpers.build_v2((64, 64, 3), 20)
    for major, minor in ((0, 1), (0, 2)):
        synth.write_weights_v2(path, stream2, major=major, minor=minor)
        assert os.path.getsize(path) == 16 + 4 * stream2.size
        assert np.array_equal(v2.load_weights(net2, path)[0].stream, stream2)


def test_launcher_config_surface(tmp_path):
    import launcher
    ini = tmp_path / "cfg" / "yolo.ini"
    ini.parent.mkdir()
    ini.write_text("[COMMON]\nversion = v3\ninput_h = 416\ninput_w = 416\ninput_c = 3\n"
                   "[TEST]\nimage_dir = ../img/\nout_dir = /abs/out\nbatch_size = 1\nthreshold = 0.5\n"
                   "iou_threshold = 0.6\nanchors = [10, 13, 16, 30]\nclass_names = [\"a\",\"b\"]\n"
                   "checkpoint_path = ../a\npretrained_weights_path = ../bin/yolov3.weights\ncpu_only = True\n")
    cfg = launcher.load_config(str(ini))
    assert cfg["TEST"]["image_dir"] == os.path.join(str(ini.parent), "../img/")
    assert cfg["TEST"]["out_dir"] == "/abs/out"
    assert cfg["TEST"]["anchors"] == [10, 13, 16, 30] and cfg["TEST"]["class_names"] == ["a", "b"]
    with pytest.raises(ValueError):
        launcher._main({"COMMON": {"version": "v9"}}, "test")
    with pytest.raises(ValueError):
        launcher._main({"COMMON": {"version": "v3"}, "TEST": {}}, "bogus")
    assert isinstance(yolo.YoloV3(), yolo.Yolo) and yolo.YoloV2.create_network is v2.create_full_network.__func__


def test_shipped_configs_parse():
    import launcher
    for name, version in (("yolo_3.ini", "v3"), ("yolo_2.ini", "v2")):
        cfg = launcher.load_config(os.path.join(ROOT, "config", name))
        p = {**cfg["TEST"], **cfg["COMMON"]}
        assert p["version"] == version
        for key in ("image_dir", "out_dir", "batch_size", "threshold", "iou_threshold", "anchors", "class_names",
                    "input_h", "input_w", "input_c", "checkpoint_path", "pretrained_weights_path", "cpu_only"):
            assert key in p, key


def test_bounding_box_record():
    b = base.BoundingBox(x=np.float32(.5), y=np.float32(.25), w=np.float64(.2), h=np.float64(.1), class_idx=3, prob=.9)
    assert b.get_top_left(100, 200) == ((.5 - .1) * 200, (.25 - .05) * 100)
    assert b.get_bottom_right() == (.5 + .2 / 2., .25 + .1 / 2.)
    assert base.non_maximum_suppression([], 0.6) == []


def test_cabi_exports_every_declared_symbol(lib_built):
    header = open(os.path.join(ROOT, "include", "yolo_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", header)))
    assert declared == _lib.EXPORTS, set(declared) ^ set(_lib.EXPORTS)
    handle = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().yb_abi_version() == 2
    assert ctypes.sizeof(_lib.yb_det) == 40 and np.dtype(_lib.DET_DTYPE).itemsize == 40
    assert ctypes.sizeof(P.yb_layer) == 4 * (7 + 4 + 1 + 32)


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device every compute entry point must fail loudly (never fall back to the CPU)."""
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    net, _, _ = helpers.build_v3((64, 64, 3), 2)
    st = net[0]._yb_state
    with pytest.raises(_lib.YoloB200Error) as ei:
        engine.Engine(st.plan(), (64, 64, 3), 2, engine.YB_DECODE_V3)
    assert ei.value.code == _lib.YB_ERR_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(_lib.YoloB200Error):
        engine.PostProcessor([(2, 2, [(1., 1.)])], 2, engine.YB_DECODE_V3)
    z = np.zeros(2, np.float32)
    with pytest.raises(_lib.YoloB200Error):
        engine.nms(z, z, z + 1, z + 1, z + .5, 0.6)
    assert len(engine.nms(z[:0], z[:0], z[:0], z[:0], z[:0], 0.6)) == 0     # empty list needs no device (base.py:196-197)
    # argument validation does not need a device either
    h = ctypes.c_void_p()
    assert _lib.lib().yb_engine_create(None, 0, 1, 1, 1, 1, 0, 0, 1, ctypes.byref(h)) == _lib.YB_ERR_INVALID


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensorflow_yolo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
    src = open(os.path.join(ROOT, "launcher.py")).read()
    assert "oracle" not in src


def test_generate_raw_batch_threaded_order_and_late_failure(tmp_path, monkeypatch, capsys, lib_built):
    """Files are decoded by a thread pool one batch ahead: same order, same batch boundaries as the reference's
    generate_test_batch (net/base.py:158-168), and an unreadable file ends the run only when its batch is reached."""
    import cv2
    from tensorflow_yolo_b200.net import base as pbase
    rs = np.random.RandomState(3)
    paths = []
    for i in range(7):
        p = str(tmp_path / ("im%02d.png" % i))
        cv2.imwrite(p, rs.randint(0, 256, size=(20 + i, 30, 3)).astype(np.uint8))
        paths.append(p)
    for threads in ("1", "4"):
        monkeypatch.setenv("YB_DECODE_THREADS", threads)
        got = list(pbase.generate_raw_batch(paths, 3))
        assert [len(b[0]) for b in got] == [3, 3, 1] and [p for _, ps in got for p in ps] == paths
        for imgs, ps in got:
            for im, p in zip(imgs, ps):
                assert np.array_equal(im, cv2.imread(p))
        bad = paths[:4] + [str(tmp_path / "missing.png")] + paths[4:]
        gen = pbase.generate_raw_batch(bad, 3)
        first = next(gen)                                   # the first batch is delivered although batch 2 will fail
        assert first[1] == paths[:3]
        with pytest.raises(TypeError):
            next(gen)
        assert "Failed to read" in capsys.readouterr().out
    assert list(pbase.generate_raw_batch([], 3)) == []
