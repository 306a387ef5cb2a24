"""GPU: the reference's entry point -- launcher config -> YoloV3().test(params) -- on synthetic images and a
synthetic darknet .weights file; results compared with the oracle run on the same files."""
import os

import numpy as np
import pytest

import helpers
from oracle import convstack, postprocess
from tensorflow_yolo_b200 import synth

pytestmark = pytest.mark.gpu


def _write_case(tmp_path, version, shape, nc, anchors, stream):
    import cv2
    img_dir = tmp_path / "img"
    img_dir.mkdir()
    rs = np.random.RandomState(7)
    for i in range(3):
        cv2.imwrite(str(img_dir / ("im%d.png" % i)), rs.randint(0, 256, size=(90 + 10 * i, 120, 3)).astype(np.uint8))
    wpath = tmp_path / "bin" / "net.weights"
    wpath.parent.mkdir()
    (synth.write_weights_v3 if version == "v3" else synth.write_weights_v2)(str(wpath), stream)
    ini = tmp_path / "config" / "net.ini"
    ini.parent.mkdir()
    ini.write_text(
        "[COMMON]\nversion = {}\ninput_h = {}\ninput_w = {}\ninput_c = 3\n[TEST]\nimage_dir = ../img/\nout_dir = ../out/\n"
        "batch_size = 2\nthreshold = 0.5\niou_threshold = 0.6\nanchors = {}\nclass_names = {}\n"
        "checkpoint_path = ../nope\npretrained_weights_path = ../bin/net.weights\ncpu_only = True\n".format(
            version, shape[0], shape[1], list(anchors), helpers.names(nc)).replace("'", '"'))
    return str(ini)


@pytest.mark.parametrize("version", ["v3", "v2"])
def test_launcher_test_mode(tmp_path, version, capsys):
    import launcher
    from tensorflow_yolo_b200.net import base as pbase
    shape = (128, 128, 3)
    if version == "v3":
        nc, anchors = 80, helpers.V3_ANCHORS
        net, topo, stream = helpers.build_v3(shape, nc, seed=2, obj_bias=-1.0)
    else:
        nc, anchors = 20, helpers.V2_ANCHORS_VOC
        net, topo, stream = helpers.build_v2(shape, nc, seed=3, obj_bias=1.0)
    ini = _write_case(tmp_path, version, shape, nc, anchors, stream)
    results = launcher._main(launcher.load_config(ini), "test")
    out = capsys.readouterr().out
    assert "Pre-trained weights loaded." in out and out.strip().endswith("Done")
    assert len(results) == 3
    for path, boxes in results.items():
        assert os.path.exists(os.path.join(str(tmp_path), "out", os.path.splitext(os.path.basename(path))[0] + "_out.png"))
        image, _ = pbase.preprocess_image(path, shape)
        ref_out = convstack.forward(topo, stream, image[None].astype(np.float32))
        if version == "v3":
            ref = postprocess.find_bounding_boxes_v3(ref_out, convstack.yolo_geometry(topo, shape), 0.5, 0.6)[0]
        else:
            ref = postprocess.find_bounding_boxes_v2(ref_out, np.reshape(anchors, [-1, 2]), nc, 0.5, 0.6)[0]
        # bf16 conv stack vs fp32: detections agree up to boxes whose score sits near the threshold
        got_rows = set()
        assert abs(len(boxes) - len(ref["row"])) <= max(3, len(ref["row"]) // 5)
        if len(ref["row"]) and len(boxes):
            ref_xy = np.stack([ref["x"], ref["y"]], 1)
            hits = 0
            for b in boxes:
                d = np.abs(ref_xy - np.asarray([b.x, b.y])).max(1)
                hits += bool(d.min() < 0.02)
            assert hits >= 0.7 * len(boxes)


def test_checkpoint_is_tried_first_and_gives_the_same_detections(tmp_path, capsys):
    """net/yolo.py:71-78: a readable checkpoint_path wins over pretrained_weights_path; the detections are those of
    the .weights run bit for bit (both sources reach the engine as the same float stream)."""
    import launcher
    from tensorflow_yolo_b200 import checkpoint as ckpt
    shape = (128, 128, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2, obj_bias=-1.0)
    ini = _write_case(tmp_path, "v3", shape, 80, helpers.V3_ANCHORS, stream)
    from_weights = launcher._main(launcher.load_config(ini), "test")
    assert "Pre-trained weights loaded." in capsys.readouterr().out
    ckpt.checkpoint_from_stream(net, stream, str(tmp_path / "ckpt" / "yolo.ckpt"), extra={"global_step": np.asarray(3, np.int64)})
    os.remove(str(tmp_path / "bin" / "net.weights"))            # the .weights file must not be needed any more
    text = open(ini).read().replace("checkpoint_path = ../nope", "checkpoint_path = ../ckpt/yolo.ckpt")
    open(ini, "w").write(text)
    from_ckpt = launcher._main(launcher.load_config(ini), "test")
    out = capsys.readouterr().out
    assert "restored." in out and "Pre-trained weights loaded." not in out
    assert sorted(from_ckpt) == sorted(from_weights)
    for path in from_weights:
        a, b = from_weights[path], from_ckpt[path]
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert (x.x, x.y, x.w, x.h, x.prob, x.class_idx) == (y.x, y.y, y.w, y.h, y.prob, y.class_idx)
