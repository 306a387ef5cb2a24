"""ANCHOR mode (net/v2.py:298-323, net/base.py:49-52,69-97) and the structured
(boxes, scores, classes) export of the TEST results.  Host-side logic only (no GPU)."""
import os

import numpy as np
import pytest

from oracle import refimport
from tensorflow_yolo_b200.net import base, v2, yolo

VOC = """<annotation><filename>{name}.jpg</filename><size><width>{w}</width><height>{h}</height><depth>3</depth></size>
{objects}</annotation>"""
OBJ = "<object><name>{n}</name><bndbox><xmin>{x1}</xmin><ymin>{y1}</ymin><xmax>{x2}</xmax><ymax>{y2}</ymax></bndbox></object>"


def _write_voc(d, rs, n_files=12):
    truth = []
    for i in range(n_files):
        w, h = int(rs.randint(200, 640)), int(rs.randint(200, 480))
        objs = []
        for j in range(int(rs.randint(1, 4))):
            big = (i + j) % 2
            bw, bh = (0.6, 0.7) if big else (0.1, 0.15)
            x1, y1 = int(rs.randint(0, w * (1 - bw) - 1)), int(rs.randint(0, h * (1 - bh) - 1))
            x2, y2 = x1 + int(bw * w), y1 + int(bh * h)
            name = "tower" if big else "car"
            objs.append(OBJ.format(n=name, x1=x1, y1=y1, x2=x2, y2=y2))
            truth.append(((x2 - x1) / w, (y2 - y1) / h, name))
        (d / ("a%02d.xml" % i)).write_text(VOC.format(name="a%02d" % i, w=w, h=h, objects="".join(objs)))
    (d / "notes.txt").write_text("ignored")
    return truth


def test_parse_annotations_and_anchor_mode(tmp_path, capsys):
    rs = np.random.RandomState(0)
    truth = _write_voc(tmp_path, rs)
    ann = base.parse_annotations(str(tmp_path), "/images", normalize=True)
    assert len(ann) == 12 and all(p.startswith("/images/a") and p.endswith(".jpg") for p, _ in ann)
    sizes = sorted((round(o[2] - o[0], 9), round(o[3] - o[1], 9), o[4]) for _, objs in ann for o in objs)
    assert sizes == sorted((round(w, 9), round(h, 9), n) for w, h, n in truth)
    raw = base.parse_annotations(str(tmp_path), "/images")
    assert all(isinstance(o[0], int) for _, objs in raw for o in objs)          # pixel coordinates stay ints
    params = {"num_anchors": "2", "image_dir": "/images", "annotation_dir": str(tmp_path), "tolerate": "0.0001", "stride": "32",
              "input_w": "416", "input_h": "416"}
    np.random.seed(0)
    anchors, names = yolo.YoloV2().generate_anchors(params)
    assert "12 annotations found." in capsys.readouterr().out
    assert names == {"tower", "car"} and anchors.shape == (4,)
    got = sorted(np.reshape(anchors, [-1, 2]).tolist())
    # two well separated clusters: centres are the cluster means, in grid units of a 416 input at stride 32
    small = np.mean([(w, h) for w, h, n in truth if n == "car"], 0) * 13
    big = np.mean([(w, h) for w, h, n in truth if n == "tower"], 0) * 13
    assert np.allclose(got, sorted([small.tolist(), big.tolist()]), rtol=1e-6)
    with pytest.raises(NotImplementedError):
        yolo.YoloV3().generate_anchors(params)                                  # the reference binds it for v2 only


def test_anchor_mode_matches_reference_on_its_own_dataset(capsys):
    """The reference's generate_anchors run live on resource/eiffel/train (its shipped ANCHOR config: 1 anchor)."""
    if not refimport.available():
        pytest.skip("/root/reference is not present")
    import launcher
    ref = refimport.load()
    root = refimport.REFERENCE_ROOT
    params = {"num_anchors": "1", "image_dir": os.path.join(root, "resource/eiffel/train/"),
              "annotation_dir": os.path.join(root, "resource/eiffel/train/"), "tolerate": "0.005", "stride": "32",
              "input_w": "416", "input_h": "416"}
    try:
        want_a, want_n = ref.v2.generate_anchors.__func__(params)
    except Exception as ex:                                                      # tqdm / sklearn missing for the reference import
        pytest.skip("reference ANCHOR mode cannot run here: {}".format(ex))
    got_a, got_n = v2.generate_anchors.__func__(params)
    assert got_n == want_n and np.allclose(got_a, want_a, rtol=1e-9)             # one cluster: the mean, deterministic
    cfg = {"COMMON": {"version": "v2", "input_w": "416", "input_h": "416"}, "ANCHOR": params}
    out_a, _ = launcher._main(cfg, "anchor")
    assert np.allclose(out_a, want_a) and "Anchors:" in capsys.readouterr().out


def test_boxes_to_arrays():
    boxes = [base.BoundingBox(x=np.float32(.5), y=np.float32(.25), w=np.float64(.2), h=np.float64(.1), class_idx=np.int64(3), prob=np.float32(.9)),
             base.BoundingBox(x=np.float32(.1), y=np.float32(.2), w=np.float64(.3), h=np.float64(.4), class_idx=np.int64(0), prob=np.float32(.6))]
    b, s, c = base.boxes_to_arrays(boxes)
    assert b.dtype == np.float64 and s.dtype == np.float32 and c.dtype == np.int64
    assert b.shape == (2, 4) and np.allclose(b[0], [.5, .25, .2, .1]) and s.tolist() == [np.float32(.9), np.float32(.6)] and c.tolist() == [3, 0]
    corners = base.boxes_to_corners(boxes, h=100, w=200)
    assert np.allclose(corners[0], [(.5 - .1) * 200, (.25 - .05) * 100, (.5 + .1) * 200, (.25 + .05) * 100])
    e = base.boxes_to_arrays([])
    assert e[0].shape == (0, 4) and e[1].shape == (0,) and e[2].shape == (0,)
    assert base.boxes_to_corners([]).shape == (0, 4)
