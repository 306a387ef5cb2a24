"""Small self-checking NMS target: a few calls that exercise the pipelined variant (several blocks, irregular boxes,
per-class mode) and the generic variant (float32 boxes), each compared with the oracle.  Written as a
compute-sanitizer target (memcheck / racecheck); the sanitizer is closed on the B200 pool of this round, so it was only
run plain -- the synchronisation argument is written out in DESIGN.md section 4.2 instead.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorflow_yolo_b200 import engine  # noqa: E402
from oracle import postprocess  # noqa: E402
import test_gpu_post as T  # noqa: E402

for k, seed, irr in ((300, 1, True), (1500, 2, False)):
    c = T._clustered_boxes(np.random.RandomState(seed), k, irr)
    for thr in (0.3, 0.6):
        got = engine.nms(c["x"], c["y"], c["w"], c["h"], c["prob"], thr)
        assert np.array_equal(got, postprocess.nms(c, thr))
    cls = np.random.RandomState(seed).randint(0, 3, k).astype(np.int32)
    engine.nms(c["x"], c["y"], c["w"], c["h"], c["prob"], 0.45, class_idx=cls, nms_mode=engine.YB_NMS_PER_CLASS)
    got32 = engine.nms(c["x"], c["y"], c["w"].astype(np.float32), c["h"].astype(np.float32), c["prob"], 0.6)
    c32 = dict(c, w=c["w"].astype(np.float32), h=c["h"].astype(np.float32))
    assert np.array_equal(got32, postprocess.nms(c32, 0.6))
print("nms sanitize target ok")
