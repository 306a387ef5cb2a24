"""GPU measurement tool (not a test): the preprocessing row of SURVEY 8(f) -- cv2.resize (INTER_LINEAR) + BGR->RGB + /255 of
decoded images (net/base.py:115-155) on the device (yb_engine_forward_raw) against the reference's host code.

  python tools/bench_pre.py [--out gpurun_out/pre.json] [--batch 128] [--steps 10]

Decoded BGR images (uint8, 640x480 and 1280x720) sit in host memory, as cv2.imread leaves them.  Timed, host wall clock
around `steps` synchronised steps of forward + decode + NMS for YOLOv3-416, batch 128:
  raw      eng.forward_raw(images): upload of the decoded images, resize + BGR->RGB on the device, conv stack
  u8       eng.forward(preprocessed uint8 batch in host memory): the bench.py e2e feed (preprocessing not included)
  cpu      the reference's preprocess_image arithmetic with cv2 on one host core (images/s), for scale
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from tensorflow_yolo_b200 import engine as yb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "pre.json"))
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import cv2
    B = args.batch
    net, state, stream, shape = bench.build_network(416, "v3")
    eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
    eng.load_weights(stream)
    eng.autotune(B, reps=3)
    out = {"workload": "YOLOv3-416 batch %d, forward + decode + NMS, decoded BGR images in host memory" % B, "rows": []}
    rs = np.random.RandomState(0)
    for (h, w) in ((480, 640), (720, 1280)):
        imgs = [rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for _ in range(B)]
        pre = yb.resize_bgr2rgb(imgs, 416, 416)                     # the same preprocessing, once, for the u8 feed
        t0 = time.perf_counter()
        n_cpu = 16
        for im in imgs[:n_cpu]:
            x = cv2.resize(im, dsize=(416, 416))[:, :, ::-1] / 255.   # net/base.py:121-122,153
        cpu_ips = n_cpu / (time.perf_counter() - t0)
        assert x.shape == (416, 416, 3)
        row = {"source_hw": [h, w], "source_bytes_per_batch": B * h * w * 3, "cpu_preprocess_images_per_s_1core": cpu_ips}
        for name, fn in (("raw", lambda: eng.forward_raw(imgs)), ("u8", lambda: eng.forward(pre))):
            for _ in range(3):
                fn(); eng.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
            eng.sync()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn(); eng.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
            eng.sync()
            dt = (time.perf_counter() - t0) / args.steps
            row[name] = {"ms_per_step": dt * 1e3, "images_per_s": B / dt}
        row["preprocess_cost_ms_per_step"] = row["raw"]["ms_per_step"] - row["u8"]["ms_per_step"]
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
