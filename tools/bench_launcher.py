"""GPU measurement tool (not a test): the reference's entry point end to end on a directory of JPEG files --
YoloV3().test(params) as launcher.py calls it (net/yolo.py:41-96): imread, preprocessing, network, decode, NMS, drawing,
saving -- with the file decoding single-threaded as in the reference and with the thread pool.

  python tools/bench_launcher.py [--images 256] [--batch 32] [--out gpurun_out/launcher.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from tensorflow_yolo_b200 import synth  # noqa: E402
from tensorflow_yolo_b200.net import yolo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "launcher.json"))
    args = ap.parse_args()
    import cv2
    tmp = tempfile.mkdtemp(prefix="yb_launcher_")
    img_dir, out_dir = os.path.join(tmp, "img"), os.path.join(tmp, "out")
    os.makedirs(img_dir)
    rs = np.random.RandomState(0)
    yy, xx = np.mgrid[0:480, 0:640]
    for i in range(args.images):                      # smooth synthetic photographs-like content: realistic JPEG sizes
        f = rs.uniform(0.005, 0.05, 6)
        im = np.stack([127 + 120 * np.sin(f[2 * c] * xx + f[2 * c + 1] * yy + i) for c in range(3)], -1)
        cv2.imwrite(os.path.join(img_dir, "im%04d.jpg" % i), (im + rs.normal(0, 6, im.shape)).clip(0, 255).astype(np.uint8))
    net, state, stream, shape = bench.build_network(416, "v3")
    wpath = os.path.join(tmp, "net.weights")
    synth.write_weights_v3(wpath, stream)
    params = {"image_dir": img_dir, "out_dir": out_dir, "batch_size": str(args.batch), "threshold": "0.5", "iou_threshold": "0.6",
              "anchors": bench.V3_ANCHORS, "class_names": ["c%d" % i for i in range(80)], "input_h": "416", "input_w": "416",
              "input_c": "3", "checkpoint_path": os.path.join(tmp, "nope"), "pretrained_weights_path": wpath, "cpu_only": "False"}
    out = {"workload": "YoloV3().test(params): %d JPEG files 640x480, batch_size %d, draw + save included" % (args.images, args.batch), "rows": []}
    devnull = open(os.devnull, "w")
    for decode_thr, draw_thr in ((1, 1), (1, 4), (8, 8), (16, 16)):
        os.environ["YB_DECODE_THREADS"] = str(decode_thr)
        os.environ["YB_DRAW_THREADS"] = str(draw_thr)
        best = None
        for rep in range(2):                          # the first pass also builds the engine and tunes it
            stdout, sys.stdout = sys.stdout, devnull
            t0 = time.perf_counter()
            try:
                res = yolo.YoloV3().test(params)
            finally:
                sys.stdout = stdout
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        row = {"decode_threads": decode_thr, "draw_threads": draw_thr, "seconds": best, "images_per_s": args.images / best,
               "images": len(res)}
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
