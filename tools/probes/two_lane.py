"""GPU probe (not a test): does splitting the batch into L concurrent lanes (L engines of batch 128/L on their own
streams) fill the per-layer tails and launch gaps of the single-stream forward?

  python tools/probes/two_lane.py [--total 128] [--lanes 1,2,4] [--iters 30] [--out gpurun_out/two_lane.json]

Every lane is a complete engine (own arena, own stream); a step enqueues forward + decode + NMS on every lane.  Wall
clock around `iters` steps after warm-up, all lanes synchronised at both ends."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from tensorflow_yolo_b200 import engine as yb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=128)
    ap.add_argument("--lanes", default="1,2,4")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "two_lane.json"))
    args = ap.parse_args()
    import torch
    net, state, stream, shape = bench.build_network(416, "v3")
    rows = []
    for lanes in [int(v) for v in args.lanes.split(",")]:
        n = args.total // lanes
        engs, xs = [], []
        for i in range(lanes):
            e = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=n, device=0)
            e.load_weights(stream)
            e.set_option("graph", 0)
            engs.append(e)
            g = torch.Generator(device="cuda"); g.manual_seed(1 + i)
            xs.append(torch.rand((n,) + shape, device="cuda", dtype=torch.float32, generator=g))
        torch.cuda.synchronize()

        def step():
            for e, x in zip(engs, xs):
                e.forward(x)
            for e in engs:
                e.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
        for _ in range(5):
            step()
        for e in engs:
            e.sync()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            step()
        for e in engs:
            e.sync()
        dt = time.perf_counter() - t0
        row = {"lanes": lanes, "batch_per_lane": n, "ms_per_step": 1e3 * dt / args.iters, "images_per_s": lanes * n * args.iters / dt}
        rows.append(row)
        print(json.dumps(row), flush=True)
        for e in engs:
            e.close()
    with open(args.out, "w") as f:
        json.dump({"what": "YOLOv3-416 forward+decode+NMS, {} images per step split over concurrent lanes (one engine and stream "
                           "per lane), wall clock".format(args.total), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
