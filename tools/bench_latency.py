"""GPU measurement tool (not a test): small-batch latency of the hot path with and without the CUDA graph of the forward.

  python tools/bench_latency.py [--out gpurun_out/latency.json] [--iters 200]

YOLOv3-416 COCO-80, random-init weights, images resident in HBM.  For each batch size: milliseconds per step (forward
+ decode + NMS) on the device (CUDA events on the engine's stream around `iters` back-to-back steps) and through the
host (wall clock around the same loop + sync), eager launches ("graph": 0) against graph replay ("graph": 1).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from tensorflow_yolo_b200 import engine as yb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "latency.json"))
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batches", default="1,2,4,8,16,32")
    ap.add_argument("--split-k", type=int, default=-1, help="latency mode (option split_k): -1 = measure both, 0 / 1 = only that setting")
    args = ap.parse_args()
    import torch
    net, state, stream, shape = bench.build_network(416, "v3")
    out = {"workload": "YOLOv3-416 COCO-80, forward (75 launches) + decode + NMS, inputs resident in HBM", "iters": args.iters, "rows": []}
    for n in [int(b) for b in args.batches.split(",")]:
        eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=n, device=0)
        eng.load_weights(stream)
        eng.autotune(n, reps=5)
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        x = torch.rand((n,) + shape, device="cuda", dtype=torch.float32, generator=g)
        row = {"batch": n}
        variants = [(0, "eager", 0), (1, "graph", 0), (1, "graph_split_k", 1)]
        if args.split_k >= 0:
            variants = [(0, "eager", args.split_k), (1, "graph", args.split_k)]
        for mode, name, sk in variants:
            eng.set_option("split_k", sk)
            eng.set_option("graph", mode)
            for _ in range(5):
                eng.forward(x); eng.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
            eng.sync()
            r0 = eng.graph_replays()
            t0 = time.perf_counter()
            eng.mark(0)
            for _ in range(args.iters):
                eng.forward(x); eng.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
            eng.mark(1)
            eng.sync()
            wall = (time.perf_counter() - t0) * 1e3 / args.iters
            dev = eng.elapsed_ms(0, 1) / args.iters
            row[name] = {"device_ms_per_step": dev, "host_ms_per_step": wall, "images_per_s": n / (wall * 1e-3),
                         "graph_replays": eng.graph_replays() - r0}
        row["speedup_host"] = row["eager"]["host_ms_per_step"] / row["graph"]["host_ms_per_step"]
        if "graph_split_k" in row:
            row["split_k_factors"] = sorted(set(eng.op_cfg(i)["splitk"] for i in range(75)))
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
        eng.close()
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
