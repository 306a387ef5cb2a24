"""GPU measurement tool (not a test): every conv class of YOLOv3-416 at batch 128 through cuDNN (torch.nn.functional.conv2d,
bf16, channels_last, cudnn.benchmark on) next to this repo's kernels (yb_engine_time_op after autotuning).

  python tools/cudnn_yardstick.py [--batch 128] [--out gpurun_out/cudnn_yardstick.json]

cuDNN is a yardstick only: it is LIBRARY code and never on the product path.  Its numbers are the bare conv
(no BN / leaky / residual epilogue, bf16 output); ours include the fused epilogue (and, for the two fused pairs of
conv_fused.cuh, both layers -- compared with the sum of the two cuDNN convs)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from tensorflow_yolo_b200 import engine as yb, plan as yplan  # noqa: E402


def time_cudnn(torch, n, cin, cout, k, stride, h, w, reps):
    import torch.nn.functional as F
    x = torch.randn((n, cin, h, w), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = torch.randn((cout, cin, k, k), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    pad = (k - 1) // 2
    for _ in range(3):
        F.conv2d(x, wt, None, stride, pad)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        F.conv2d(x, wt, None, stride, pad)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cudnn_yardstick.json"))
    args = ap.parse_args()
    import torch
    torch.backends.cudnn.benchmark = True
    B = args.batch
    net, state, stream, shape = bench.build_network(416, "v3")
    specs = state.graph.specs
    eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
    eng.load_weights(stream)
    eng.autotune(B, reps=3)
    # launched ops in plan order: collect per conv op (layer, shape) and time it; the producer of a fused pair has 0 FLOPs
    n_ops = len([s for s in specs if s.kind == yplan.KIND_CONV])
    classes = {}
    op_i = 0
    while True:
        try:
            info = eng.op_info(op_i)
        except Exception:
            break
        layer = info["layer"]
        spec = specs[layer]
        if spec.kind == yplan.KIND_CONV:
            cin = specs[spec.src[0]].shape[2]
            h, w, cout = spec.shape
            key = (cin, cout, spec.ksize, spec.stride, h * spec.stride, w * spec.stride)
            c = classes.setdefault(key, {"count": 0, "ours_ms": 0.0, "paths": set(), "ops": []})
            c["count"] += 1
            c["paths"].add(info["path"])
            c["ops"].append(op_i)
        op_i += 1
    rows = []
    for key, c in classes.items():
        cin, cout, k, stride, hin, win = key
        ours = eng.time_op(c["ops"][0], B, reps=args.reps)          # one representative op per class
        cud = time_cudnn(torch, B, max(cin, 1), cout, k, stride, hin, win, args.reps) if cin >= 3 else None
        flops = 2.0 * B * (hin // stride) * (win // stride) * cout * k * k * cin
        rows.append({"cin": cin, "cout": cout, "k": k, "stride": stride, "in_hw": [hin, win], "launches_per_step": c["count"],
                     "paths": sorted(c["paths"]), "ours_ms": ours, "cudnn_ms": cud, "gflop": flops / 1e9,
                     "flop_share": None})
    total = sum(r["gflop"] * r["launches_per_step"] for r in rows)
    for r in rows:
        r["flop_share"] = r["gflop"] * r["launches_per_step"] / total
    # fused pairs: our time sits on the consumer (path 3), the producer reports ~0; compare with the sum of both cuDNN convs
    out = {"batch": B, "note": "ours includes BN/leaky/residual epilogues; path 3 = fused pair (time of both layers on the consumer row, "
                               "~0 on the producer row); cuDNN = bare bf16 NHWC conv, cudnn.benchmark", "rows": rows}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("%-34s %4s %9s %9s %7s %6s" % ("conv class (cin->cout k s @in)", "n", "ours ms", "cuDNN ms", "ratio", "share"))
    for r in sorted(rows, key=lambda r: -r["flop_share"]):
        ratio = (r["cudnn_ms"] / r["ours_ms"]) if r["cudnn_ms"] and r["ours_ms"] > 1e-4 else float("nan")
        print("%4d->%-4d k%d s%d @%dx%-12d %4d %9.4f %9s %7.2f %5.1f%%" % (
            r["cin"], r["cout"], r["k"], r["stride"], r["in_hw"][0], r["in_hw"][1], r["launches_per_step"], r["ours_ms"],
            "%.4f" % r["cudnn_ms"] if r["cudnn_ms"] else "-", ratio, 100 * r["flop_share"]))


if __name__ == "__main__":
    main()
