"""GPU measurement tool (not a test): per-layer-class roofline probes of the tcgen05 conv kernel.

  python tools/conv_probe.py [--batch 128] [--size 416] [--out gpurun_out/probe.json]

For every distinct conv shape of YOLOv3 it times the op in isolation (CUDA events, back-to-back launches) with the
kernel's ablation switches (yb_engine_set_option("ablate", mask)): full kernel, mainloop only (no epilogue work),
loads only, MMA only, epilogue only ...  The slowest single role is what bounds the layer.  Also times whole
forwards with programmatic dependent launch on/off and with the heuristic vs the autotuned configurations.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from tensorflow_yolo_b200 import engine as yb  # noqa: E402

ABLATIONS = [("full", 0), ("no_epi", 1), ("no_mma", 2), ("loads_only", 3), ("a_only", 11), ("b_only", 7),
             ("mma_epi", 12), ("mma_only", 13), ("epi_only", 14)]


def forward_ms(eng, x, steps=10):
    for _ in range(3):
        eng.forward(x)
    eng.sync()
    eng.mark(0)
    for _ in range(steps):
        eng.forward(x)
    eng.mark(1)
    eng.sync()
    return eng.elapsed_ms(0, 1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.json"))
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    B = args.batch
    net, state, stream, shape = bench.build_network(args.size)
    eng = yb.Engine(state.plan(), shape, bench.NUM_CLASSES, yb.YB_DECODE_V3, max_batch=B)
    eng.load_weights(stream)
    x = torch.rand((B,) + shape, device="cuda", dtype=torch.float32)
    specs = state.graph.specs
    result = {"batch": B, "size": args.size}

    result["forward_ms_heuristic_pdl"] = forward_ms(eng, x)
    eng.set_option("pdl", 0)
    result["forward_ms_heuristic_nopdl"] = forward_ms(eng, x)
    eng.set_option("pdl", 1)

    # ---- ablation table per conv class (heuristic configuration) ----
    n_ops = len(eng.profile(x))
    seen, table = set(), []
    for op_i in range(n_ops):
        info = eng.op_info(op_i)
        if info["path"] != 0:
            continue
        spec = specs[info["layer"]]
        cin = specs[spec.src[0]].shape[2]
        key = (tuple(spec.shape), spec.ksize, spec.stride, cin)
        if key in seen:
            continue
        seen.add(key)
        row = {"op": op_i, "layer": info["layer"], "out_hwc": list(spec.shape), "k": spec.ksize, "stride": spec.stride, "cin": cin,
               "gflop": info["flops_per_image"] * B / 1e9}
        for name, mask in ABLATIONS:
            eng.set_option("ablate", mask)
            row[name] = eng.time_op(op_i, B, args.reps)
        eng.set_option("ablate", 0)
        # in-kernel cycle counters: which role waits on which (full kernel, and the epilogue running alone)
        eng.set_option("cycles", 1)
        for tag, mask in (("roles", 0), ("roles_epi_only", 14)):
            eng.set_option("ablate", mask)
            eng.read_cycles(reset=True)
            eng.time_op(op_i, B, args.reps)
            cyc = eng.read_cycles(reset=True)
            pt, mt, et = max(cyc["prod_total"], 1), max(cyc["mma_total"], 1), max(cyc["epi_total"], 1)
            row[tag] = {"prod_wait_empty": cyc["prod_wait_empty"] / pt, "mma_wait_full": cyc["mma_wait_full"] / mt,
                        "mma_wait_tmem": cyc["mma_wait_tmem"] / mt, "epi_wait_acc": cyc["epi_wait_acc"] / et,
                        "epi_wait_res": cyc["epi_wait_res"] / et, "epi_wait_buf": cyc["epi_wait_buf"] / et,
                        "epi_tmem_ld": cyc["epi_tmem_ld"] / et, "epi_math": cyc["epi_math"] / et,
                        "epi_fence_store": cyc["epi_fence_store"] / et}
        eng.set_option("ablate", 0)
        eng.set_option("cycles", 0)
        row["cfg"] = eng.op_cfg(op_i)
        row["tflops"] = row["gflop"] / row["full"] if row["full"] > 0 else 0.0
        table.append(row)
        print("op {:3d} {:>16s} k{} s{} cin{:4d} cfg {} | ".format(op_i, str(tuple(spec.shape)), spec.ksize, spec.stride, cin, row["cfg"]) +
              " ".join("{}={:.3f}".format(n, row[n]) for n, _ in ABLATIONS) + " | {:.0f} TF".format(row["gflop"] / row["full"]), flush=True)
        print("        roles: " + " ".join("{}={:.0%}".format(k, v) for k, v in row["roles"].items()), flush=True)
        print("        epi_only roles: " + " ".join("{}={:.0%}".format(k, v) for k, v in row["roles_epi_only"].items() if k.startswith("epi")), flush=True)
    result["ablation"] = table

    # ---- autotune ----
    eng.forward(x)
    report = eng.autotune(B, reps=5)
    result["tune"] = report
    # interleaved A/B (power/thermal drift moves single measurements by a few percent): tuned vs heuristic, PDL on
    eng.forward(x)
    eng.sync()
    tuned_cfg = {o: eng.op_cfg(o) for o in range(n_ops) if eng.op_info(o)["path"] == 0}
    ab = {"tuned": [], "heuristic": []}
    for _ in range(4):
        for o, c in tuned_cfg.items():
            eng.set_conv_cfg(o, c["bn"], c["pair"], c["bstat"], c["tma_epi"], c["ksub"])
        ab["tuned"].append(forward_ms(eng, x))
        for o in tuned_cfg:
            eng.set_conv_cfg(o, 0)
        ab["heuristic"].append(forward_ms(eng, x))
    result["forward_ms_ab"] = ab
    for o, c in tuned_cfg.items():
        eng.set_conv_cfg(o, c["bn"], c["pair"], c["bstat"], c["tma_epi"], c["ksub"])
    result["forward_ms_tuned_pdl"] = forward_ms(eng, x)
    eng.set_option("pdl", 0)
    result["forward_ms_tuned_nopdl"] = forward_ms(eng, x)
    eng.set_option("pdl", 1)
    for o in report["ops"]:
        best = min(o["candidates"], key=lambda c: c["ms"])
        print("tune op {:3d} cin{:4d} cout{:4d} k{} s{} {}x{} default {:.4f} best {:.4f} {} chosen {}".format(
            o["op"], o["cin"], o["cout"], o["k"], o["stride"], o["ho"], o["wo"], o["default_ms"], best["ms"],
            {k: best[k] for k in ("bn", "pair", "bstat", "tma_epi", "stages", "ksub")}, o["chosen"]), flush=True)
    print({k: v for k, v in result.items() if k.startswith("forward_ms")}, flush=True)
    with open(args.out, "w") as f:
        json.dump(result, f, indent=0)


if __name__ == "__main__":
    main()
