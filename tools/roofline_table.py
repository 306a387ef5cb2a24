"""Per-layer-class roofline table from a bench.py --dump-profile JSON (markdown on stdout).

  python tools/roofline_table.py profiles/r01_per_op_events_v3_416_b128_pdl.json > profiles/r01_per_layer_roofline.md

For every distinct conv shape: launches per step, measured time (CUDA events between launches, summed over the class),
the tensor-core floor (FLOPs / measured burst bf16 peak), the HBM floor (input + output activation bytes + weights,
each moved once, / measured HBM peak) and measured / max(floors).  The event pass serialises the launches (no PDL
overlap), so the class times add up to more than the step time reported by bench.py.
"""
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    d = json.load(open(sys.argv[1]))
    peaks = {"bf16_tflops": 1652.4, "hbm_gbs": 6539.9}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks.update({k: v for k, v in json.load(open(pk)).items() if k in peaks})
    B = d["batch"]
    g = defaultdict(lambda: {"n": 0, "ms": 0.0, "flop": 0.0, "bytes": 0.0, "cfg": None})
    ops = [o for o in d["ops"] if o["kind"] == "conv"]
    for n_op, o in enumerate(ops):
        h, w, c = o["out_hwc"]
        k, s, cin = o["ksize"], o["stride"], o["cin"]
        if o["path"] != 3 and o["ms"] < 2e-2 and n_op + 1 < len(ops) and ops[n_op + 1]["path"] == 3:
            continue                                     # the producer of a fused pair: accounted for on the consumer's row
        key = (h, w, c, k, s, cin, o["path"])
        e = g[key]
        e["n"] += 1
        e["ms"] += o["ms"]
        e["flop"] += 2.0 * B * h * w * c * k * k * cin
        out_b = 4 if (c % 85 == 0 and k == 1) else 2          # the heads are written in fp32
        e["bytes"] += B * (h * s * w * s * cin * (2 if cin > 3 else 4) + h * w * c * out_b) + c * k * k * cin * 2
        if o["path"] == 3:
            # fused pair (conv_fused.cuh): + the producer's FLOPs; bytes = the producer's input + this output (+ nothing for the
            # residual: it is the producer's input, read once), the intermediate never reaches HBM
            p = ops[n_op - 1]
            ph, pw, pc = p["out_hwc"]
            e["flop"] += 2.0 * B * ph * pw * pc * p["ksize"] * p["ksize"] * p["cin"]
            e["bytes"] += B * (ph * pw * p["cin"] * (2 if p["cin"] > 3 else 4)) - B * (h * s * w * s * cin * 2)
            e["cfg"] = "fused pair: {}x{} {}->{} ({}) + this (tcgen05 BN64 BK32)".format(
                p["ksize"], p["ksize"], p["cin"], pc, "mma.sync producer" if p["cin"] == 3 else "tcgen05 producer, the TMA patch as operand")
        else:
            e["cfg"] = "BN{} BK{}{}".format(o["bn"], o["bk"], " pair" if o.get("pair") else "") if o["path"] == 0 else "mma.sync"
    tot_ms = sum(e["ms"] for e in g.values())
    print("# Per-layer-class roofline, YOLOv3-416 batch {} on one B200".format(B))
    print()
    print("Source: `{}` (step {:.3f} ms without events, {:.3f} ms with an event between launches). Peaks: {:.1f} TFLOP/s bf16 "
          "(measured burst), {:.1f} GB/s HBM (measured).".format(os.path.basename(sys.argv[1]), d["step_ms"], d.get("step_ms_with_events", 0.0),
                                                             peaks["bf16_tflops"], peaks["hbm_gbs"]))
    print()
    print("| conv (out HxWxC, k, stride, Cin) | launches | kernel | measured ms | TFLOP/s | tensor floor ms | HBM floor ms | measured / roofline |")
    print("|---|---|---|---|---|---|---|---|")
    tot_floor = 0.0
    for key, e in sorted(g.items(), key=lambda kv: -kv[1]["ms"]):
        h, w, c, k, s, cin, _path = key
        t_tc = e["flop"] / (peaks["bf16_tflops"] * 1e12) * 1e3
        t_hbm = e["bytes"] / (peaks["hbm_gbs"] * 1e9) * 1e3
        floor = max(t_tc, t_hbm)
        tot_floor += floor
        print("| {}x{}x{} k{} s{} cin{} | {} | {} | {:.3f} | {:.0f} | {:.3f} | {:.3f} | {:.2f} |".format(
            h, w, c, k, s, cin, e["n"], e["cfg"], e["ms"], e["flop"] / (e["ms"] * 1e-3) / 1e12, t_tc, t_hbm, e["ms"] / floor))
    print("| **all convs** | {} | | **{:.3f}** | {:.0f} | | | **{:.2f}** (sum of per-class rooflines {:.3f} ms) |".format(
        sum(e["n"] for e in g.values()), tot_ms, sum(e["flop"] for e in g.values()) / (tot_ms * 1e-3) / 1e12, tot_ms / tot_floor, tot_floor))


if __name__ == "__main__":
    main()
