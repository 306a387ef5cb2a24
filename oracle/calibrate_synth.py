"""Measures the pre-BN statistics the synthetic-weight generator needs -- TEST INFRASTRUCTURE.

Run in the build container:  python -m oracle.calibrate_synth
Writes tensorflow_yolo_b200/synth_calib.json: per network, the mean/variance of every BN conv's
pre-BN output and the input second moment of every head conv, measured with the fp32 oracle
(oracle.convstack) on seed-2 kernels and seed-1 images at 416x416.  The product-side generator
(tensorflow_yolo_b200/synth.py) only reads the table; it never imports the oracle.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import convstack  # noqa: E402
from tensorflow_yolo_b200 import plan as _plan, synth  # noqa: E402
from tensorflow_yolo_b200.net import v2 as pv2, v3 as pv3  # noqa: E402

V3_ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]
V2_ANCHORS = [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828]


def calibrate(topology, specs, num_classes, shape, seed=2, n_images=1):
    unit = synth.weight_stream(specs, seed=seed, num_classes=num_classes, calib={"bn": [(0.0, 1.0)] * 200, "head_m2": [1.0] * 8})
    params, _ = convstack.split_stream(topology, unit)
    x0 = torch.as_tensor(synth.images(n_images, shape[0], shape[1])).permute(0, 3, 1, 2).contiguous()
    outs, bn_tab, head_tab = [], [], []
    with torch.no_grad():
        for i, rec in enumerate(topology):
            kind = rec["kind"]
            src = [outs[s] for s in rec["src"]]
            if kind == "input":
                y = x0
            elif kind == "conv":
                p = params[i]
                k, s = rec["ksize"], rec["stride"]
                x = src[0]
                w = torch.as_tensor(p["kernel"])
                if rec["bn"]:
                    y = F.conv2d(convstack._pad(x, k), w, None, stride=s) if s > 1 else F.conv2d(x, w, None, padding=(k - 1) // 2)
                    m, v = float(y.mean()), float(y.var())
                    bn_tab.append([m, v])
                    mean = m + torch.as_tensor(p["moving_mean"]) * np.sqrt(v)     # unit table: mean = jitter
                    var = v * torch.as_tensor(p["moving_variance"])              # unit table: var = jitter
                    inv = torch.rsqrt(var + convstack.BN_EPS) * torch.as_tensor(p["gamma"])
                    y = y * inv.view(1, -1, 1, 1) + (torch.as_tensor(p["beta"]) - mean * inv).view(1, -1, 1, 1)
                    y = torch.maximum(y * convstack.LEAKY, y)
                else:
                    head_tab.append(float((x * x).mean()))
                    y = x[:, :1]       # heads are not consumed by any conv
            elif kind == "maxpool":
                y = F.max_pool2d(src[0], 2, 2)
            elif kind == "route":
                y = torch.cat(src, dim=1)
            elif kind == "reorg":
                st = rec["stride"]
                n, c, h, w_ = src[0].shape
                y = src[0].reshape(n, c, h // st, st, w_ // st, st).permute(0, 3, 5, 1, 2, 4).reshape(n, st * st * c, h // st, w_ // st)
            elif kind == "shortcut":
                y = src[0] + src[1]
            elif kind == "upsample":
                y = src[0].repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
            else:
                y = src[0]
            outs.append(y)
    return {"signature": synth._signature(specs), "seed": seed, "input": list(shape), "bn": bn_tab, "head_m2": head_tab}


def main():
    tables = []
    names80 = ["c"] * 80
    shape = (416, 416, 3)
    net = pv3.create_network(np.reshape(V3_ANCHORS, [-1, 2]), names80, False, input_shape=shape)
    specs = net[0]._yb_state.graph.specs
    tables.append(calibrate(convstack.topology_v3(80, np.reshape(V3_ANCHORS, [-1, 2]), shape), specs, 80, shape))
    net = pv2.create_full_network(np.reshape(V2_ANCHORS, [-1, 2]), names80, False, input_shape=shape)
    specs = net[0]._yb_state.graph.specs
    tables.append(calibrate(convstack.topology_v2(80, 5, shape), specs, 80, shape))
    out = os.path.join(ROOT, "tensorflow_yolo_b200", "synth_calib.json")
    with open(out, "w") as f:
        json.dump({"made_by": "oracle/calibrate_synth.py", "tables": tables}, f, indent=0)
    print("wrote", out, [len(t["bn"]) for t in tables])


if __name__ == "__main__":
    main()
