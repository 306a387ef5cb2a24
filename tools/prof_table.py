"""Prints a per-layer-class table from bench.py --dump-profile JSON files (optionally side by side)."""
import json
import sys
from collections import defaultdict


def load(p):
    d = json.load(open(p))
    g = defaultdict(lambda: [0, 0.0, 0.0, None])
    for o in d["ops"]:
        key = (o["kind"], tuple(o["out_hwc"]), o["ksize"], o["stride"], o["cin"])
        g[key][0] += 1; g[key][1] += o["ms"]; g[key][2] += o["tflops"] * o["ms"]; g[key][3] = (o["bn"], o["bk"], o["stages"])
    return d, g


def main():
    files = sys.argv[1:]
    loaded = [load(f) for f in files]
    base = loaded[0][1]
    print("step ms:", [round(d["step_ms"], 3) for d, _ in loaded], " sum of ops:", [round(sum(v[1] for v in g.values()), 3) for _, g in loaded])
    B = loaded[0][0]["batch"]
    for k, (n, ms, tf, cfg) in sorted(base.items(), key=lambda kv: -kv[1][1]):
        h, w, c = k[1]
        cin, ks, st = k[4], k[2], k[3]
        hin, win = h * st, w * st
        bytes_min = B * n * (hin * win * cin * 2 + h * w * c * 2) if k[0] == "conv" else 0
        line = "%-8s %-16s k%d s%d cin%-4d n=%2d hbm_floor %.3f |" % (k[0], str(k[1]), ks, st, cin, n, bytes_min / 6.54e12 * 1e3)
        for d, g in loaded:
            n2, ms2, tf2, cfg2 = g.get(k, (0, 0.0, 0.0, None))
            line += " %-13s %6.3f ms %5.0f TF |" % (str(cfg2), ms2, tf2 / ms2 if ms2 else 0)
        print(line)


if __name__ == "__main__":
    main()
