"""Times selected conv ops in isolation (yb_engine_time_op: back-to-back launches, CUDA events) -- A/B helper.
  python tools/probes/time_layers.py [op indices ...]      (default: the 208x208 / 104x104 layers and one of each big class)
"""
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tensorflow_yolo_b200 import engine as yb  # noqa: E402
import torch  # noqa: E402

ops = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5, 6, 10, 11, 28, 29, 50, 51]
net, state, stream, shape = bench.build_network(416, "v3")
B = 128
eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
eng.load_weights(stream)
x = torch.rand((B,) + shape, device="cuda")
eng.forward(x); eng.sync()
res = {}
for rep in range(3):
    for op in ops:
        ms = eng.time_op(op, B, reps=20)
        res[op] = min(res.get(op, 1e9), ms)
print(" ".join("op%d=%.4f" % (op, res[op]) for op in ops), " sum=%.4f" % sum(res.values()))
for _ in range(3):
    eng.forward(x)
eng.sync(); eng.mark(0)
for _ in range(20):
    eng.forward(x)
eng.mark(1); eng.sync()
print("forward ms %.4f" % (eng.elapsed_ms(0, 1) / 20))
