"""Sweeps launch configurations of selected conv ops in isolation (yb_engine_set_conv_cfg + yb_engine_time_op).
  python tools/probes/cfg_sweep.py [op indices ...]
"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tensorflow_yolo_b200 import engine as yb  # noqa: E402
import torch  # noqa: E402

ops = [int(a) for a in sys.argv[1:]] or [1, 3]
net, state, stream, shape = bench.build_network(416, "v3")
B = 128
eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
eng.load_weights(stream)
x = torch.rand((B,) + shape, device="cuda")
eng.forward(x); eng.sync()
for op in ops:
    base = eng.time_op(op, B, reps=20)
    print("op", op, "default", eng.op_cfg(op), "%.4f ms" % base)
    rows = []
    for bn, pair, bstat, tma_epi, ksub in itertools.product((32, 64, 128), (0, 1), (0, 1), (0, 1), (1, 3, 9)):
        try:
            eng.set_conv_cfg(op, bn, pair, bstat, tma_epi, ksub)
            ms = min(eng.time_op(op, B, reps=20) for _ in range(2))
            rows.append((ms, bn, pair, bstat, tma_epi, ksub, eng.op_cfg(op)["stages"]))
        except Exception as e:      # configuration not available for this op
            pass
    eng.set_conv_cfg(op, 0)
    for r in sorted(rows)[:8]:
        print("   %.4f ms  bn=%d pair=%d bstat=%d tma_epi=%d ksub=%d stages=%d" % r)
