"""GPU probe: in-kernel cycle counters of the fused stem kernels (conv_fused.cuh), per tile of thread 0 of every CTA."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from tensorflow_yolo_b200 import engine as yb

NAMES = ("prod_stage", "prod_wait_a", "prod_conv", "prod_total", "mma_wait_a", "mma_wait_acc", "mma_total", "epi_wait_mma", "epi_wait_res", "epi_total", "tiles")

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    net, state, stream, shape = bench.build_network(416, "v3")
    eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
    eng.load_weights(stream)
    out = {}
    for op in (1, 3):
        ms = eng.time_op(op, B, reps=10)
        eng.set_option("cycles", 1)
        eng.read_cycles(reset=True)
        eng.time_op(op, B, reps=1)      # 1 untimed + 1 timed launch
        import numpy as np, ctypes
        from tensorflow_yolo_b200 import _lib
        arr = np.zeros(16, dtype=np.uint64)
        _lib.check(_lib.lib().yb_engine_read_cycles(eng._h, arr.ctypes.data, 1))
        eng.read_cycles_raw = [int(v) for v in arr]
        eng.set_option("cycles", 0)
        vals = eng.read_cycles_raw
        tiles = max(vals[10], 1)
        out["op%d" % op] = {"ms": ms, "cycles_per_tile": {n: v / tiles for n, v in zip(NAMES[:10], vals[:10])}, "tiles_counted": tiles}
        print(op, json.dumps(out["op%d" % op]))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused_cycles.json"), "w"), indent=1)

main()
