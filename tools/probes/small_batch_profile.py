"""GPU probe: per-op CUDA-event times of the forward at a small batch (eager launches, PDL off by the events) next to the
graph-replayed step time, to see which layers bound the latency configuration."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from tensorflow_yolo_b200 import engine as yb, plan as yplan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net, state, stream, shape = bench.build_network(416, "v3")
eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
eng.load_weights(stream)
eng.autotune(B, reps=5)
x = torch.rand((B,) + shape, device="cuda", dtype=torch.float32)
def step():
    eng.forward(x); eng.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
for _ in range(5): step()
eng.sync(); eng.mark(0)
for _ in range(100): step()
eng.mark(1); eng.sync()
step_ms = eng.elapsed_ms(0, 1) / 100
eng.set_option("graph", 0)
eng.profiling(True)
for _ in range(20): step()
eng.sync()
prof, n_fwd = eng.profile_read()
eng.profiling(False)
rows = []
specs = state.graph.specs
for op_i, (layer, ms) in enumerate(prof):
    info = eng.op_info(op_i); cfg = eng.op_cfg(op_i); sp = specs[layer]
    rows.append({"op": op_i, "layer": layer, "out": list(sp.shape), "k": sp.ksize, "s": sp.stride, "path": info["path"], "cfg": cfg, "us": 1e3 * ms / n_fwd,
                 "gflop": info["flops_per_image"] * B / 1e9})
tot = sum(r["us"] for r in rows)
print("batch", B, "graph step ms", step_ms, "sum of per-op events us", tot)
for r in sorted(rows, key=lambda r: -r["us"])[:28]:
    print("op %3d L%3d %-16s k%d s%d path %d bn %3d pair %d st %2d ksub %d splitk %d  %7.1f us  %6.0f TF/s" % (r["op"], r["layer"], r["out"], r["k"], r["s"], r["path"], r["cfg"]["bn"], r["cfg"]["pair"], r["cfg"]["stages"], r["cfg"]["ksub"], r["cfg"].get("splitk", 1), r["us"], r["gflop"] / max(r["us"], 1e-3)))
json.dump({"batch": B, "graph_step_ms": step_ms, "ops": rows}, open(os.path.join(ROOT, "gpurun_out", "small_batch_profile_b%d%s.json" % (B, "_splitk" if os.environ.get("YB_SPLIT_K") else "")), "w"), indent=0)
