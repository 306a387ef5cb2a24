"""Short GPU program for ncu captures of the round-2 kernels: two forwards of YOLOv3-416 at batch 128 (fused stem kernels,
the tcgen05 convs) and two dense decodes of [128, 10647, 85] heads (decode_v3_bulk_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from tensorflow_yolo_b200 import engine as yb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
net, state, stream, shape = bench.build_network(416, "v3")
eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
eng.load_weights(stream)
x = torch.rand((B,) + shape, device="cuda", dtype=torch.float32)
for _ in range(2):
    eng.forward(x)
    eng.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
eng.sync()
post = yb.PostProcessor(bench.v3_scales(416), 80, yb.YB_DECODE_V3, max_batch=B)
head = torch.randn((B, 10647, 85), device="cuda", dtype=torch.float32)
for _ in range(2):
    post.run(head, 0.001, 0.6, fetch=False)
post.sync()
print("ok")
