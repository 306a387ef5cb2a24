"""Short GPU program for ncu captures of the Darknet-19 kernels: two forwards of YOLOv2-VOC 416 at batch 64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from tensorflow_yolo_b200 import engine as yb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net, state, stream, shape = bench.build_network(416, "v2voc")
eng = yb.Engine(state.plan(), shape, 20, yb.YB_DECODE_V2, max_batch=B, device=0)
eng.load_weights(stream)
x = torch.rand((B,) + shape, device="cuda", dtype=torch.float32)
for _ in range(2):
    eng.forward(x)
eng.sync()
print("ok")
