import sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from tensorflow_yolo_b200 import engine as yb
import torch
net, state, stream, shape = bench.build_network(416, "v3")
B=128
eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=B, device=0)
eng.load_weights(stream)
x8 = torch.randint(0,256,(B,)+shape,dtype=torch.uint8,device='cuda')
xf = (x8.float()/255.)
for name,x in (('f32',xf),('u8',x8)):
    best=1e9
    for _ in range(5):
        p = eng.profile(x)
        best=min(best,p[0][1])
    print(name,'first conv ms',best)
