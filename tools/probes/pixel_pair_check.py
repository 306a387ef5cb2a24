import os, sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import helpers
from tensorflow_yolo_b200 import engine, synth
shape = (128, 160, 3)
net, topo, stream = helpers.build_v3(shape, 80, seed=4)
x = synth.images(3, shape[0], shape[1], seed=5)
outs = {}
for v in ("1", "0"):
    os.environ["YB_PIXEL_PAIRS"] = v
    os.environ["YB_KEEP_ALL"] = "1"
    eng = engine.Engine(net[0]._yb_state.plan(), shape, 80, engine.YB_DECODE_V3, max_batch=3)
    eng.load_weights(stream)
    eng.forward(x)
    outs[v] = (eng.read_layer(5)[0], eng.read_output())
    print(v, [ (i, eng.op_cfg(i)) for i in (3,)])
    eng.close()
a, b = outs["1"], outs["0"]
print("layer5 identical:", np.array_equal(a[0], b[0]), "max abs diff", float(np.abs(a[0] - b[0]).max()))
print("output identical:", np.array_equal(a[1], b[1]), "max abs diff", float(np.abs(a[1] - b[1]).max()))
