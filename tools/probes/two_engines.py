"""GPU probe: a small per-GPU batch as ONE engine (one stream, 75 dependent kernels) against TWO engines of half the batch on
two streams launched alternately from one host thread -- do the two dependent chains fill each other's tails?"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from tensorflow_yolo_b200 import engine as yb

def make(n, state, stream, shape):
    eng = yb.Engine(state.plan(), shape, 80, yb.YB_DECODE_V3, max_batch=n, device=0)
    eng.load_weights(stream)
    eng.autotune(n, reps=3)
    eng.set_option("graph", 1)
    return eng

def main():
    iters = 200
    net, state, stream, shape = bench.build_network(416, "v3")
    out = []
    for B in [int(a) for a in (sys.argv[1:] or ["16", "32", "8"])]:
        x = torch.rand((B,) + shape, device="cuda", dtype=torch.float32)
        one = make(B, state, stream, shape)
        def step1():
            one.forward(x); one.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
        for _ in range(10): step1()
        one.sync(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters): step1()
        one.sync()
        t_one = (time.perf_counter() - t0) * 1e3 / iters
        one.close()
        row = {"batch": B, "one_engine_ms": t_one}
        for parts in (2, 4):
            if B % parts: continue
            h = B // parts
            engs = [make(h, state, stream, shape) for _ in range(parts)]
            xs = [x[i * h:(i + 1) * h].contiguous() for i in range(parts)]
            def stepn():
                for e, xi in zip(engs, xs):
                    e.forward(xi); e.detect_async(bench.THRESHOLD, bench.IOU_THRESHOLD)
            for _ in range(10): stepn()
            for e in engs: e.sync()
            t0 = time.perf_counter()
            for _ in range(iters): stepn()
            for e in engs: e.sync()
            row["%d_engines_ms" % parts] = (time.perf_counter() - t0) * 1e3 / iters
            for e in engs: e.close()
        print(json.dumps(row), flush=True)
        out.append(row)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "two_engines.json"), "w"), indent=1)

main()
