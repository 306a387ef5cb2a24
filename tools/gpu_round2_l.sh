#!/bin/bash
# split-K latency mode: parity test + small-batch latency with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "split_k or autotune or cuda_graph" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2l_tests.log
timeout 900 python tools/bench_latency.py --batches 1,8,16,32,64 --iters 100 --out gpurun_out/r2l_latency.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['batch'], 'eager %.3f graph %.3f split %.3f ms' % (r['eager']['device_ms_per_step'], r['graph']['device_ms_per_step'], r['graph_split_k']['device_ms_per_step']), r.get('split_k_factors'))
"
