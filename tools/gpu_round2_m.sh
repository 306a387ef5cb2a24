#!/bin/bash
for cfg in "YB_SPLIT_K_OVERHEAD=2000 YB_SPLIT_K_GAIN=0.95" "YB_SPLIT_K_OVERHEAD=4000 YB_SPLIT_K_GAIN=0.92" "YB_SPLIT_K_OVERHEAD=0 YB_SPLIT_K_GAIN=0.97"; do
echo "== $cfg"
env $cfg timeout 600 python tools/bench_latency.py --batches 8,16,32,64 --iters 100 --out gpurun_out/r2m_latency.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['batch'], 'eager %.3f graph %.3f split %.3f ms' % (r['eager']['device_ms_per_step'], r['graph']['device_ms_per_step'], r['graph_split_k']['device_ms_per_step']), r.get('split_k_factors'))
"
done
