#!/bin/bash
# fused stem kernels, TMA-fed 16-warp version: parity tests, role cycle counters, bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2i_fused_test.log 2>&1; rc=$?
echo "fused test rc=$rc"; tail -15 gpurun_out/r2i_fused_test.log
if [ $rc -eq 0 ]; then
  timeout 200 python tools/probes/fused_cycles.py 2>&1 | tail -3
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra --dump-profile gpurun_out/r2i_prof.json > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2i_bench.json'));print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks']);p=json.load(open('gpurun_out/r2i_prof.json'));print([(o['layer'],round(o['ms'],4)) for o in p['ops'][:6]])"
fi
