#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2c_fused_test.log 2>&1; rc=$?
echo "fused test rc=$rc"; tail -12 gpurun_out/r2c_fused_test.log
if [ $rc -eq 0 ]; then
  timeout 200 python tools/probes/fused_cycles.py 2>&1 | tail -3
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2c_prof_fused.json > gpurun_out/r2c_bench_fused.json 2> gpurun_out/r2c_bench_fused.err; echo "bench fused rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2c_bench_fused.json'));print('fused', d['value'], d['e2e']['value'], d['ms_per_step']);p=json.load(open('gpurun_out/r2c_prof_fused.json'));print([(o['layer'],round(o['ms'],4)) for o in p['ops'][:6]])"
fi
