#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2b_fused_test.log 2>&1; rc=$?
echo "fused test rc=$rc"; tail -15 gpurun_out/r2b_fused_test.log
if [ $rc -eq 0 ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -8 gpurun_out/r2b_gpu_tests.log
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2b_prof_fused.json > gpurun_out/r2b_bench_fused.json 2> gpurun_out/r2b_bench_fused.err; echo "bench fused rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2b_bench_fused.json'));print('fused', d['value'], d['e2e']['value'], d['ms_per_step']);p=json.load(open('gpurun_out/r2b_prof_fused.json'));print([(o['layer'],round(o['ms'],4)) for o in p['ops'][:6]])"
fi
