#!/bin/bash
# full GPU suite + bench after the fused-kernel rewrite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/r2k_gpu_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra --dump-profile gpurun_out/r2k_prof.json > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2k_bench.json'));print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks'], d.get('sustained'));p=json.load(open('gpurun_out/r2k_prof.json'));print([(o['layer'],round(o['ms'],4)) for o in p['ops'][:6]])"
