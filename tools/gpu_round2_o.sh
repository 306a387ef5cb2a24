#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_conv.py -x -q -k "max_pool or v2" > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2o_tests.log
for fp in 1 0; do
YB_FUSE_POOL=$fp timeout 300 python bench.py --net v2voc --batch 64 --steps 20 --no-cpu-baseline --no-extra --sustain-seconds 0 --dump-profile gpurun_out/r2o_prof_fp$fp.json > gpurun_out/r2o_bench_fp$fp.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r2o_bench_fp$fp.json'));print('fuse_pool=$fp value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']);p=json.load(open('gpurun_out/r2o_prof_fp$fp.json'));print([(o['layer'],o['kind'],round(o['ms'],4)) for o in p['ops'][:8]])"
done
