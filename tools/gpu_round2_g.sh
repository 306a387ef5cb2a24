#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_conv.py -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2g_tests.log
for fp in 1 0; do
YB_FUSE_POOL=$fp timeout 300 python bench.py --net v2voc --batch 64 --steps 20 --warmup 3 --no-cpu-baseline --no-extra --dump-profile gpurun_out/r2g_prof_v2_fp$fp.json > gpurun_out/r2g_bench_v2_fp$fp.json 2> gpurun_out/r2g_err; echo "bench v2 fuse_pool=$fp rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_v2_fp$fp.json'))
print('v2voc b64 value %.0f e2e %.0f ms %.3f kept/img %.1f other_ms %.3f conv_ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['kept_detections_per_image'], d['roofline']['conv_stack']['other_forward_ms_per_step'], d['roofline']['conv_stack']['ms_per_step']))
p=json.load(open('gpurun_out/r2g_prof_v2_fp$fp.json'))
print([(o['layer'],o['kind'],round(o['ms'],4)) for o in p['ops'] if o['kind']!='conv' or o['layer']<4])
PY
done
