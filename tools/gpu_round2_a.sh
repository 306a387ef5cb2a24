#!/bin/bash
# first GPU call of round 2: fused-kernel parity, whole GPU suite, bench A/B (fused / unfused), two-lane probe
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2a_fused_test.log 2>&1; rc=$?
echo "fused test rc=$rc"; tail -15 gpurun_out/r2a_fused_test.log
if [ $rc -eq 0 ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/r2a_gpu_tests.log
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2a_prof_fused.json > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err; echo "bench fused rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2a_bench_fused.json'));print('fused', d['value'], d['e2e']['value'], d['ms_per_step'])"
else
  export YB_FUSE_STEM=0 YB_FUSE_BLOCK=0
fi
YB_FUSE_STEM=0 YB_FUSE_BLOCK=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r2a_prof_unfused.json > gpurun_out/r2a_bench_unfused.json 2> gpurun_out/r2a_bench_unfused.err; echo "bench unfused rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2a_bench_unfused.json'));print('unfused', d['value'], d['e2e']['value'], d['ms_per_step'])"
timeout 400 python tools/probes/two_lane.py --out gpurun_out/r2a_two_lane.json 2>&1 | tail -5
