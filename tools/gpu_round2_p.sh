#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fused.py -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2p_tests.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra --dump-profile gpurun_out/r2p_prof.json > gpurun_out/r2p_bench.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r2p_bench.json'));print('value %.0f e2e %.0f ms %.3f sust %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['value']), d['clocks']['sm_mhz'])"
done
python tools/roofline_table.py gpurun_out/r2p_prof.json | sed -n 6,20p | cut -c1-120
