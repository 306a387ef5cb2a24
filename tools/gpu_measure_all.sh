#!/bin/bash
# end-of-round measurement pass: GPU tests, bench lines of every config, per-op table, ncu launch list + full captures
mkdir -p gpurun_out
T=${1:-r02f}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --dump-profile gpurun_out/${T}_prof.json > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${T}_bench_1gpu.json'));print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks'], d.get('sustained'))"
timeout 300 python bench.py --net v2voc --batch 64 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_config2_v2voc_b64.json 2>/dev/null; echo "c2 rc=$?"
timeout 300 python bench.py --net v2coco --batch 1 --steps 200 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/${T}_bench_config1_v2coco_b1.json 2>/dev/null; echo "c1 rc=$?"
timeout 300 python bench.py --size 608 --batch 64 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_config4_v3_608_b64.json 2>/dev/null; echo "c4 rc=$?"
timeout 300 python tools/bench_post.py > gpurun_out/${T}_post.log 2>&1; cp gpurun_out/post_c5.json gpurun_out/${T}_post_c5.json 2>/dev/null; tail -2 gpurun_out/${T}_post.log
timeout 200 python tools/probes/fused_cycles.py > gpurun_out/${T}_fused_cycles.log 2>&1; cp gpurun_out/fused_cycles.json gpurun_out/${T}_fused_cycles.json
# ncu: launch list of a 2-step bench (after the same command ran clean above), then full captures
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/${T}_b2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/${T}_ncu_list.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_fused|block_fused|decode_v3_bulk|sort_nms' -c 8 -o gpurun_out/${T}_ncu_fused_post python tools/probes/ncu_targets.py > gpurun_out/${T}_ncu_full1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_persist' --launch-skip 8 -c 6 -o gpurun_out/${T}_ncu_conv_tc python tools/probes/ncu_targets.py > gpurun_out/${T}_ncu_full2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/${T}_*.ncu-rep
