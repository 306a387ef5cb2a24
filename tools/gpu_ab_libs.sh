#!/bin/bash
# same-box A/B of library builds: tools/ab/lib_<name>.so are swapped in turn.  usage: gpu_ab_libs.sh "base cur" [reps]
mkdir -p gpurun_out
cp tensorflow_yolo_b200/libyolo_b200.so /tmp/lib_keep.so
R=${2:-3}
for i in $(seq 1 $R); do
  for name in $1; do
    cp tools/ab/lib_$name.so tensorflow_yolo_b200/libyolo_b200.so
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
    python -c "import json,sys;d=json.load(open('gpurun_out/ab_tmp.json'));print(sys.argv[1], 'value %.0f e2e %.0f ms %.3f sustained %.0f clk %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['sustained']['value'], d['clocks']['sm_mhz']))" "$name"
  done
done
cp /tmp/lib_keep.so tensorflow_yolo_b200/libyolo_b200.so
