#!/bin/bash
# A/B on one box: env var settings alternate, bench.py value/e2e per run.  usage: gpu_ab.sh "<envA>" "<envB>" [reps]
A="$1"; B="$2"; R=${3:-2}
for i in $(seq 1 $R); do
  for cfg in "$A" "$B"; do
    env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
    python -c "import json,sys;d=json.load(open('gpurun_out/ab_tmp.json'));print(sys.argv[1], 'value %.0f e2e %.0f ms %.3f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']))" "$cfg"
  done
done
