#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_e2e.py -x -q > gpurun_out/r2d_e2e.log 2>&1; echo "e2e rc=$?"; tail -12 gpurun_out/r2d_e2e.log
cat gpurun_out/detection_flip_rate.json 2>/dev/null; echo
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench_1gpu.json 2> gpurun_out/r2d_bench_1gpu.err; echo "bench1 rc=$?"; tail -3 gpurun_out/r2d_bench_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_1gpu.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'sustained',d['sustained'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','frac_of_burst','traffic','kernel')})
print('extra',json.dumps(d['extra'])[:1800])
print('cpu',d['cpu_baseline'])
PY
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2d_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_2gpu.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'sustained',d['sustained'])
print('extra',json.dumps(d['extra'])[:1500])
PY
