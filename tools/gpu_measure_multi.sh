#!/bin/bash
# multi-GPU pass: N = number of visible GPUs.  2-device launcher test, weak and strong scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -k "shards_over_devices" > gpurun_out/r02f_multi_tests_${N}gpu.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r02f_multi_tests_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r02f_bench_${N}gpu.json 2> gpurun_out/r02f_bench_${N}gpu.err; echo "weak rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --scaling strong --no-cpu-baseline --no-extra > gpurun_out/r02f_bench_${N}gpu_strong.json 2> gpurun_out/r02f_bench_${N}gpu_strong.err; echo "strong rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r02f_bench_${N}gpu.json", "gpurun_out/r02f_bench_${N}gpu_strong.json"):
    try:
        d = json.load(open(f))
        print(f, d["n_gpus"], d["scaling"], "value %.0f e2e %.0f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"].get("batch_per_gpu"), (d.get("extra") or {}).get("config3_strong"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
