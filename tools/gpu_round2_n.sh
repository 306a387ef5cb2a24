#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/cudnn_yardstick.py > gpurun_out/r02f_cudnn.log 2>&1; echo "yardstick rc=$?"; tail -4 gpurun_out/r02f_cudnn.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 --no-autotune > gpurun_out/r02f_b2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 --no-autotune > gpurun_out/r02f_ncu_list.log 2>&1; echo "list rc=$?"
