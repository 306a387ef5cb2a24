#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -6 gpurun_out/r2e_gpu_tests.log
cat gpurun_out/detection_flip_rate.json; echo
timeout 600 python tools/cudnn_yardstick.py --out gpurun_out/r2e_cudnn_yardstick.json 2>&1 | tail -32
