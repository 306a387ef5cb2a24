#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post.py -x -q > gpurun_out/r2f_post.log 2>&1; echo "post rc=$?"; tail -6 gpurun_out/r2f_post.log
timeout 300 python tools/bench_post.py --out gpurun_out/r2f_post_c5.json 2>&1 | tail -3
