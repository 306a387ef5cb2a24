#!/bin/bash
# quick loop for the fused kernels: one parity test + role cycle counters
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q -k "fused_pairs or batch_position" > gpurun_out/r2j_fused_test.log 2>&1; rc=$?
echo "fused test rc=$rc"; tail -5 gpurun_out/r2j_fused_test.log
timeout 200 python tools/probes/fused_cycles.py 2>&1 | tail -3
