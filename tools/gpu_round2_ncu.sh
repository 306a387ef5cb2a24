#!/bin/bash
mkdir -p gpurun_out
python tools/probes/ncu_targets.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'stem_fused|block_fused|decode_v3_bulk' -c 6 -o gpurun_out/r02_prof_fused_decode python tools/probes/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_launches.csv python tools/probes/ncu_targets.py 32 > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
ls -la gpurun_out/*.ncu-rep
