#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 4 -f -o gpurun_out/r02c_msda_runs python tools/ncu_target_runs.py > gpurun_out/r02c_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02c_ncu.log
