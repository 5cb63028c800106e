#!/bin/bash
mkdir -p gpurun_out
timeout 60 build/tmem_ld_layout_probe > gpurun_out/r02f_tmem_ld_layout_probe.txt 2>&1; echo "probe rc=$?"; head -40 gpurun_out/r02f_tmem_ld_layout_probe.txt
for w in 1 2 3; do
  echo "== DATR_WGRAD_WAVES=$w" >> gpurun_out/r02f_wgrad_waves.txt
  DATR_WGRAD_WAVES=$w timeout 300 python -c "
import sys; sys.path.insert(0,'tools'); import bench_linear; bench_linear.wgrad()" 2>&1 >> gpurun_out/r02f_wgrad_waves.txt
done
cut -c1-200 gpurun_out/r02f_wgrad_waves.txt
