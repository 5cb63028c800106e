#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 8 9; do
  echo "== DATR_LINEAR_DEBUG=$d" >> gpurun_out/r02e_linear_epilogue_experiment.txt
  DATR_LINEAR_DEBUG=$d timeout 300 python -c "
import sys; sys.path.insert(0,'tools'); import bench_linear; bench_linear.main()" 2>&1 | sed 's/cuBLAS-fp32.*//' >> gpurun_out/r02e_linear_epilogue_experiment.txt
done
cat gpurun_out/r02e_linear_epilogue_experiment.txt | cut -c1-170
