#!/bin/bash
# MSDeformAttn kernel iteration: parity tests, micro-benchmark vs the reference CUDA kernels, one ncu full capture.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py -x -q > gpurun_out/pytest_msda.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_msda.log
tail -8 gpurun_out/pytest_msda.log
timeout 600 python tools/microbench_msda.py > gpurun_out/microbench.txt 2>&1; grep -v " ref " gpurun_out/microbench.txt
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 4 -o gpurun_out/prof python tools/ncu_target.py 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
fi
