#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q > gpurun_out/r02i_tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02i_tests.txt | cut -c1-200
timeout 600 python tools/profile_dino.py > /dev/null 2>&1; cp gpurun_out/dino_step_kernels.txt gpurun_out/r02i_dino_step_kernels_graphs.txt
timeout 600 python tools/gap_report.py > /dev/null 2>&1; cp gpurun_out/dino_step_gaps.txt gpurun_out/r02i_dino_step_gaps.txt 2>/dev/null
head -75 gpurun_out/r02i_dino_step_kernels_graphs.txt | cut -c1-200
head -30 gpurun_out/r02i_dino_step_gaps.txt | cut -c1-200
