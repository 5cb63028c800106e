#!/bin/bash
# round 2, GPU call 2/3: MSDeformAttn run kernels -- parity tests, micro-benchmark rows vs runs vs reference CUDA
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_runs_gpu.py tests/test_msda_gpu.py tests/test_msda_fused_gpu.py tests/test_ema_gpu.py -q > gpurun_out/r02b_msda_tests.txt 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r02b_msda_tests.txt
timeout 600 python tools/microbench_msda.py > gpurun_out/r02b_msda_microbench.txt 2>&1; echo "microbench rc=$?"; cat gpurun_out/r02b_msda_microbench.txt
