#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_model_gpu.py tests/test_conv_gpu.py -q -x > gpurun_out/r02g_tests.txt 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02g_tests.txt
timeout 600 python tools/bench_linear.py > gpurun_out/r02g_bench_linear.txt 2>&1; cut -c1-200 gpurun_out/r02g_bench_linear.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02g_bench_default.json
