#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q > gpurun_out/r02o_tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02o_tests.txt | cut -c1-220
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_default.json 2> gpurun_out/r02o_bench_default.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r02o_bench_default.json
