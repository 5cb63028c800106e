#!/bin/bash
# new kernels first (fused MSDeformAttn, padding mask, LayerNorm forward), then the whole GPU suite and the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_fused_gpu.py tests/test_layernorm_gpu.py -x -q > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log; tail -15 gpurun_out/pytest_new.log | cut -c1-220
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log | cut -c1-220
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err | cut -c1-200
