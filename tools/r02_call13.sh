#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload teacher --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_teacher.json 2> gpurun_out/r02l_bench_teacher.err; echo "teacher rc=$?"; cut -c1-200 gpurun_out/r02l_bench_teacher.json; tail -3 gpurun_out/r02l_bench_teacher.err | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py -q > gpurun_out/r02l_tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02l_tests.txt | cut -c1-200
