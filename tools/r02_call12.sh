#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q > gpurun_out/r02k_tests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02k_tests.txt | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_default.json 2> gpurun_out/r02k_bench_default.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r02k_bench_default.json; tail -2 gpurun_out/r02k_bench_default.err | cut -c1-300
DATR_JOINT_DECODER=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r02k_bench_split_decoder.json 2> /dev/null; cut -c1-200 gpurun_out/r02k_bench_split_decoder.json
timeout 900 python bench.py --workload teacher --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r02k_bench_teacher.json 2> /dev/null; cut -c1-200 gpurun_out/r02k_bench_teacher.json
timeout 900 python bench.py --workload dino5 --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r02k_bench_dino5.json 2> /dev/null; cut -c1-200 gpurun_out/r02k_bench_dino5.json
