#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_linear_gpu.py -q > gpurun_out/r02h_tests.txt 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02h_tests.txt | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_default.json 2> gpurun_out/r02h_bench_default.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r02h_bench_default.json; tail -2 gpurun_out/r02h_bench_default.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-260 gpurun_out/r02h_bench_2gpu.json; tail -2 gpurun_out/r02h_bench_2gpu.err | cut -c1-300
