#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_conv_backward.py > gpurun_out/r02n_bench_conv_backward.txt 2>&1; cat gpurun_out/r02n_bench_conv_backward.txt | cut -c1-200
