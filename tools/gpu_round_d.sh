#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wgrad_tf32|layernorm256|colsum" -s 4 -c 4 -o gpurun_out/prof_small python tools/ncu_target_small.py > gpurun_out/ncu_small.log 2>&1; tail -2 gpurun_out/ncu_small.log
timeout 300 python tools/bench_linear.py > gpurun_out/bench_linear.txt 2>&1; cat gpurun_out/bench_linear.txt
