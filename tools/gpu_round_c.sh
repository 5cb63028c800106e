#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tf32 -s 3 -c 3 -o gpurun_out/prof_conv python tools/ncu_target_conv.py > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tf32 -s 5 -c 5 -o gpurun_out/prof_linear python tools/ncu_target_linear.py > gpurun_out/ncu_linear.log 2>&1; tail -2 gpurun_out/ncu_linear.log
DATR_GRAPHS=0 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 45000 --csv --log-file gpurun_out/launches_dino.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log; ls -la gpurun_out/launches_dino.csv
timeout 300 python tools/bench_linear.py > gpurun_out/bench_linear.txt 2>&1; cat gpurun_out/bench_linear.txt
timeout 300 python tools/bench_conv.py > gpurun_out/bench_conv.txt 2>&1
timeout 300 python tools/bench_small_kernels.py > gpurun_out/bench_small.txt 2>&1
