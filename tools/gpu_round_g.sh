#!/bin/bash
# ncu launch lists (gpu__time_duration only): the msda workload of bench.py, and the hand-written kernels of the default
# (dino) workload in eager mode (kernel-name filter keeps the other ~8 000 launches per step out of the profiler)
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_msda.csv python bench.py --workload msda --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_msda.log 2>&1; tail -1 gpurun_out/launches_msda.csv | cut -c1-120
DATR_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:msda_|linear_tf32|wgrad_tf32|layernorm256|softmax_|conv3x3|colsum|zero_masked" --launch-skip 3000 -c 3200 --csv --log-file gpurun_out/launches_dino_handwritten.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_dino.log 2>&1; tail -1 gpurun_out/launches_dino_handwritten.csv | cut -c1-120
