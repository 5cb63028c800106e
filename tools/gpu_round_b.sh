#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tf32 -s 5 -c 5 -o gpurun_out/prof_linear python tools/ncu_target_linear.py > gpurun_out/ncu_linear.log 2>&1; tail -2 gpurun_out/ncu_linear.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_dino_2gpu.json 2> gpurun_out/bench_dino_2gpu.err; cat gpurun_out/bench_dino_2gpu.json; tail -5 gpurun_out/bench_dino_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload msda --steps 5 --warmup 3 > gpurun_out/bench_msda_2gpu.json 2>> gpurun_out/bench_dino_2gpu.err; cat gpurun_out/bench_msda_2gpu.json
