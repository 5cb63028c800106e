#!/bin/bash
# profiles of the round's final state: ncu --set full of the fused MSDeformAttn / LayerNorm forward / padding-mask kernels,
# kernel table of a graph-replayed step, bench lines (default, msda workload, reference arm)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:msda_|layernorm256_fwd|zero_masked" -s 6 -c 6 -o gpurun_out/prof_fused -f python tools/ncu_target_fused.py > gpurun_out/ncu_fused.log 2>&1; tail -2 gpurun_out/ncu_fused.log
timeout 300 python tools/profile_dino.py > gpurun_out/profile_dino.log 2>&1; head -3 gpurun_out/dino_step_kernels.txt
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
timeout 600 python bench.py --workload msda --steps 10 --warmup 3 > gpurun_out/bench_msda.json 2> gpurun_out/bench_msda.err; cut -c1-300 gpurun_out/bench_msda.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
