#!/bin/bash
# round 2, GPU call 1: probes left over from round 1 + the reference model on the B200 (baseline + full-size parity)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout 90 build/gemm_2cta_probe > gpurun_out/r02a_gemm_2cta_probe.txt 2>&1; echo "2cta rc=$?"
timeout 180 build/msda_tile_probes > gpurun_out/r02a_msda_tile_probes.txt 2>&1; echo "tile probes rc=$?"
timeout 1200 python tools/bench_reference_gpu.py --config 4scale --steps 5 --out gpurun_out/r02a_reference_gpu_4scale.json > gpurun_out/r02a_reference_gpu_4scale.log 2>&1; echo "ref 4scale rc=$?"
grep -E "^\[(parity|reference|ours)" gpurun_out/r02a_reference_gpu_4scale.log | cut -c1-600
timeout 900 python tools/bench_reference_gpu.py --config 5scale --steps 3 --out gpurun_out/r02a_reference_gpu_5scale.json > gpurun_out/r02a_reference_gpu_5scale.log 2>&1; echo "ref 5scale rc=$?"
grep -E "^\[(parity|reference|ours)" gpurun_out/r02a_reference_gpu_5scale.log | cut -c1-600
tail -5 gpurun_out/r02a_gemm_2cta_probe.txt; tail -12 gpurun_out/r02a_msda_tile_probes.txt
