#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both workloads, both arms), micro-benchmark vs the
# reference CUDA kernels, step kernel breakdown, ncu launch list and one full capture of the MSDeformAttn kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_dino.json 2> gpurun_out/bench_dino.err; cat gpurun_out/bench_dino.json; tail -5 gpurun_out/bench_dino.err
timeout 600 python bench.py --workload msda --steps 10 --warmup 3 > gpurun_out/bench_msda.json 2> gpurun_out/bench_msda.err; cat gpurun_out/bench_msda.json; tail -3 gpurun_out/bench_msda.err
timeout 600 python tools/microbench_msda.py > gpurun_out/microbench.txt 2>&1; cat gpurun_out/microbench.txt
timeout 600 python tools/profile_dino.py > gpurun_out/profile_dino.log 2>&1; tail -80 gpurun_out/profile_dino.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_msda.csv python bench.py --workload msda --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 4 -o gpurun_out/prof python tools/ncu_target.py 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_dino_ref.json 2>> gpurun_out/bench_dino.err; cat gpurun_out/bench_dino_ref.json
ls -la gpurun_out
