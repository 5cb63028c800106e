#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --workload dino --steps 5 --warmup 3 > gpurun_out/bench_dino.json 2> gpurun_out/bench_dino.err; cat gpurun_out/bench_dino.json; tail -5 gpurun_out/bench_dino.err
timeout 600 python bench.py --workload dino --impl reference --steps 1 --warmup 0 > gpurun_out/bench_dino_ref.json 2>> gpurun_out/bench_dino.err; cat gpurun_out/bench_dino_ref.json; tail -3 gpurun_out/bench_dino.err
