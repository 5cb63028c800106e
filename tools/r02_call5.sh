#!/bin/bash
# round 2, GPU call 5: forward-kernel occupancy variants, new bench workloads, reference arm on the box's host cores
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py tests/test_ema_gpu.py -q > gpurun_out/r02d_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02d_tests.txt
for v in 0 1 2 3 4 5 6; do
  echo "== DATR_MSDA_FWD_VARIANT=$v" >> gpurun_out/r02d_fwd_variants.txt
  DATR_MSDA_FWD_VARIANT=$v timeout 300 python tools/microbench_msda.py --no-ref --fused --cases=cfg2_enc,cfg2_enc_init,cfg2_dec1100,cfg4_enc >> gpurun_out/r02d_fwd_variants.txt 2>&1
done
cat gpurun_out/r02d_fwd_variants.txt | grep -v "^$" | cut -c1-150
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_default.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02d_bench_default.json; tail -3 gpurun_out/r02d_bench_default.err
timeout 900 python bench.py --workload dino5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_dino5.json 2> gpurun_out/r02d_bench_dino5.err; echo "dino5 rc=$?"; cut -c1-400 gpurun_out/r02d_bench_dino5.json; tail -3 gpurun_out/r02d_bench_dino5.err
timeout 900 python bench.py --workload teacher --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_teacher.json 2> gpurun_out/r02d_bench_teacher.err; echo "teacher rc=$?"; cut -c1-400 gpurun_out/r02d_bench_teacher.json; tail -3 gpurun_out/r02d_bench_teacher.err
