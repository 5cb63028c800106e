#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 3; do
  echo "== DATR_MSDA_BWD_MERGE=$m" >> gpurun_out/r02p_msda_bwd_merge.txt
  DATR_MSDA_BWD_MERGE=$m timeout 300 python tools/microbench_msda.py --no-ref --fused --cases=cfg2_enc,cfg2_enc_init,cfg2_enc_init+0.3px,cfg2_enc_uniform,cfg4_enc >> gpurun_out/r02p_msda_bwd_merge.txt 2>&1
done
cut -c1-150 gpurun_out/r02p_msda_bwd_merge.txt
DATR_MSDA_BWD_MERGE=1 timeout 600 python -m pytest tests/test_msda_gpu.py tests/test_msda_fused_gpu.py -q -x 2>&1 | tail -3
