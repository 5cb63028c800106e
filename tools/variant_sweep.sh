#!/bin/bash
# tuning helper: times the MSDeformAttn kernel variants (DATR_MSDA_VARIANT) at the cfg2 shapes
for v in 9 1 2 3 4; do
  echo "== variant $v"
  DATR_MSDA_VARIANT=$v python tools/microbench_msda.py 2>&1 | grep -E "cfg2_enc |cfg2_dec" | grep "ours  cold"
done
DATR_MSDA_VARIANT=4 python -m pytest tests/test_msda_gpu.py -x -q 2>&1 | tail -2
