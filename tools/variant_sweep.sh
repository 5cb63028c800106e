for v in 0 1; do echo "== variant $v"; DATR_MSDA_VARIANT=$v python tools/microbench_msda.py 2>&1 | grep -E "ours  cold"; done
DATR_MSDA_VARIANT=1 python -m pytest tests/test_msda_gpu.py -x -q 2>&1 | tail -2
