"""ncu target: the MSDeformAttn run kernels (strategy 2) and row kernels (strategy 1) on one config-2 encoder call with
run-coherent locations (initial offset pattern + 0.3 px jitter).  Usage under gpurun:
  ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 4 -o gpurun_out/prof python tools/ncu_target_runs.py
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import msda_cases as mc
from test_msda_runs_gpu import coherent_inputs
from datr_b200 import MultiScaleDeformableAttention as MSDA

jit = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
inp = coherent_inputs(2, 8, 4, mc.CFG2_LEVELS, jit, 1)
d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
args = (d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"])
for it in range(2):                 # launches 0-3 warm-up, 4-7 profiled: runs fwd, runs bwd, rows fwd, rows bwd
    for st in (2, 1):
        MSDA.set_strategy(st)
        MSDA.ms_deform_attn_forward(*args, 64)
        MSDA.ms_deform_attn_backward(*args, d["grad_out"], 64)
torch.cuda.synchronize()
print("done")
