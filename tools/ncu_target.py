"""Short target for ncu: a few MSDeformAttn encoder/decoder forward+backward calls at the
BASELINE.json configs[1] shapes (N=2, S=22223).  Usage under gpurun:
  ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 4 -o gpurun_out/prof python tools/ncu_target.py
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench
from datr_b200 import MultiScaleDeformableAttention as MSDA

dev = torch.device("cuda", 0)
S, _ = bench.step_plan()
shapes = torch.tensor(bench.CFG2_LEVELS, dtype=torch.int64, device=dev)
hw = shapes[:, 0] * shapes[:, 1]
lstart = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sets = [bench.synth_call(2, S, S, "enc", 1, dev), bench.synth_call(2, S, 1100, "dec", 2, dev)]
for _ in range(iters):
    for c in sets:
        MSDA.ms_deform_attn_forward(c["value"], shapes, lstart, c["loc"], c["attn"], 64)
        MSDA.ms_deform_attn_backward(c["value"], shapes, lstart, c["loc"], c["attn"], c["grad_out"], 64)
torch.cuda.synchronize()
print("done")
