"""Short target for ncu: the config-2 encoder call (N=2, S=Lq=22223) of the fused MSDeformAttn forward on fp32 rows and
on bf16 pair rows, the pack kernel, and the fused backward with the vector-reduction scatter and with the TMA scatter.
  ncu --set full --clock-control none --import-source on -k "regex:msda_" -s 5 -c 5 -o gpurun_out/prof python tools/ncu_target_pairs.py
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import msda_cases as mc
from datr_b200 import MultiScaleDeformableAttention as MSDA
from datr_b200 import native

levels = mc.CFG2_LEVELS
N, L = 2, len(levels)
inp = mc.coherent_inputs(N, 8, 4, levels, 0.3, 1)          # the initial offset pattern + 0.3 px jitter
d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
S = d["value"].shape[1]
refp = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(
    mc.encoder_reference_points(levels)[None, :, None, :], (N, S, L, 2)), dtype=np.float32)).cuda()
wh = d["shapes"].flip(-1).float()
off = ((d["loc"] - refp[:, :, None, :, None, :]) * wh[None, None, None, :, None, :]).contiguous()
lg = d["attn"].flatten(3).log().contiguous()
fa = (d["value"], d["shapes"], d["level_start"], off, lg, refp)
lib = native.lib()
for _ in range(2):
    MSDA.ms_deform_attn_fused_forward(*fa)
    pairs = MSDA.pack_value_pairs(d["value"], d["shapes"], d["level_start"], torch.bfloat16)
    MSDA.ms_deform_attn_fused_forward(*fa, pairs=pairs)
    lib.datr_msda_set_backward_stages(0)
    MSDA.ms_deform_attn_fused_backward(*fa, d["grad_out"])
    lib.datr_msda_set_backward_stages(2)
    MSDA.ms_deform_attn_fused_backward(*fa, d["grad_out"])
torch.cuda.synchronize()
print("done")
