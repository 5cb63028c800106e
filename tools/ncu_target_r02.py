"""Short target for ncu: the kernels added in round 2 at the config-2 shapes -- fused self-attention forward / backward
(4 images x 8 heads x 1100 queries), the bf16 FFN GEMMs (88 892 rows: linear1 -> bf16 with the TMA-store epilogue, linear2,
masked dgrad, both weight gradients), the one-launch AdamW and the sine embedding.
  ncu --set full --clock-control none --import-source on -k "regex:attn_|linear_tf32_kernel|wgrad_tf32_kernel|adamw_step|sine_embed" \
      -o gpurun_out/prof python tools/ncu_target_r02.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import attention, linear as dl
from datr_b200.models.dino.utils import gen_sineembed_for_position
from test_attention_gpu import dn_mask

dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, H, T, C = 4, 8, 1100, 256
qk = torch.randn(N, T, 2 * C, device=dev, requires_grad=True)
v = torch.randn(N, T, C, device=dev, requires_grad=True)
blocked = dn_mask(T, 200, 10, None).to(dev)
bits = attention.pack_mask(blocked, T, dev)
go = torch.randn(N, T, C, device=dev)
M, d, dff = 88892, 256, 2048
x = torch.randn(M, d, device=dev)
xb = x.bfloat16()
w1b = (torch.randn(dff, d, device=dev) / 16).bfloat16(); w2b = (torch.randn(d, dff, device=dev) / 45).bfloat16()
b1 = torch.zeros(dff, device=dev); b2 = torch.zeros(d, device=dev)
gy = torch.randn(M, d, device=dev); gb = gy.bfloat16()
w2t, w1t = w2b.t().contiguous(), w1b.t().contiguous()
lin = torch.nn.Linear(1000, 1000).to(dev)
from datr_b200.parallel import FlatGradients
from datr_b200.optim import FlatAdamW
big = torch.nn.Linear(4096, 4096).to(dev)
grads = FlatGradients(big)
opt = FlatAdamW([{"params": list(big.parameters()), "lr": 1e-4}], grads, weight_decay=1e-4)
grads.flat.normal_()
pos = torch.rand(N, T, 4, device=dev)
for _ in range(2):
    attention.fused_self_attention(qk, v, H, blocked, bits=bits).backward(go)
    h = dl._launch_bf16(xb, w1b, b1, None, 1, True)
    dl._launch_bf16(h, w2b, b2, x, 0, False)
    dz1 = dl._launch_bf16(gb, w2t, None, h, 3, True, residual_bf16=True)
    dl._launch_bf16(dz1, w1t, None, gy, 0, False)
    dl._wgrad_bf16(gb, h, True)
    dl._wgrad_bf16(dz1, xb, True)
    opt.clip_and_step(0.1)
    gen_sineembed_for_position(pos)
torch.cuda.synchronize()
print("done")
