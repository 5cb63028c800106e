"""Short ncu target: weight-gradient, LayerNorm-backward and column-sum kernels at the DINO-4scale encoder shapes.
  ncu --set full --clock-control none --import-source on -k regex:'wgrad_tf32|layernorm256|colsum' -s 4 -c 4 -o gpurun_out/prof_small python tools/ncu_target_small.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200.linear import _wgrad, _colsum
from datr_b200.layernorm import layer_norm

M = 44446
dz1 = torch.randn(M, 256, device="cuda"); x1 = torch.randn(M, 256, device="cuda")
dz2 = torch.randn(M, 2048, device="cuda"); y2 = torch.randn(M, 2048, device="cuda")
x = torch.randn(2, 22223, 256, device="cuda", requires_grad=True); g = torch.randn(2, 22223, 256, device="cuda")
norm = torch.nn.LayerNorm(256).cuda()
for _ in range(2):
    _wgrad(dz1, x1, True)
    _wgrad(dz2, x1, True)
    torch.autograd.grad(layer_norm(norm, x), (x, norm.weight, norm.bias), g)
    _colsum(dz2, y2)
torch.cuda.synchronize()
print("done")
