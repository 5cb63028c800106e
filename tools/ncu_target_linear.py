"""Short ncu target: the tcgen05 linear kernel at the DINO-4scale encoder shapes (M = 44446).
  ncu --set full --clock-control none --import-source on -k regex:linear_tf32 -s 5 -c 5 -o gpurun_out/prof_linear python tools/ncu_target_linear.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200 import linear as dl

dl.set_mode("tf32")
M = 44446
shapes = [(256, 256, False, False), (128, 256, False, False), (2048, 256, True, False), (256, 2048, False, True), (256, 256, False, True)]
data = []
for N, K, relu, res in shapes:
    data.append((torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(N, device="cuda"),
                 torch.randn(M, N, device="cuda") if res else None, relu))
for _ in range(2):
    for x, w, b, r, relu in data:
        dl.linear(x, w, b, relu=relu, residual=r)
torch.cuda.synchronize()
print("done")
