"""Short ncu target: implicit-GEMM 3x3 convolution at three ResNet-50 conv2 shapes (4 images of 1333x800).
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tf32 -s 3 -c 3 -o gpurun_out/prof_conv python tools/ncu_target_conv.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200.conv import conv3x3_bias_relu

data = []
for C, H, W, s in [(64, 200, 334, 1), (128, 100, 167, 1), (256, 50, 84, 1)]:
    x = torch.randn(4, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = (torch.randn(C, C, 3, 3, device="cuda") / (3 * C ** 0.5)).contiguous(memory_format=torch.channels_last)
    data.append((x, w, torch.randn(C, device="cuda"), s))
with torch.no_grad():
    for _ in range(2):
        for x, w, b, s in data:
            conv3x3_bias_relu(x, w, b, s)
torch.cuda.synchronize()
print("done")
