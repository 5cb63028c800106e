"""Micro-benchmark of the implicit-GEMM 3x3 convolution (+bias+ReLU) vs cuDNN (TF32, NHWC) + separate bias/ReLU at the
ResNet-50 conv2 shapes of a 4-image 1333x800 batch (GPU box only)."""
import os, sys, json
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200.conv import conv3x3_bias_relu

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = PEAKS.get("hbm_gbs", 6650.0)


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


torch.backends.cudnn.allow_tf32 = True
for name, C, H, W, s in [("layer1", 64, 200, 334, 1), ("layer2 first", 128, 200, 334, 2), ("layer2", 128, 100, 167, 1),
                         ("layer3 first", 256, 100, 167, 2), ("layer3", 256, 50, 84, 1), ("layer4 first", 512, 50, 84, 2),
                         ("layer4", 512, 25, 42, 1)]:
    x = torch.randn(4, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = (torch.randn(C, C, 3, 3, device="cuda") / (3 * C ** 0.5)).contiguous(memory_format=torch.channels_last)
    b = torch.randn(C, device="cuda")
    with torch.no_grad():
        t_ours = timeit(lambda: conv3x3_bias_relu(x, w, b, s))
        t_lib = timeit(lambda: F.relu(F.conv2d(x, w, b, stride=s, padding=1)))
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    flops = 2.0 * 4 * Ho * Wo * C * C * 9
    bytes_ = 4.0 * (4 * H * W * C + 9 * C * C + 4 * Ho * Wo * C)
    print(f"{name:13s} C={C:4d} {H}x{W} s={s}: ours {t_ours*1e3:7.1f} us ({flops/t_ours/1e9:6.1f} TF/s, {bytes_/t_ours/1e6:7.1f} GB/s = {bytes_/t_ours/1e6/HBM:5.3f} of HBM)"
          f"   cuDNN+bias+relu {t_lib*1e3:7.1f} us", flush=True)
