"""Micro-benchmark of the 3x3 convolution backward at the shapes of one DINO-4scale DA step (4 images, 1333x800):
input gradient (ours = the forward kernel on the rotated filter, stride 1 only) and weight gradient (ours =
datr_conv3x3_wgrad_nhwc_tf32) against ATen's convolution_backward (cuDNN, TF32, NHWC).  GPU box only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200 import native
from datr_b200.conv import _launch


def timeit(fn, iters=15):
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


def own_wgrad(gz, x, w, stride):
    gw = torch.empty_like(w)
    n, cin, h, wd = x.shape
    rc = native.lib().datr_conv3x3_wgrad_nhwc_tf32(gz.data_ptr(), x.data_ptr(), gw.data_ptr(), None, n, h, wd, cin, w.shape[0], stride,
                                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    return gw


torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
CASES = [("resnet layer2 first", 128, 128, 200, 334, 2), ("resnet layer2", 128, 128, 100, 167, 1), ("resnet layer3 first", 256, 256, 100, 167, 2),
         ("resnet layer3", 256, 256, 50, 84, 1), ("resnet layer4 first", 512, 512, 50, 84, 2), ("resnet layer4", 512, 512, 25, 42, 1),
         ("extra level", 2048, 256, 25, 42, 2)]
for lv, (h, w) in enumerate([(100, 167), (50, 84), (25, 42), (13, 21)]):
    CASES += [(f"D_img conv1 L{lv}", 256, 256, h, w, 1), (f"D_img conv2 L{lv}", 256, 128, h, w, 1), (f"D_img conv3 L{lv}", 128, 128, h, w, 1)]
for name, cin, cout, H, W, s in CASES:
    x = torch.randn(4, cin, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)).contiguous(memory_format=torch.channels_last)
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    gz = torch.randn(4, cout, Ho, Wo, device="cuda").contiguous(memory_format=torch.channels_last)
    bw = lambda need: torch.ops.aten.convolution_backward(gz, x, w, None, [s, s], [1, 1], [1, 1], False, [0, 0], 1, need)
    t_lib_d = timeit(lambda: bw([True, False, False]))
    t_lib_w = timeit(lambda: bw([False, True, False]))
    t_own_w = timeit(lambda: own_wgrad(gz, x, w, s)) if cin % 128 == 0 else float("nan")
    if s == 1 and cout % 32 == 0:
        w_rot = w.flip(2, 3).transpose(0, 1).contiguous(memory_format=torch.channels_last)
        t_own_d = timeit(lambda: _launch(gz, w.flip(2, 3).transpose(0, 1).contiguous(memory_format=torch.channels_last), None, 1, 0))
    else:
        t_own_d = float("nan")
    print(f"{name:20s} {cin:4d}->{cout:4d} {H:3d}x{W:3d} s={s}: dgrad ours {t_own_d*1e3:7.1f} us  cuDNN {t_lib_d*1e3:7.1f} us | "
          f"wgrad ours {t_own_w*1e3:7.1f} us  cuDNN {t_lib_w*1e3:7.1f} us", flush=True)

print("forward: ours (bias + activation fused) vs cuDNN conv + bias, then the activation as a separate pass")
import torch.nn.functional as F
from datr_b200.conv import conv3x3_bias_act
for name, cin, cout, H, W, s in CASES:
    x = torch.randn(4, cin, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)).contiguous(memory_format=torch.channels_last)
    b = torch.randn(cout, device="cuda")
    act = 2 if name.startswith("D_img") else (0 if name.startswith("extra") else 1)
    with torch.no_grad():
        t_own = timeit(lambda: conv3x3_bias_act(x, w, b, s, act))
        if act == 2:
            t_lib = timeit(lambda: F.leaky_relu_(F.conv2d(x, w, b, stride=s, padding=1), 0.2))
        elif act == 1:
            t_lib = timeit(lambda: F.relu_(F.conv2d(x, w, b, stride=s, padding=1)))
        else:
            t_lib = timeit(lambda: F.conv2d(x, w, b, stride=s, padding=1))
    print(f"{name:20s} {cin:4d}->{cout:4d} {H:3d}x{W:3d} s={s} act={act}: forward ours {t_own*1e3:7.1f} us  cuDNN {t_lib*1e3:7.1f} us", flush=True)
