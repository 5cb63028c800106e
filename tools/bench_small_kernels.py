"""Micro-benchmark of the LayerNorm-backward and column-sum kernels vs ATen at the DINO-4scale shapes (GPU box only)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200.linear import _colsum
from datr_b200.layernorm import layer_norm
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


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


M = 44446
for cols in (256, 128, 2048):
    dy = torch.randn(M, cols, device="cuda"); y = torch.randn(M, cols, device="cuda")
    t1 = timeit(lambda: _colsum(dy)); t2 = timeit(lambda: dy.sum(0))
    t3 = timeit(lambda: _colsum(dy, y)); t4 = timeit(lambda: torch.ops.aten.threshold_backward(dy, y, 0.0).sum(0))
    b1, b3 = 4.0 * M * cols, 12.0 * M * cols
    print(f"colsum cols={cols:5d}: ours {t1*1e3:7.1f} us ({b1/t1/1e6:7.1f} GB/s = {b1/t1/1e6/HBM:5.3f}) aten {t2*1e3:7.1f} us | relu_bwd+colsum ours {t3*1e3:7.1f} us ({b3/t3/1e6:7.1f} GB/s = {b3/t3/1e6/HBM:5.3f}) aten {t4*1e3:7.1f} us", flush=True)
x = torch.randn(2, 22223, 256, device="cuda", requires_grad=True); g = torch.randn(2, 22223, 256, device="cuda")
norm = torch.nn.LayerNorm(256).cuda()
yo = layer_norm(norm, x); ya = norm(x)
t1 = timeit(lambda: torch.autograd.grad(yo, (x, norm.weight, norm.bias), g, retain_graph=True))
t2 = timeit(lambda: torch.autograd.grad(ya, (x, norm.weight, norm.bias), g, retain_graph=True))
b = 12.0 * M * 256
print(f"layernorm256 bwd: ours {t1*1e3:7.1f} us ({b/t1/1e6:7.1f} GB/s = {b/t1/1e6/HBM:5.3f}) aten {t2*1e3:7.1f} us")
