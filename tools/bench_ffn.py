"""FFN block (linear1 + ReLU + linear2 + residual, d = 256, d_ffn = 2048) forward and forward + backward at the encoder
(M = 4 x 22223 tokens) and decoder (M = 4 x 1100) sizes: TF32 operands vs bf16 operands, per-GEMM times included.
GPU box only; L2 flushed between calls."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200 import linear as dl

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


dl.set_mode("tf32")
for M in (88892, 4400):
    d, dff = 256, 2048
    x = torch.randn(M, d, device="cuda", requires_grad=True)
    w1 = (torch.randn(dff, d, device="cuda") / 16).requires_grad_(True); b1 = torch.zeros(dff, device="cuda", requires_grad=True)
    w2 = (torch.randn(d, dff, device="cuda") / 45).requires_grad_(True); b2 = torch.zeros(d, device="cuda", requires_grad=True)
    gy = torch.randn(M, d, device="cuda")
    for kind in ("tf32", "bf16"):
        dl.set_ffn_precision(kind)
        with torch.no_grad():
            pass
        t_f = timeit(lambda: dl.ffn(x, w1, b1, w2, b2))
        t_fb = timeit(lambda: dl.ffn(x, w1, b1, w2, b2).backward(gy))
        print(f"M={M} {kind}: forward {t_f:8.1f} us, forward + backward {t_fb:8.1f} us", flush=True)
    xb, w1b, w2b = x.detach().bfloat16(), w1.detach().bfloat16(), w2.detach().bfloat16()
    h = dl._launch_bf16(xb, w1b, b1.detach(), None, 1, True)
    gb = gy.bfloat16()
    xd = x.detach()
    rows = [("linear1 bf16 -> bf16 (+bias, ReLU)", lambda: dl._launch_bf16(xb, w1b, b1.detach(), None, 1, True)),
            ("linear2 bf16 -> fp32 (+bias, +residual)", lambda: dl._launch_bf16(h, w2b, b2.detach(), xd, 0, False)),
            ("dgrad2 bf16 -> bf16 (ReLU mask)", lambda: dl._launch_bf16(gb, w2b.t().contiguous(), None, h, 3, True, residual_bf16=True)),
            ("dgrad1 bf16 -> fp32 (+residual)", lambda: dl._launch_bf16(h, w1b.t().contiguous(), None, gy, 0, False)),
            ("wgrad2 bf16", lambda: dl._wgrad_bf16(gb, h, True)), ("wgrad1 bf16", lambda: dl._wgrad_bf16(h, xb, True)),
            ("cast fp32 -> bf16 [M, 256]", lambda: xd.to(torch.bfloat16)),
            ("linear1 tf32", lambda: dl._launch(xd, w1.detach(), b1.detach(), None, 1)),
            ("linear2 tf32", lambda: dl._launch(h.float(), w2.detach(), b2.detach(), xd, 0)),
            ("wgrad2 tf32", lambda: dl._wgrad(gy, h.float(), True)), ("wgrad1 tf32", lambda: dl._wgrad(h.float(), xd, True))]
    for name, fn in rows:
        print(f"   M={M} {name:42s} {timeit(fn):8.1f} us", flush=True)
