"""Weight-gradient kernel at decoder sizes (M = 4 x 1100 rows): time per launch for the minimum slab height given in
DATR_WGRAD_MIN_ROWS (read once at library load).  GPU box only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200 import linear as dl

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


out = []
for M in (4400, 2200):
    for N, K in ((256, 256), (384, 256), (512, 256), (2048, 256), (256, 2048), (256, 512)):
        dz = torch.randn(M, N, device="cuda"); x = torch.randn(M, K, device="cuda")
        out.append(f"M={M} N={N} K={K}: {timeit(lambda: dl._wgrad(dz, x, True)):6.1f} us")
print(f"min rows per slab {os.environ.get('DATR_WGRAD_MIN_ROWS', '128 (default)')}: " + " | ".join(out))
