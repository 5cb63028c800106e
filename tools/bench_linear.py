"""Micro-benchmark of the tcgen05 linear kernel vs cuBLAS (TF32 and fp32) at the DINO-4scale shapes (GPU box only)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from datr_b200 import linear as dl

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


def main():
    M = 44446
    for name, N, K, relu, res in [("value_proj/out_proj", 256, 256, False, False), ("offsets", 256, 256, False, False),
                                  ("attn_weights", 128, 256, False, False), ("linear1+relu", 2048, 256, True, False),
                                  ("linear2+res", 256, 2048, False, True)]:
        x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda") if res else None
        flops = 2.0 * M * N * K
        bytes_ = 4.0 * (M * K + N * K + M * N * (2 if res else 1))
        dl.set_mode("tf32")
        t_ours = timeit(lambda: dl.linear(x, w, b, relu=relu, residual=r))
        dl.set_mode("fp32")

        def lib():
            y = torch.nn.functional.linear(x, w, b)
            if relu:
                y = torch.relu_(y)
            if r is not None:
                y = y + r
            return y
        torch.backends.cuda.matmul.allow_tf32 = True
        t_tf32 = timeit(lib)
        torch.backends.cuda.matmul.allow_tf32 = False
        t_fp32 = timeit(lib)
        print(f"{name:20s} M={M} N={N:5d} K={K:5d}  ours {t_ours*1e3:8.1f} us ({flops/t_ours/1e9:7.1f} TF/s, {bytes_/t_ours/1e6:7.1f} GB/s = {bytes_/t_ours/1e6/HBM:5.3f} of HBM)"
              f"   cuBLAS-tf32 {t_tf32*1e3:8.1f} us   cuBLAS-fp32 {t_fp32*1e3:8.1f} us", flush=True)


def wgrad():
    """Weight + bias gradient kernel vs cuBLAS TF32 (dz^T @ x) + ATen column sum."""
    from datr_b200.linear import _wgrad
    M = 44446
    for name, N, K in [("value/out/offsets proj", 256, 256), ("attn_weights", 128, 256), ("linear1", 2048, 256), ("linear2", 256, 2048)]:
        dz = torch.randn(M, N, device="cuda"); x = torch.randn(M, K, device="cuda")
        flops = 2.0 * M * N * K
        bytes_ = 4.0 * (M * N + M * K + N * K)
        t_ours = timeit(lambda: _wgrad(dz, x, True))
        torch.backends.cuda.matmul.allow_tf32 = True
        t_lib = timeit(lambda: (dz.t() @ x, dz.sum(0)))
        torch.backends.cuda.matmul.allow_tf32 = False
        print(f"wgrad {name:24s} M={M} N={N:5d} K={K:5d}  ours {t_ours*1e3:8.1f} us ({flops/t_ours/1e9:7.1f} TF/s, {bytes_/t_ours/1e6:7.1f} GB/s = {bytes_/t_ours/1e6/HBM:5.3f} of HBM)"
              f"   cuBLAS-tf32 + sum {t_lib*1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
    wgrad()
