"""Decoder self-attention at the config-2 shapes (4 images x 8 heads, T = 1100, d = 32, de-noising mask): the one-kernel
tensor-core attention (csrc/attn_fused.cu) against the round-1 path (two cuBLAS batched GEMMs around the softmax kernel)
and torch SDPA.  GPU box only; prints microseconds per call (L2 flushed between calls)."""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import attention
from test_attention_gpu import dn_mask

torch.backends.cuda.matmul.allow_tf32 = True
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


for N, H, T in ((4, 8, 1100), (2, 8, 1100), (4, 8, 900)):
    C = 32 * H
    qk = torch.randn(N, T, 2 * C, device="cuda"); v = torch.randn(N, T, C, device="cuda")
    blocked = dn_mask(T, 200, 10, None).cuda()
    bits = attention.pack_mask(blocked, T, qk.device)
    attention._BACKWARD = "fused"
    go = torch.randn(N, T, C, device="cuda")
    q4, k4 = (t.reshape(N, T, H, 32).transpose(1, 2) for t in (qk[..., :C], qk[..., C:]))
    v4 = v.reshape(N, T, H, 32).transpose(1, 2)
    t_pack = timeit(lambda: attention.pack_mask(blocked, T, qk.device))
    with torch.no_grad():
        t_fused = timeit(lambda: attention.fused_self_attention(qk, v, H, blocked, bits=bits))
        t_old = timeit(lambda: attention.self_attention(q4, k4, v4, blocked))
        t_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4, attn_mask=~blocked))
    qk_g, v_g = qk.clone().requires_grad_(True), v.clone().requires_grad_(True)
    attention._BACKWARD = "gemm"
    t_fused_p = timeit(lambda: attention.fused_self_attention(qk_g, v_g, H, blocked, bits=bits))

    def fb_new():
        attention.fused_self_attention(qk_g, v_g, H, blocked, bits=bits).backward(go)

    def fb_old():
        q_, k_ = (t.reshape(N, T, H, 32).transpose(1, 2) for t in (qk_g[..., :C], qk_g[..., C:]))
        attention.self_attention(q_, k_, v_g.reshape(N, T, H, 32).transpose(1, 2), blocked).transpose(1, 2).reshape(N, T, C).backward(go)
    t_fb_new, t_fb_old = timeit(fb_new), timeit(fb_old)
    attention._BACKWARD = "fused"
    t_fb_fused = timeit(fb_new)
    flops = 4.0 * N * H * T * T * 32
    print(f"N={N} H={H} T={T}: fused fwd {t_fused:7.1f} us ({flops / t_fused / 1e6:6.1f} TFLOP/s) | fused fwd + probabilities out "
          f"{t_fused_p:7.1f} | bmm+softmax+bmm fwd {t_old:7.1f} | SDPA fwd {t_sdpa:7.1f} | mask pack {t_pack:5.1f} | "
          f"fwd+bwd: fused fwd + fused bwd {t_fb_fused:7.1f}, fused fwd + GEMM bwd {t_fb_new:7.1f}, round-1 path {t_fb_old:7.1f}", flush=True)
