"""Decoder self-attention at the DINO shapes (N=2, 8 heads, 32 channels/head, T = 1100 with the de-noising mask / 900
without): forward + backward time of the SDPA backends, CUDA events, GPU box only."""
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"


def run(T, masked, backend, iters=30):
    g = torch.Generator(device="cpu").manual_seed(1)
    qkv = [torch.randn(2, T, 8, 32, generator=g).to(dev).transpose(1, 2).requires_grad_(True) for _ in range(3)]
    mask = None
    if masked:
        mask = (torch.rand(T, T, generator=g) > 0.2).to(dev)
    go = torch.randn(2, 8, T, 32, device=dev)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for i in range(iters + 5):
        with sdpa_kernel(backend):
            e[0].record()
            o = F.scaled_dot_product_attention(*qkv, attn_mask=mask)
            e[1].record()
            o.backward(go)
            e[2].record()
        torch.cuda.synchronize()
        if i >= 5:
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    return tf / iters * 1e3, tb / iters * 1e3


for T, masked in ((1100, True), (900, False)):
    for name, be in (("efficient", SDPBackend.EFFICIENT_ATTENTION), ("math", SDPBackend.MATH)):
        try:
            f, b = run(T, masked, be)
            print(f"T={T} masked={masked} {name:10s} fwd {f:7.1f} us  bwd {b:7.1f} us")
        except Exception as ex:  # noqa: BLE001
            print(f"T={T} masked={masked} {name}: {type(ex).__name__}: {str(ex)[:100]}")
