"""Short target for ncu: the module-level fused MSDeformAttn kernels (encoder call: 2-d reference points, merged
offsets+logits GEMM output as strided input; decoder call: 4-d reference boxes), the LayerNorm(256) forward and the
padding-mask kernel at the BASELINE.json configs[1] shapes (N=2, S=22223).  Usage under gpurun:
  ncu --set full --clock-control none --import-source on -k "regex:msda_|layernorm256_fwd|zero_masked" -s 6 -c 6 \
      -o gpurun_out/prof_fused python tools/ncu_target_fused.py
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench
from datr_b200 import MultiScaleDeformableAttention as MSDA
from datr_b200 import native

dev = torch.device("cuda", 0)
S, _ = bench.step_plan()
L, M, P = len(bench.CFG2_LEVELS), 8, 4
T = M * L * P
shapes = torch.tensor(bench.CFG2_LEVELS, dtype=torch.int64, device=dev)
hw = shapes[:, 0] * shapes[:, 1]
lstart = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
g = torch.Generator(device="cpu").manual_seed(5)


def call(Lq, ref_dim):
    value = torch.randn((2, S, M, 32), generator=g).to(dev)
    merged = torch.cat((torch.randn((2, Lq, 2 * T), generator=g) * 2.0, torch.randn((2, Lq, T), generator=g)), -1).to(dev)
    if ref_dim == 2:
        refs = []
        for h, w in bench.CFG2_LEVELS:
            ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
            refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
        ref = torch.cat(refs)[None, :, None, :].expand(2, Lq, L, 2).contiguous().to(dev)
    else:
        ref = torch.cat((torch.rand((2, Lq, 1, 2), generator=g), torch.rand((2, Lq, 1, 2), generator=g) * 0.3 + 0.02), -1) \
            .expand(2, Lq, L, 4).contiguous().to(dev)
    gout = torch.randn((2, Lq, M * 32), generator=g).to(dev)
    return value, merged, ref, gout


sets = [call(S, 2), call(1100, 4)]
x = torch.randn(2 * S, 256, device=dev); w = torch.ones(256, device=dev); b = torch.zeros(256, device=dev)
y = torch.empty_like(x); st = torch.empty(2, 2 * S, device=dev)
mask = torch.zeros(2 * S, dtype=torch.bool, device=dev); mask[-3000:] = True
lib = native.lib()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for value, merged, ref, gout in sets:
        N, Lq = merged.shape[:2]
        off = merged[..., :2 * T].view(N, Lq, M, L, P, 2); lg = merged[..., 2 * T:].view(N, Lq, M, L * P)
        MSDA.ms_deform_attn_fused_forward(value, shapes, lstart, off, lg, ref)
        MSDA.ms_deform_attn_fused_backward(value, shapes, lstart, off, lg, ref, gout, merged_grad=torch.empty_like(merged))
    lib.datr_layernorm256_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), 1e-5, y.data_ptr(), st[0].data_ptr(),
                                  st[1].data_ptr(), 2 * S, torch.cuda.current_stream().cuda_stream)
    lib.datr_zero_masked_rows(y.data_ptr(), mask.data_ptr(), 2 * S, 256, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done")
