"""Second half of __graft_entry__.smoke(): one tiny DINO training step on cuda:0 (forward, losses, backward) with
the CUDA MSDeformAttn kernels, checked for finiteness and for launch counts."""
import torch


def run():
    from datr_b200 import native
    from datr_b200.config import dino_args
    from datr_b200.models.dino.dino import build_dino
    torch.manual_seed(0)
    args = dino_args(device="cuda", enc_layers=2, dec_layers=2, dim_feedforward=128, num_queries=50, num_classes=9,
                     dn_labelbook_size=9, num_select=50)
    model, criterion, _ = build_dino(args)
    model.cuda().train()
    criterion.train()
    imgs = [torch.randn(3, 160, 200, device="cuda"), torch.randn(3, 144, 176, device="cuda"),
            torch.randn(3, 160, 192, device="cuda"), torch.randn(3, 128, 200, device="cuda")]
    targets = [{"labels": torch.tensor([1, 3], device="cuda"), "boxes": torch.tensor([[.5, .5, .2, .3], [.3, .6, .1, .2]], device="cuda")},
               {"labels": torch.tensor([2], device="cuda"), "boxes": torch.tensor([[.4, .4, .3, .3]], device="cuda")}]
    n0 = native.launch_count()
    out = model(imgs, targets)
    losses = criterion(out, targets)
    loss = sum(losses[k] * criterion.weight_dict[k] for k in losses if k in criterion.weight_dict)
    loss.backward()
    torch.cuda.synchronize()
    launches = native.launch_count() - n0
    assert torch.isfinite(loss).item(), "non-finite loss"
    # source + target halves share ONE encoder and ONE decoder pass: (2 enc + 2 dec layers) x (fwd + bwd)
    assert launches == 2 * (2 + 2), launches
    g = model.transformer.encoder.layers[0].self_attn.sampling_offsets.weight.grad
    assert g is not None and torch.isfinite(g).all().item()
    print(f"[smoke] DINO DA training step ok: loss={loss.item():.4f}, MSDeformAttn launches={launches}")
