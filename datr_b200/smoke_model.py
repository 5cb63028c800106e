"""Second half of __graft_entry__.smoke(): one tiny DINO training step on cuda:0 (forward, losses, backward) with
the CUDA MSDeformAttn kernels, checked for finiteness and for launch counts; the same step in the benchmarked
tensor-core mode (+ clip / AdamW); and the step replayed as CUDA-graph segments the way bench.py runs it."""
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

    # the same step in the benchmarked mode: tcgen05 linear / weight-gradient kernels, fused self-attention (forward and
    # backward), GPU matcher, then clip + AdamW in one launch; the loss must agree with the fp32 step inside the TF32 class
    from datr_b200 import linear as dl
    from datr_b200.optim import FlatAdamW
    from datr_b200.parallel import FlatGradients, param_groups
    for p in model.parameters():
        p.grad = None
    grads = FlatGradients(model)
    opt = FlatAdamW(param_groups(model, 1e-4, 1e-5), grads, weight_decay=1e-4)
    counts = lambda: (native.linear_launch_count(), native.wgrad_launch_count(), native.attn_launch_count(), native.all_launch_count())
    c0 = counts()
    dl.set_mode("tf32")
    try:
        torch.manual_seed(0)
        out = model(imgs, targets)
        losses = criterion(out, targets)
        loss_tc = sum(losses[k] * criterion.weight_dict[k] for k in losses if k in criterion.weight_dict)
        loss_tc.backward()
        opt.clip_and_step(0.1)
        # bf16 FFN block (the model above is too small to reach its row threshold)
        x = torch.randn(8192, 256, device="cuda", requires_grad=True)
        w1 = (torch.randn(512, 256, device="cuda") / 16).requires_grad_(True); b1 = torch.zeros(512, device="cuda", requires_grad=True)
        w2 = (torch.randn(256, 512, device="cuda") / 22).requires_grad_(True); b2 = torch.zeros(256, device="cuda", requires_grad=True)
        y = dl.ffn(x, w1, b1, w2, b2)
        y.sum().backward()
        ref = torch.relu(x.detach() @ w1.detach().t()) @ w2.detach().t() + x.detach()
        err = float((y.detach() - ref).abs().max() / ref.abs().max())
    finally:
        dl.set_mode("fp32")
    torch.cuda.synchronize()
    c1 = counts()
    assert torch.isfinite(loss_tc).item() and all(torch.isfinite(p).all().item() for p in model.parameters())
    assert c1[0] > c0[0] and c1[1] > c0[1] and c1[2] >= c0[2] + 3 * 2, (c0, c1)      # linear, wgrad, attention fwd + 2 x bwd per layer
    assert err < 2e-2, err
    print(f"[smoke] tensor-core mode ok: loss={loss_tc.item():.4f}, hand-written kernel launches={c1[3] - c0[3]} "
          f"(linear {c1[0] - c0[0]}, weight gradient {c1[1] - c0[1]}, attention {c1[2] - c0[2]}), bf16 FFN rel err {err:.1e}")

    # and as the benchmark replays it (datr_b200.graphs): every static segment as a CUDA graph, parameter gradients put
    # into the flat buffer inside the captured backward, small weight gradients on the graph's parallel branch, the image
    # discriminator on a second stream.  First step captures, second replays.
    from datr_b200 import graphs
    # the eager steps' autograd graphs must be gone before the first capture: their AccumulateGrad nodes belong to the default
    # stream, and a capture must not make that stream wait (torch's own make_graphed_callables has the same precondition)
    del out, losses, loss, loss_tc, y, ref, x, w1, b1, w2, b2, g
    import gc
    gc.collect()
    fast = model.to(memory_format=torch.channels_last)
    sg = graphs.StepGraphs()
    graphs.ACTIVE = sg
    dl.set_mode("tf32")
    try:
        flat = FlatGradients(fast)
        timgs = torch.zeros(4, 3, 160, 200, device="cuda")
        from datr_b200.util.misc import NestedTensor
        mask = torch.ones(4, 160, 200, dtype=torch.bool, device="cuda")
        for i, im in enumerate(imgs):
            timgs[i, :, :im.shape[1], :im.shape[2]] = im
            mask[i, :im.shape[1], :im.shape[2]] = False
        samples = NestedTensor(timgs.contiguous(memory_format=torch.channels_last), mask)
        for it in range(2):
            sg.begin_step()
            flat.zero()
            torch.manual_seed(0)
            out = fast(samples, targets)
            losses = criterion(out, targets)
            loss_g = losses["_weighted_total"] if "_weighted_total" in losses else \
                sum(losses[k] * criterion.weight_dict[k] for k in losses if k in criterion.weight_dict)
            loss_g.backward()
        torch.cuda.synchronize()
    finally:
        graphs.ACTIVE = None
        dl.set_mode("fp32")
    tgs = [g for g, _ in sg.cache.values() if isinstance(g, graphs._TrainingGraph)]
    assert torch.isfinite(loss_g).item() and torch.isfinite(flat.flat).all().item()
    assert float(flat.flat.abs().max()) > 0 and flat.check_views()
    assert sg.captures >= 5 and sum(g.n_sunk for g in tgs) > 0, (sg.captures, sum(g.n_sunk for g in tgs))
    print(f"[smoke] graph replay ok: loss={float(loss_g.detach()):.4f}, {sg.captures} segments, "
          f"{sum(g.n_sunk for g in tgs)} parameter gradients accumulated inside the captured backward, "
          f"{sum(g.side_launches for g in tgs)} launches on the parallel branch")

