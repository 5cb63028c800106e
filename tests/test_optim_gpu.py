"""GPU parity of the one-launch clip + AdamW step (datr_b200.optim.FlatAdamW, include/datr_adamw.h) against
torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW, the tail of the reference's iteration (engine.py:108-111, main.py:152)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 5, 1),
                            torch.nn.Flatten(), torch.nn.Linear(5 * 36, 33), torch.nn.Linear(33, 7)).cuda()
    return m.to(memory_format=torch.channels_last)


@pytest.mark.parametrize("max_norm", [0.1, None])
def test_flat_adamw_matches_torch_adamw(max_norm):
    from datr_b200 import native
    from datr_b200.optim import FlatAdamW
    from datr_b200.parallel import FlatGradients
    ours, ref = _model(1), _model(1)
    groups = lambda m: [{"params": [p for n, p in m.named_parameters() if not n.startswith("0.")], "lr": 1e-3},
                        {"params": [p for n, p in m.named_parameters() if n.startswith("0.")], "lr": 1e-4}]
    grads = FlatGradients(ours)
    opt = FlatAdamW(groups(ours), grads, weight_decay=1e-2)
    topt = torch.optim.AdamW(groups(ref), lr=1e-3, weight_decay=1e-2, fused=True)
    g = torch.Generator(device="cpu").manual_seed(2)
    n0 = native.all_launch_count()
    for step in range(30):
        x = torch.randn(4, 3, 6, 6, generator=g).cuda().contiguous(memory_format=torch.channels_last)
        grads.zero()
        (ours(x) ** 2).sum().backward()
        for a, b in zip(ours.parameters(), ref.parameters()):       # both optimizers see the same gradients
            b.grad = a.grad.detach().clone()
        norm = opt.clip_and_step(max_norm)
        if max_norm is not None:
            tnorm = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
            assert abs(float(norm) - float(tnorm)) <= 1e-5 * float(tnorm)
        topt.step()
        if step == 15:           # learning-rate drop (lr_scheduler of main.py): the table follows the group values
            for grp in opt.param_groups + topt.param_groups:
                grp["lr"] *= 0.1
    assert native.all_launch_count() == n0 + 30
    for (n, a), b in zip(ours.named_parameters(), ref.parameters()):
        assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(b.abs().max())), n
    st = topt.state_dict()["state"]
    # moments: the flat buffers hold what torch keeps per parameter
    order = [p for grp in groups(ours) for p in grp["params"]]
    for i, p in enumerate(order):
        off = (grads.views[[id(q) for q in grads.params].index(id(p))].data_ptr() - grads.flat.data_ptr()) // 4
        m = opt.exp_avg[off:off + p.numel()]
        want = st[i]["exp_avg"]
        assert float((m - want.as_strided((p.numel(),), (1,)).reshape(-1)).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))
