"""CPU, world_size 2 over gloo: the data-parallel gradient path of datr_b200/parallel.py (one flat buffer, one
all-reduce, DDP's averaging semantics, unused parameters contribute zeros), in both buffer-filling modes, against
the single-process average of the two ranks' gradients.  Mirrors what main.py:156 gets from
DistributedDataParallel(find_unused_parameters=True)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from datr_b200.parallel import FlatGradients, broadcast_parameters


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(8, 16)
        self.b = torch.nn.Linear(16, 4)
        self.unused = torch.nn.Linear(3, 3)          # never on the path: must come out as zeros, not None
        self.conv = torch.nn.Conv2d(2, 4, 3).to(memory_format=torch.channels_last)

    def forward(self, x, img):
        h = self.b(torch.relu(self.a(x))) + self.b(torch.tanh(self.a(x * 0.5)))     # shared weights: two arrivals
        return h.sum() + self.conv(img).mean()


def batch(rank):
    g = torch.Generator().manual_seed(42 + rank)                                    # main.py:138: seed + rank
    return torch.randn(5, 8, generator=g), torch.randn(2, 2, 6, 6, generator=g)


def local_grads(rank):
    torch.manual_seed(0)
    net = Net()
    net(*batch(rank)).backward()
    return [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in net.parameters()]


def worker(rank, world, port, gather, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                 # different initial weights per rank: the broadcast must fix that
        net = Net()
        broadcast_parameters(net)
        grads = FlatGradients(net, gather=gather)
        for step in range(2):                   # second step: the buffer is reused
            grads.zero()
            net(*batch(rank)).backward()
            grads.all_reduce()
            norm = grads.clip_(0.1)
        assert grads.check_views()
        out[rank] = ([p.grad.clone() for p in net.parameters()], float(norm), [p.detach().clone() for p in net.parameters()])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gather", [True, False])
def test_flat_gradient_allreduce_world2(gather):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(worker, args=(2, port, gather, out), nprocs=2, join=True)
    torch.manual_seed(0)
    ref_params = [p.detach() for p in Net().parameters()]
    want = [(a + b) / 2 for a, b in zip(local_grads(0), local_grads(1))]
    total = torch.sqrt(sum((g ** 2).sum() for g in want))
    scale = min(1.0, 0.1 / (float(total) + 1e-6))
    for rank in (0, 1):
        got, norm, params = out[rank]
        for p, q in zip(params, ref_params):
            assert torch.equal(p, q), "rank 0's initial weights must reach every rank"
        assert abs(norm - float(total)) < 1e-5 * max(1.0, float(total))
        for g, w in zip(got, want):
            assert torch.allclose(g, w * scale, rtol=1e-5, atol=1e-7)
    assert all(float(g.abs().max()) == 0.0 for g in out[0][0][4:6]), "unused parameters contribute zeros"


def test_single_process_modes_agree():
    res = []
    for gather in (True, False):
        torch.manual_seed(0)
        net = Net()
        grads = FlatGradients(net, gather=gather)
        grads.zero()
        net(*batch(0)).backward()
        grads.all_reduce()
        res.append(grads.flat.clone())
    assert torch.allclose(res[0], res[1], rtol=1e-6, atol=1e-8)


class TwoStage(torch.nn.Module):
    """`stem` runs first in the forward, so its gradients arrive last (the role of DINO's backbone)."""

    def __init__(self):
        super().__init__()
        self.head = torch.nn.Linear(16, 4)
        self.stem = torch.nn.Linear(8, 16)
        self.fired = 0

    def forward(self, x, on_features_grad=None):
        f = torch.relu(self.stem(x))
        if on_features_grad is not None:
            f.register_hook(lambda g: (on_features_grad(), None)[1])       # every head gradient is final when this fires
        return self.head(f).sum()


def overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = TwoStage()
        grads = FlatGradients(net, late=lambda n: n.startswith("stem"))
        assert grads.split == sum(p.numel() for p in net.head.parameters())
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(42 + rank))
        for step in range(2):
            grads.zero()
            net(x, on_features_grad=grads.reduce_early).backward()
            assert grads._early_handle is not None                           # the hook started the first collective
            grads.all_reduce()
            norm = grads.clip_(0.1)
        out[rank] = ({n: p.grad.clone() for n, p in net.named_parameters()}, float(norm))
    finally:
        dist.destroy_process_group()


def test_early_reduce_from_a_backward_hook_gives_the_same_averaged_clipped_gradients():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(overlap_worker, args=(2, port, out), nprocs=2, join=True)
    local = []
    for rank in (0, 1):
        torch.manual_seed(0)
        net = TwoStage()
        net(torch.randn(5, 8, generator=torch.Generator().manual_seed(42 + rank))).backward()
        local.append({n: p.grad.clone() for n, p in net.named_parameters()})
    want = {n: (local[0][n] + local[1][n]) / 2 for n in local[0]}
    total = torch.sqrt(sum((g ** 2).sum() for g in want.values()))
    scale = min(1.0, 0.1 / (float(total) + 1e-6))
    for rank in (0, 1):
        got, norm = out[rank]
        assert abs(norm - float(total)) < 1e-5 * max(1.0, float(total))
        for n in want:
            assert torch.allclose(got[n], want[n] * scale, rtol=1e-5, atol=1e-7), n
