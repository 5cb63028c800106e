"""CPU, world_size 2 over gloo: the drop-in model under the reference's own data-parallel wrapper,
torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True) (main.py:156): one DA training step per
rank on different targets; the all-reduced gradients equal the average of the two single-process gradients.
(The product's native alternative, datr_b200.parallel.FlatGradients, is covered by tests/test_parallel_cpu.py.)
The MSDeformAttn op is served by the oracle's grid_sample port inside the workers (host-logic test)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import model_cases as mcase

TARGET_SEEDS = {0: 5, 1: 6}


def _build():
    from oracle import msda as om
    from datr_b200.models.dino.ops.modules import ms_deform_attn as mod

    class OracleFn:
        @staticmethod
        def apply(value, shapes, level_start, loc, attn, step):
            return om.core_torch(value, shapes, loc, attn)
    mod.MSDeformAttnFunction = OracleFn
    from datr_b200.models.dino.dino import build_dino
    torch.manual_seed(0)
    model, crit, _ = build_dino(mcase.small_args(enc_layers=1, dec_layers=1))
    model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
    model.train(); crit.train()
    return model, crit


def _step(model, crit, rank, num_boxes_override=None):
    inner = model.module if hasattr(model, "module") else model
    inner.global_proto = None
    torch.manual_seed(7)
    tg = mcase.targets(seed=TARGET_SEEDS[rank])
    out = model(mcase.images(), tg)
    losses = crit(out, tg)
    total = mcase.total_loss(losses, crit.weight_dict)
    model.zero_grad()
    total.backward()
    return {k: p.grad.reshape(-1)[::53].clone() for k, p in inner.named_parameters() if p.grad is not None}


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, crit = _build()
        ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
        out[rank] = {k: v.numpy() for k, v in _step(ddp, crit, rank).items()}
    finally:
        dist.destroy_process_group()


def test_model_trains_under_distributed_data_parallel_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0].keys() == out[1].keys()
    for k in out[0]:
        assert np.array_equal(out[0][k], out[1][k]), k                   # DDP leaves identical gradients on every rank
    # single-process reference: same two steps with the criterion's num_boxes averaged over the "ranks" by hand
    model, crit = _build()
    counts = {r: sum(len(t["labels"]) for t in mcase.targets(seed=TARGET_SEEDS[r])) for r in (0, 1)}
    mean_boxes = (counts[0] + counts[1]) / 2
    grads = []
    for r in (0, 1):
        import datr_b200.models.dino.matcher as mt
        orig = mt.BatchedMatch.result
        mt.BatchedMatch.result = lambda self, _o=orig: (_o(self)[0], max(mean_boxes, 1.0))     # what the all-reduce yields
        try:
            grads.append(_step(model, crit, r))
        finally:
            mt.BatchedMatch.result = orig
    for k in out[0]:
        want = ((grads[0][k] + grads[1][k]) / 2).numpy()
        scale = max(float(np.abs(want).max()), 1e-6)
        assert float(np.abs(out[0][k] - want).max()) / scale < 2e-3, k
