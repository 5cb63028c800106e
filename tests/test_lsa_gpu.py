"""GPU linear sum assignment (include/datr_lsa.h) against scipy.optimize.linear_sum_assignment, the solver the reference's
matcher calls (models/dino/matcher.py:91): identical index vectors -- the north star asks for bit-exact assignment work --
on random fp32 cost matrices, matrices full of ties (integers, constants, duplicated queries) and the matcher's own
batched layout."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu


def solve(costs):
    """costs: list of [nq, nt] float32 arrays -> list of (i, j) from the GPU solver (one launch for all)."""
    from datr_b200 import native
    lib = native.lib()
    flat, rows, off, eoff = [], [], 0, 0
    for c in costs:
        nq, nt = c.shape
        rows.append([eoff, nt, nq, nt, off])
        flat.append(np.ascontiguousarray(c, dtype=np.float32).reshape(-1))
        eoff += nq * nt
        off += 2 * nt
    cost = torch.from_numpy(np.concatenate(flat) if flat else np.zeros(1, np.float32)).cuda()
    table = torch.tensor(rows, dtype=torch.int64, device="cuda")
    out = torch.full((max(off, 1),), -7, dtype=torch.int64, device="cuda")
    rc = lib.datr_lsa_solve(cost.data_ptr(), table.data_ptr(), len(costs), max(c.shape[0] for c in costs),
                            max(c.shape[1] for c in costs), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.datr_lsa_last_error().decode()
    out = out.cpu().numpy()
    res = []
    for (_, _, _, nt, o) in rows:
        res.append((out[o:o + nt], out[o + nt:o + 2 * nt]))
    return res


@pytest.mark.parametrize("seed", range(4))
def test_random_cost_matrices_match_scipy_exactly(seed):
    rng = np.random.default_rng(seed)
    costs = [rng.standard_normal((nq, nt)).astype(np.float32) * rng.choice([1e-3, 1.0, 50.0])
             for nq, nt in [(900, 1), (900, 7), (900, 20), (900, 63), (1100, 100), (300, 300), (37, 5), (5, 5), (900, 0), (64, 33)]]
    for c, (i, j) in zip(costs, solve(costs)):
        wi, wj = linear_sum_assignment(c)
        assert np.array_equal(i, wi) and np.array_equal(j, wj), c.shape


def test_ties_are_broken_like_scipy():
    rng = np.random.default_rng(5)
    costs = [rng.integers(0, 3, (900, 20)).astype(np.float32),           # small integers: massive ties
             np.zeros((50, 10), np.float32),                             # constant matrix (scipy issue 11602: identity)
             np.ones((12, 12), np.float32),
             rng.integers(0, 2, (200, 40)).astype(np.float32),
             np.repeat(rng.standard_normal((30, 9)).astype(np.float32), 10, axis=0),   # every query duplicated 10 times
             np.tile(rng.standard_normal((100, 1)).astype(np.float32), (1, 8))]        # every box identical
    for c, (i, j) in zip(costs, solve(costs)):
        wi, wj = linear_sum_assignment(c)
        assert np.array_equal(i, wi) and np.array_equal(j, wj), c.shape


def test_batched_match_on_the_device_equals_the_host_solver(monkeypatch):
    """The matcher's batched layout ([sets * images, queries, all boxes], one problem per (set, image) column block)."""
    from datr_b200.models.dino import matcher as mm
    torch.manual_seed(3)
    bs, nq, nc = 2, 900, 91
    sets = [{"pred_logits": torch.randn(bs, nq, nc, device="cuda"), "pred_boxes": torch.rand(bs, nq, 4, device="cuda") * 0.5 + 0.2}
            for _ in range(7)]
    targets = [{"labels": torch.randint(0, nc, (n,), device="cuda"), "boxes": torch.rand(n, 4, device="cuda") * 0.4 + 0.2}
               for n in (13, 4)]
    m = mm.HungarianMatcher(2.0, 5.0, 2.0, 0.25)
    monkeypatch.setattr(mm, "DEVICE_SOLVER", True)
    dev_handle = mm.BatchedMatch(m, sets, targets)
    assert dev_handle.flat_dev is not None
    got, nb = dev_handle.result()
    monkeypatch.setattr(mm, "DEVICE_SOLVER", False)
    want, nb_host = mm.BatchedMatch(m, sets, targets).result()
    assert nb == nb_host == 17.0
    for g, w in zip(got, want):
        for (gi, gj), (wi, wj) in zip(g, w):
            assert torch.equal(gi.cpu(), wi.cpu()) and torch.equal(gj.cpu(), wj.cpu())
