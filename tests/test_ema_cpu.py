"""CPU: the EMA mirror (datr_b200/models/dino/EMA.py) against the reference's own models/dino/EMA.py imported from
/root/reference (build container only) and against the update rule written out, bit-exact."""
import importlib.util
import os

import pytest
import torch

from datr_b200.models.dino import EMA

REF = "/root/reference/models/dino/EMA.py"


def net(seed):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    m(torch.randn(4, 7))                      # moves the BatchNorm buffers (float) and num_batches_tracked (int)
    return m


def ref_module():
    spec = importlib.util.spec_from_file_location("_ref_ema", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.exists(REF), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("cls,kw", [("ModelEMA", dict(decay=0.999)), ("CosineEMA", dict(decay_start=0.9, decay_end=0.99, total_epoch=10))])
def test_matches_reference_classes_bit_exact(cls, kw):
    R = ref_module()
    student = net(1)
    ours, theirs = getattr(EMA, cls)(net(0), **kw), getattr(R, cls)(net(0), **kw)
    for step in range(5):
        with torch.no_grad():
            for p in student.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        if cls == "CosineEMA":
            ours.update_decay(step + 1); theirs.update_decay(step + 1)
            assert ours.decay == theirs.decay
        ours.update(student); theirs.update(student)
    for (k, a), (_, b) in zip(ours.ema.state_dict().items(), theirs.ema.state_dict().items()):
        assert torch.equal(a, b), k
    assert ours.updates == theirs.updates and not any(p.requires_grad for p in ours.ema.parameters())
    assert not ours.ema.training


def test_update_rule_and_api():
    student, teacher = net(1), net(0)
    before = {k: v.clone() for k, v in teacher.state_dict().items()}
    e = EMA.SemiSupModelEMA(teacher, decay=0.9)
    e.update(student)
    for k, v in e.ema.state_dict().items():
        if v.dtype.is_floating_point:
            want = before[k] * 0.9
            want += (1.0 - 0.9) * student.state_dict()[k]
            assert torch.equal(v, want), k
        else:
            assert torch.equal(v, before[k]), k          # integer buffers are left alone (EMA.py:48)
    student.some_flag = 3
    e.update_attr(student, include=("some_flag",))
    assert e.ema.some_flag == 3
    assert EMA.is_parallel(student) is False
    m = EMA.ModelEMA(teacher, decay=0.9999, updates=10)
    assert abs(m.decay(2000) - 0.9999 * (1 - 2.718281828459045 ** -1)) < 1e-12


def test_alias_grouping_routes_partial_overlaps_to_the_sequential_path():
    """Host logic of StateDictEMA on CPU tensors (everything takes the reference's sequential ops there) and the
    grouping of aliased names: identical storages fold into one entry with a multiplicity."""
    from datr_b200.ema import StateDictEMA
    a, b = torch.randn(10), torch.randn(10)
    want = a.clone()
    pair = StateDictEMA([a, a, a[2:6]], [b, b, b[2:6]])
    assert not pair.fast and len(pair.slow) == 3
    pair.update(0.5)
    for e, m in ((want, b), (want, b), (want[2:6], b[2:6])):
        e *= 0.5
        e += 0.5 * m
    assert torch.equal(a, want)
