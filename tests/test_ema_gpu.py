"""GPU: the multi-tensor EMA kernel (include/datr_ema.h) against the reference's per-tensor loop
(models/dino/EMA.py:47-50: `v *= d; v += (1 - d) * m`) on the same tensors -- bit-exact -- including ragged sizes,
unaligned views and a full DINO-sized state dict in one launch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_kernel_is_bit_exact_with_the_reference_loop():
    from datr_b200 import native
    from datr_b200.ema import StateDictEMA
    g = torch.Generator(device="cpu").manual_seed(0)
    sizes = [1, 3, 4, 5, 255, 256, 16384, 16385, 40000, 256 * 2048, 7]
    base = torch.randn(sum(sizes) + 1, generator=g).cuda()
    ema, off = [], 1                                         # offset 1: every view is 4-byte but not 16-byte aligned
    for n in sizes:
        ema.append(base[off:off + n]); off += n
    ema += [torch.randn(64, 3, 7, 7, generator=g).cuda(), torch.arange(5).cuda()]          # + an integer buffer
    model = [torch.randn(t.shape, generator=g).cuda() if t.dtype.is_floating_point else t.clone() for t in ema]
    want = [t.clone() for t in ema]
    pair = StateDictEMA(ema, model)
    assert len(pair.fast) == len(sizes) + 1 and not pair.slow
    n0 = native.ema_launch_count()
    for d in (0.9, 0.9997, 0.99990001):
        pair.update(d)
        for v, m in zip(want, model):
            if v.dtype.is_floating_point:
                v *= d
                v += (1.0 - d) * m
    assert native.ema_launch_count() == n0 + 3
    torch.cuda.synchronize()
    for a, b in zip(ema, want):
        assert torch.equal(a, b)


def test_model_ema_classes_on_cuda_match_cpu():
    from datr_b200.models.dino import EMA
    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(300, 200), torch.nn.LayerNorm(200), torch.nn.Linear(200, 10))
    t_cpu, s_cpu = mk(), mk()
    t_gpu, s_gpu = mk().cuda(), mk().cuda()
    t_gpu.load_state_dict(t_cpu.state_dict()); s_gpu.load_state_dict(s_cpu.state_dict())
    a, b = EMA.ModelEMA(t_cpu, decay=0.99), EMA.ModelEMA(t_gpu, decay=0.99)
    for _ in range(3):
        a.update(s_cpu); b.update(s_gpu)
    for (k, x), (_, y) in zip(a.ema.state_dict().items(), b.ema.state_dict().items()):
        assert torch.equal(x, y.cpu()), k


def test_aliased_state_dict_entries_get_one_update_per_name_like_the_reference_loop():
    """A state dict that lists one tensor under several names (DINO's shared box / class heads appear 12 times): the
    reference loop updates it once per name; the kernel path must reproduce exactly that, deterministically."""
    from datr_b200 import native
    from datr_b200.ema import StateDictEMA
    g = torch.Generator(device="cpu").manual_seed(1)
    shared_e, shared_m = torch.randn(256, 256, generator=g).cuda(), torch.randn(256, 256, generator=g).cuda()
    other_e, other_m = torch.randn(40000, generator=g).cuda(), torch.randn(40000, generator=g).cuda()
    ema = [shared_e, other_e] + [shared_e] * 11
    model = [shared_m, other_m] + [shared_m] * 11
    want_shared, want_other = shared_e.clone(), other_e.clone()
    pair = StateDictEMA(ema, model)
    assert sorted(k for _, k in pair.fast) == [1, 12] and not pair.slow
    n0 = native.ema_launch_count()
    for d in (0.5, 0.999):
        pair.update(d)
        for _ in range(12):
            want_shared *= d
            want_shared += (1.0 - d) * shared_m
        want_other *= d
        want_other += (1.0 - d) * other_m
    assert native.ema_launch_count() == n0 + 2 * 12
    assert torch.equal(shared_e, want_shared) and torch.equal(other_e, want_other)


def test_real_dino_with_shared_heads_matches_the_reference_loop():
    import model_cases as mcase
    from datr_b200.models.dino import EMA
    from datr_b200.models.dino.dino import build_dino
    torch.manual_seed(0)
    student = build_dino(mcase.small_args(device="cuda"))[0].cuda()
    teacher = EMA.ModelEMA(student, decay=0.9)
    sd = teacher.ema.state_dict()
    assert len({v.data_ptr() for v in sd.values()}) < len(sd)          # the state dict really aliases
    with torch.no_grad():
        for p in student.parameters():
            p.add_(torch.randn_like(p) * 0.1)
    want = {k: v.clone() for k, v in sd.items()}
    alias = {}
    for k, v in sd.items():                                            # clones must alias like the originals
        alias.setdefault(v.data_ptr(), want[k])
        want[k] = alias[v.data_ptr()]
    for _ in range(2):
        teacher.update(student)
        d = teacher.decay(teacher.updates)
        msd = student.state_dict()
        for k, v in want.items():                                      # models/dino/EMA.py:47-50
            if v.dtype.is_floating_point:
                v *= d
                v += (1.0 - d) * msd[k].detach()
    for k, v in teacher.ema.state_dict().items():
        assert torch.equal(v, want[k]), k
