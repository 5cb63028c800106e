"""GPU parity of the bias-gradient kernels (include/datr_colsum.h) against fp64 torch on the same inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(44446, 256), (44446, 2048), (2200, 128), (7, 4), (1, 384), (100000, 64), (333, 1028)])
def test_colsum_and_relu_bwd(rows, cols):
    from datr_b200 import native
    from datr_b200.linear import _colsum
    g = torch.Generator(device="cpu").manual_seed(rows + cols)
    dy = torch.randn(rows, cols, generator=g).cuda()
    y = torch.randn(rows, cols, generator=g).cuda()
    n0 = native.colsum_launch_count()
    _, db = _colsum(dy)
    dz, db2 = _colsum(dy, y)
    torch.cuda.synchronize()
    assert native.colsum_launch_count() == n0 + 2
    want = dy.double().sum(0)
    assert float((db.double() - want).abs().max()) < 1e-4 * max(1.0, float(want.abs().max()))
    mask = y > 0
    assert torch.equal(dz, dy * mask)                      # bit-exact elementwise
    want2 = (dy.double() * mask).sum(0)
    assert float((db2.double() - want2).abs().max()) < 1e-4 * max(1.0, float(want2.abs().max()))
