"""CPU: pin the oracle (oracle/msda.py, oracle/msda_oracle.c) against golden vectors produced by the
reference's own ms_deform_attn_core_pytorch (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import msda_cases as mc
from oracle import msda as om

TOL = {"f64": 1e-12, "f32": 2e-5}
DT = {"f64": np.float64, "f32": np.float32}


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_reference_known_answer_recipe(golden, tag):
    """ops/test.py:21-60 recipe (seed 3; SURVEY.md §4 lists the first values)."""
    dt = DT[tag]
    v, l, a = (golden[f"kat_{tag}_{k}"].astype(dt) for k in ("value", "loc", "attn"))
    want = golden[f"kat_{tag}_out"]
    got_c = om.fwd(v, golden["kat_shapes"], None, l, a)
    got_t = om.core_torch(torch.from_numpy(v), golden["kat_shapes"], torch.from_numpy(l), torch.from_numpy(a)).numpy()
    assert mc.rel_err(got_c, want) < TOL[tag]
    assert mc.rel_err(got_t, want) < TOL[tag]
    if tag == "f64":  # the value quoted in SURVEY.md §4
        assert abs(want.reshape(-1)[0] - 0.0018993784157779181) < 1e-15


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name", [c[0] for c in mc.SMALL_CASES])
def test_c_oracle_matches_reference_goldens(golden, name, tag):
    i = mc.small_case(name, DT[tag])
    out = om.fwd(i["value"], i["shapes"], i["level_start"], i["loc"], i["attn"])
    gv, gl, ga = om.bwd(i["value"], i["shapes"], i["level_start"], i["loc"], i["attn"], i["grad_out"])
    for got, key in ((out, "out"), (gv, "gv"), (gl, "gl"), (ga, "ga")):
        assert mc.rel_err(got, golden[f"{name}_{tag}_{key}"]) < TOL[tag], key


@pytest.mark.parametrize("name", ["d32_l4_outside", "d32_l5", "d71_generic"])
def test_torch_port_matches_reference_goldens(golden, name):
    i = {k: torch.from_numpy(v) for k, v in mc.small_case(name, np.float64).items()}
    value, loc, attn = (i[k].clone().requires_grad_(True) for k in ("value", "loc", "attn"))
    out = om.core_torch(value, i["shapes"], loc, attn)
    out.backward(i["grad_out"].view_as(out))
    assert mc.rel_err(out.detach().numpy(), golden[f"{name}_f64_out"]) < 1e-13
    assert mc.rel_err(value.grad.numpy(), golden[f"{name}_f64_gv"]) < 1e-13
    assert mc.rel_err(loc.grad.numpy(), golden[f"{name}_f64_gl"]) < 1e-13
    assert mc.rel_err(attn.grad.numpy(), golden[f"{name}_f64_ga"]) < 1e-13


@pytest.mark.parametrize("name,Lq,mode,seed", [("cfg1_enc", -1, "encoder", 101), ("cfg1_dec", 900, "uniform", 102)])
def test_c_oracle_config1_shapes(golden, name, Lq, mode, seed):
    """BASELINE.json configs[0]: 1 image 800x800, 4 levels, 8 heads, 4 points (S=13294)."""
    i = mc.make_inputs(1, 8, 32, Lq, 4, mc.CFG1_LEVELS, mode, seed, np.float32)
    stride = int(golden["meta_stride"])
    out = om.fwd(i["value"], i["shapes"], i["level_start"], i["loc"], i["attn"])
    gv, gl, ga = om.bwd(i["value"], i["shapes"], i["level_start"], i["loc"], i["attn"], i["grad_out"])
    for got, key in ((out, "out"), (gv, "gv"), (gl, "gl"), (ga, "ga")):
        sub = golden[f"{name}_f32_{key}_sub"]
        assert mc.rel_err(got.reshape(-1)[::stride], sub) < 5e-5, key
        dg = golden[f"{name}_f32_{key}_digest"]
        a = got.astype(np.float64).reshape(-1)
        assert abs(np.abs(a).sum() - dg[1]) / dg[1] < 1e-5, key
        assert abs((a * a).sum() - dg[2]) / dg[2] < 1e-5, key


def test_empty_contribution_and_edges():
    """All samples outside every map -> exact zeros everywhere (validity guard, cuh:288)."""
    i = mc.make_inputs(1, 2, 4, 3, 2, [(3, 3), (2, 2)], "uniform", 5, np.float64)
    i["loc"][:] = -7.0
    out = om.fwd(i["value"], i["shapes"], i["level_start"], i["loc"], i["attn"])
    gv, gl, ga = om.bwd(i["value"], i["shapes"], i["level_start"], i["loc"], i["attn"], i["grad_out"])
    assert not out.any() and not gv.any() and not gl.any() and not ga.any()
    # a sample exactly on a pixel centre returns that pixel times the weight
    i["loc"][:] = -7.0
    i["loc"][0, 0, 0, 0, 0] = [(1 + 0.5) / 3, (2 + 0.5) / 3]   # x=1, y=2 on the 3x3 map
    out = om.fwd(i["value"], i["shapes"], i["level_start"], i["loc"], i["attn"])
    want = i["value"][0, 2 * 3 + 1, 0] * i["attn"][0, 0, 0, 0, 0]
    np.testing.assert_allclose(out[0, 0, :4], want, rtol=1e-13)
