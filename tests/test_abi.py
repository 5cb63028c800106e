"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute calls here (no GPU); argument validation paths return before touching CUDA."""
import ctypes
import glob
import os
import re

import pytest
import torch

from datr_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(datr_\w+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    assert "datr_msda_forward" in names and "datr_msda_backward" in names


def test_library_builds_loads_and_exports_all_declared_symbols():
    lib = native.lib()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(native.EXPORTS) <= set(declared_functions())
    assert lib.datr_abi_version() == 2


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_bad_arguments_return_error_codes_without_touching_the_gpu():
    lib = native.lib()
    rc = lib.datr_msda_forward(None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 0, None, None)
    assert rc == -1 and b"null" in lib.datr_last_error()
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    rc = lib.datr_msda_forward(p, p, p, p, p, 0, 1, 1, 1, 1, 1, 1, 0, p, None)
    assert rc == -1 and b"positive" in lib.datr_last_error()
    rc = lib.datr_msda_forward(p, p, p, p, p, 1, 1, 1, 1, 1, 1, 1, 7, p, None)
    assert rc == -1 and b"dtype" in lib.datr_last_error()
    rc = lib.datr_msda_backward(p, p, p, p, p, None, 1, 1, 1, 1, 1, 1, 1, 0, p, p, p, None)
    assert rc == -1
    rc = lib.datr_msda_forward(p + 2, p, p, p, p, 1, 1, 1, 1, 1, 1, 1, 0, p, None)
    assert rc == -2
    # fused entry points: reference_points arity, unsupported configuration, row strides
    rc = lib.datr_msda_fused_forward(p, p, p, p, 0, p, 0, p, 3, 1, 1, 1, 32, 1, 1, 4, 0, p, None)
    assert rc == -1 and b"2 or 4" in lib.datr_last_error()
    rc = lib.datr_msda_fused_forward(p, p, p, p, 0, p, 0, p, 2, 1, 1, 1, 16, 1, 1, 4, 0, p, None)
    assert rc == -4
    rc = lib.datr_msda_fused_backward(p, p, p, p, 5, p, 0, p, 2, p, 1, 1, 8, 32, 4, 1, 4, 0, p, p, p, None)
    assert rc == -1 and b"row strides" in lib.datr_last_error()
    assert lib.datr_attn_softmax_forward(None, None, 1.0, 1, 1, 1, None) == -1
    assert lib.datr_attn_softmax_forward(p, None, 1.0, 1, 4096, 1, None) == -1
    assert lib.datr_ema_update(None, None, 1, 0.5, 0.5, None) == -1
    assert lib.datr_zero_masked_rows(p, p, 4, 3, None) == -1
    assert lib.datr_layernorm256_forward(p, p, p, 1e-5, p, p, p, 0, None) == -1
    # round-2 entry points
    assert lib.datr_attn_fused_forward(None, 512, None, 512, None, 256, None, 1, 8, 16, 0.1, None, None, None, None) == -1
    assert lib.datr_attn_fused_forward(p, 100, p, 512, p, 256, p, 1, 8, 16, 0.1, p, None, None, None) == -1       # row stride < H * 32
    assert b"row strides" in lib.datr_attn_fused_last_error()
    assert lib.datr_attn_pack_mask(None, 0, None, None, None) == -1
    assert lib.datr_attn_mask_words(1100) == 36 and lib.datr_attn_mask_words(128) == 4
    assert lib.datr_linear_bf16(p, p, None, None, 0, p, 0, 8, 64, 96, 0, None) == -1 and b"64" in lib.datr_linear_last_error()
    assert lib.datr_linear_bf16(p, p, None, p, 0, p, 1, 8, 256, 64, 0, None) == -1                               # bf16 output + fp32 residual
    assert lib.datr_linear_wgrad_bf16(None, p, p, None, 8, 64, 64, None) == -1
    # backward-pass entry points of DESIGN.md 4.15
    assert lib.datr_linear_tf32_bt_masked(p, p, None, None, p, 128, 64, 32, None) == -1 and b"mask" in lib.datr_linear_last_error()
    assert lib.datr_linear_tf32_bt_masked(None, p, None, p, p, 128, 64, 32, None) == -1
    assert lib.datr_linear_tf32_bt_masked(p, p, None, p, p, 128, 64, 33, None) == -1 and b"32" in lib.datr_linear_last_error()
    assert lib.datr_linear_tf32_bt(p, p, None, None, p, 128, 64, 32, 4, None) == -1            # relu 4 is the masked entry point's
    assert lib.datr_linear_wgrad_tf32_acc(None, p, p, None, 8, 64, 64, None) == -1
    assert lib.datr_linear_wgrad_tf32_acc(p, p, p, None, 8, 62, 64, None) == -1 and b"multiples of 4" in lib.datr_linear_wgrad_last_error()
    assert lib.datr_linear_wgrad_bf16_acc(p, p, p, None, 8, 60, 64, None) == -1 and b"multiples of 8" in lib.datr_linear_wgrad_last_error()
    assert lib.datr_sine_embed(p, p, 4, 3, p, None) == -1
    assert lib.datr_adamw_step(None, None, 1, None, 0.9, 0.999, 1e-8, 0.1, 0.03, None) == -1
    assert lib.datr_adamw_step(p, p, 1, None, 0.9, 0.999, 1e-8, 0.0, 0.03, None) == -1                           # bias correction must be > 0
    assert lib.datr_lsa_solve(None, None, 1, 900, 10, None, None) == -1
    assert lib.datr_lsa_solve(p, p, 1, 10, 900, p, None) == -4 and b"more boxes" in lib.datr_lsa_last_error()
    assert lib.datr_msda_pack_value_pairs(p, p, p, 1, 1, 1, 1, 7, p, None) == -1 and b"storage" in lib.datr_last_error()
    lib.datr_msda_set_backward_stages(2)
    assert lib.datr_msda_get_backward_stages() == 2
    lib.datr_msda_set_backward_stages(-1)


def test_shim_refuses_cpu_tensors_like_the_reference():
    """ms_deform_attn.h:38,60 -> 'Not implemented on the CPU'; there is no CPU fallback."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    from datr_b200.models.dino.ops.functions import MSDeformAttnFunction
    v = torch.zeros(1, 5, 2, 4)
    shapes = torch.tensor([[1, 5]])
    start = torch.tensor([0])
    loc = torch.zeros(1, 3, 2, 1, 2, 2)
    attn = torch.zeros(1, 3, 2, 1, 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_forward(v, shapes, start, loc, attn, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_backward(v, shapes, start, loc, attn, torch.zeros(1, 3, 8), 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDeformAttnFunction.apply(v, shapes, start, loc, attn, 64)


def test_dropin_module_name():
    import sys
    from datr_b200 import MultiScaleDeformableAttention as shim
    shim.install()
    import MultiScaleDeformableAttention as MSDA
    assert MSDA.ms_deform_attn_forward is shim.ms_deform_attn_forward
    assert MSDA.ms_deform_attn_backward is shim.ms_deform_attn_backward
    del sys.modules["MultiScaleDeformableAttention"]


def test_registry_serves_the_reference_entry_point_the_way_main_py_uses_it():
    """main.py:79-85 (build_model_main): membership test on the private dict, then .get(modelname)(args)."""
    import datr_b200
    datr_b200.install_dropin()
    from models.registry import MODULE_BUILD_FUNCS
    import models  # noqa: F401  (importing the package registers 'dino', like the reference's models/__init__.py:8)
    assert "dino" in MODULE_BUILD_FUNCS._module_dict and MODULE_BUILD_FUNCS._name == MODULE_BUILD_FUNCS.name
    assert MODULE_BUILD_FUNCS.get("dino").__name__ == "build_dino" and MODULE_BUILD_FUNCS.get("missing") is None
    assert len(MODULE_BUILD_FUNCS) >= 1 and "dino" in repr(MODULE_BUILD_FUNCS)
    import pytest as _pytest
    with _pytest.raises(KeyError):
        MODULE_BUILD_FUNCS.register(MODULE_BUILD_FUNCS.get("dino"), module_name="dino")
    with _pytest.raises(TypeError):
        MODULE_BUILD_FUNCS.register(object())
