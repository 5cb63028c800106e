"""Generate tests/golden/msda_golden.npz from the REFERENCE's own code.

Runs only in the build container (needs /root/reference).  It imports the
reference's ms_deform_attn_core_pytorch (models/dino/ops/functions/
ms_deform_attn_func.py:41-61) unmodified -- the module-level
`import MultiScaleDeformableAttention` (:18) is satisfied by an empty stub --
and records, for the seeded inputs of tests/msda_cases.py:
  * forward outputs (fp64 and fp32),
  * autograd gradients w.r.t. value / sampling_locations / attention_weights
    for a fixed grad_output (fp64 and fp32),
  * the reference's own test recipe (ops/test.py:21-60: seed 3, draw order
    value, loc, attn; double case first, float case second) with its inputs,
  * for the config-1 shapes (800x800, S=13294; encoder Lq=S and decoder
    Lq=900): checksums + a strided subsample instead of full tensors.

Usage:  python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import msda_cases as mc  # noqa: E402

REF = "/root/reference/models/dino/ops/functions/ms_deform_attn_func.py"
STRIDE = 997  # subsample stride for big tensors


def load_reference_fn():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("_ref_msda_func", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def run_ref(fn, inp, dtype):
    t = {k: torch.from_numpy(np.asarray(v)) for k, v in inp.items()}
    value = t["value"].to(dtype).requires_grad_(True)
    loc = t["loc"].to(dtype).requires_grad_(True)
    attn = t["attn"].to(dtype).requires_grad_(True)
    out = fn(value, t["shapes"], loc, attn)
    out.backward(t["grad_out"].to(dtype).view_as(out))
    return (out.detach().numpy(), value.grad.numpy(), loc.grad.numpy(), attn.grad.numpy())


def digest(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum(), np.abs(a).max()], dtype=np.float64)


def main():
    fn = load_reference_fn()
    G = {}

    # --- the reference's own test recipe (ops/test.py) ---------------------------------
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
        out = fn(value.to(dt), shapes, loc.to(dt), attn.to(dt))
        G[f"kat_{tag}_value"], G[f"kat_{tag}_loc"], G[f"kat_{tag}_attn"] = value.numpy(), loc.numpy(), attn.numpy()
        G[f"kat_{tag}_out"] = out.numpy()
    G["kat_shapes"] = shapes.numpy()

    # --- seeded small cases: full tensors ---------------------------------------------
    for case in mc.SMALL_CASES:
        name = case[0]
        for tag, dt, npdt in (("f64", torch.float64, np.float64), ("f32", torch.float32, np.float32)):
            inp = mc.small_case(name, npdt)
            out, gv, gl, ga = run_ref(fn, inp, dt)
            G[f"{name}_{tag}_out"], G[f"{name}_{tag}_gv"] = out, gv
            G[f"{name}_{tag}_gl"], G[f"{name}_{tag}_ga"] = gl, ga

    # --- config-1 shapes: digests + strided subsample (fp32) ----------------------------
    for name, Lq, mode, seed in (("cfg1_enc", -1, "encoder", 101), ("cfg1_dec", 900, "uniform", 102)):
        inp = mc.make_inputs(1, 8, 32, Lq, 4, mc.CFG1_LEVELS, mode, seed, np.float32)
        res = run_ref(fn, inp, torch.float32)
        for key, arr in zip(("out", "gv", "gl", "ga"), res):
            G[f"{name}_f32_{key}_digest"] = digest(arr)
            G[f"{name}_f32_{key}_sub"] = arr.reshape(-1)[::STRIDE].copy()

    G["meta_torch_version"] = np.array(torch.__version__)
    G["meta_stride"] = np.array(STRIDE)
    out_path = os.path.join(HERE, "msda_golden.npz")
    np.savez_compressed(out_path, **G)
    print("wrote", out_path, os.path.getsize(out_path), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
