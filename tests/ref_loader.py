"""TEST INFRASTRUCTURE: import the reference's own Python modules -- from /root/reference in the build container,
from the staged copy `baseline/_ref/` (oracle/stage_ref.py, git-ignored, travels with gpurun) on the GPU box -- so
that tests, the golden generators and the baseline tools can run the reference model itself.

Nothing is copied: the reference packages are imported under a private alias (`_ref_models`, `_ref_util`)
with stubs for the packages missing from this image (timm, the native MultiScaleDeformableAttention
extension) and with MSDeformAttnFunction.apply routed to the reference's own pure-PyTorch op
(ms_deform_attn_core_pytorch, func.py:41-61).  `.cuda()` / `.to('cuda')` calls hard-wired in the
reference's training path are neutralised while a reference call runs.
"""
import contextlib
import importlib
import os
import sys
import types

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference", os.path.join(_ROOT, "baseline", "_ref"))
REF = next((p for p in _CANDIDATES if os.path.isdir(os.path.join(p, "models", "dino"))), _CANDIDATES[0])


def available():
    return os.path.isdir(os.path.join(REF, "models", "dino"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load(cuda_ext=False):
    """Returns a namespace with the reference modules: .dino, .transformer, .backbone, .utils, .dn, .matcher,
    .da, .misc, .box_ops, .func, .msda_module.

    cuda_ext=False: MSDeformAttnFunction.apply is routed to the reference's pure-PyTorch op (CPU runs).
    cuda_ext=True : the reference's own autograd Function is left in place and its native module
                    `MultiScaleDeformableAttention` is the reference CUDA extension built unmodified by
                    oracle/build_ref.py (oracle/_ref/*.so) -- the stock GPU path of the reference."""
    global _loaded
    if _loaded is not None:
        assert _loaded.cuda_ext == cuda_ext, "the reference is already loaded in the other mode"
        return _loaded
    assert available(), "neither /root/reference nor baseline/_ref is present"
    ext = None
    if cuda_ext:
        if _ROOT not in sys.path:
            sys.path.insert(0, _ROOT)
        from oracle import build_ref
        ext = build_ref.load()
        assert ext is not None, "oracle/_ref/*.so (the reference CUDA extension) is not built"
    saved = {k: sys.modules.get(k) for k in ("models", "util", "MultiScaleDeformableAttention", "timm", "timm.models",
                                               "timm.models.layers")}
    saved_path = list(sys.path)
    # the reference imports `models.*` / `util.*` absolutely: expose /root/reference first
    for k in list(sys.modules):
        if k == "models" or k.startswith("models.") or k == "util" or k.startswith("util."):
            del sys.modules[k]
    sys.path.insert(0, REF)
    if ext is not None:
        sys.modules["MultiScaleDeformableAttention"] = ext
    else:
        _stub("MultiScaleDeformableAttention")
    ident = lambda *a, **k: None
    _stub("timm"); _stub("timm.models")
    _stub("timm.models.layers", DropPath=torch.nn.Identity, to_2tuple=lambda x: (x, x), trunc_normal_=ident)
    try:
        ns = types.SimpleNamespace()
        ns.func = importlib.import_module("models.dino.ops.functions.ms_deform_attn_func")
        ns.msda_module = importlib.import_module("models.dino.ops.modules.ms_deform_attn")
        ns.transformer = importlib.import_module("models.dino.deformable_transformer")
        ns.utils = importlib.import_module("models.dino.utils")
        ns.misc = importlib.import_module("util.misc")
        ns.box_ops = importlib.import_module("util.box_ops")
        ns.dn = importlib.import_module("models.dino.dn_components")
        ns.matcher = importlib.import_module("models.dino.matcher")
        ns.da = importlib.import_module("models.dino.DA_utils")
        ns.posenc = importlib.import_module("models.dino.position_encoding")
        ns.backbone = importlib.import_module("models.dino.backbone")
        ns.dino = importlib.import_module("models.dino.dino")
        ns.registry = importlib.import_module("models.registry")
        try:        # needs cv2 and torchvision (present in the build container)
            ns.selftrain = importlib.import_module("models.dino.self_training_utils")
        except ImportError:
            ns.selftrain = None
        core = ns.func.ms_deform_attn_core_pytorch

        class _CpuMSDA:
            @staticmethod
            def apply(value, shapes, level_start, loc, attn, im2col_step):
                return core(value, shapes, loc, attn)
        if not cuda_ext:
            ns.msda_module.MSDeformAttnFunction = _CpuMSDA
        ns.cuda_ext = cuda_ext
        ns.backbone.is_main_process = lambda: False          # no pretrained-weight download
    finally:
        # keep the reference modules alive under private names, give `models` / `util` back to the caller
        for k in list(sys.modules):
            if k == "models" or k.startswith("models.") or k == "util" or k.startswith("util."):
                sys.modules["_ref_" + k] = sys.modules.pop(k)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
        sys.path[:] = saved_path
    _loaded = ns
    return ns


@contextlib.contextmanager
def cpu_cuda_shim():
    """While active, Tensor.cuda() is a no-op and .to('cuda') / device='cuda' land on the CPU, so the
    reference's training path (dn_components.py:36-113, dino.py:106-107,790-818) runs without a GPU."""
    orig_cuda, orig_to, orig_mod_cuda = torch.Tensor.cuda, torch.Tensor.to, torch.nn.Module.cuda

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return orig_to(self, *a, **k)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.Tensor.to, torch.nn.Module.cuda = orig_cuda, orig_to, orig_mod_cuda
