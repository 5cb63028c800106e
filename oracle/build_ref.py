"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's own CUDA extension (models/dino/ops/src/**, unmodified,
from where it lies under /root/reference) for sm_100a into
oracle/_ref/MultiScaleDeformableAttention_ref*.so so that the GPU tests can
cross-check our kernels against the reference kernels themselves and bench.py can
time them as the GPU baseline.  Nothing from the reference is copied into the repo.

The reference does not compile against torch >= 2.x as is: two
`AT_DISPATCH_FLOATING_TYPES(value.type(), ...)` sites
(cuda/ms_deform_attn_cuda.cu:64,134) pass a DeprecatedTypeProperties where the
macro now calls `::detail::scalar_type(ScalarType)`.  Instead of patching the
sources we force-include `ref_compat.h`, which adds the removed overload.
The reference's own setup.py refuses to build without a visible GPU
(ops/setup.py:48-49), so we drive nvcc/g++ directly.

Runs only where /root/reference exists (the build container); the GPU box uses
the prebuilt .so that travels with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/models/dino/ops/src"
MODNAME = "MultiScaleDeformableAttention_ref"


def so_path() -> str:
    return os.path.join(OUT, MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False) -> str | None:
    """Returns the path of the built module, or None when /root/reference is absent."""
    target = so_path()
    if not os.path.isdir(REF_SRC):
        return target if os.path.exists(target) else None
    if os.path.exists(target) and not force:
        return target
    import torch
    from torch.utils.cpp_extension import include_paths, library_paths

    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{p}" for p in include_paths("cuda")] + [f"-I{REF_SRC}", f"-I{sysconfig.get_paths()['include']}"]
    defs = ["-DWITH_CUDA", f"-DTORCH_EXTENSION_NAME={MODNAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    compat = os.path.join(HERE, "ref_compat.h")
    objs = []
    cu = os.path.join(REF_SRC, "cuda", "ms_deform_attn_cuda.cu")
    o = os.path.join(OUT, "ms_deform_attn_cuda.o")
    subprocess.run(["nvcc", "-c", cu, "-o", o, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-Xcompiler", "-fPIC", "-include", compat, "--expt-relaxed-constexpr", "-w",
                    "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                    "-D__CUDA_NO_HALF2_OPERATORS__"] + inc + defs, check=True)
    objs.append(o)
    for name in ("vision.cpp", os.path.join("cpu", "ms_deform_attn_cpu.cpp")):
        src = os.path.join(REF_SRC, name)
        o = os.path.join(OUT, os.path.basename(name).replace(".cpp", ".o"))
        subprocess.run(["g++", "-c", src, "-o", o, "-O2", "-std=c++17", "-fPIC", "-w", "-include", compat]
                       + inc + defs, check=True)
        objs.append(o)
    libs = [f"-L{p}" for p in library_paths("cuda")] + ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python",
                                                          "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    rpath = [f"-Wl,-rpath,{p}" for p in library_paths("cuda")]
    subprocess.run(["g++", "-shared", "-o", target] + objs + libs + rpath, check=True)
    return target


def load():
    """Import the built reference extension (needs a CUDA-capable torch to be useful)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    path = so_path()
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(MODNAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
