"""Loader/builder of the C-ABI CUDA library (include/datr_msda.h).

The library is plain CUDA C++ (no torch headers) compiled in-tree for sm_100a:
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared ...
There is NO fallback: if the library cannot be built or loaded, every op raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SOURCES = [os.path.join(_PKG, "csrc", "msda.cu"), os.path.join(_PKG, "csrc", "linear_tf32.cu"),
           os.path.join(_PKG, "csrc", "layernorm.cu"), os.path.join(_PKG, "csrc", "colsum.cu"),
           os.path.join(_PKG, "csrc", "conv3x3_tf32.cu"), os.path.join(_PKG, "csrc", "wgrad_tf32.cu"),
           os.path.join(_PKG, "csrc", "rowmask.cu"), os.path.join(_PKG, "csrc", "attn_softmax.cu"),
           os.path.join(_PKG, "csrc", "attn_fused.cu"), os.path.join(_PKG, "csrc", "decoder_ops.cu"), os.path.join(_PKG, "csrc", "adamw.cu"), os.path.join(_PKG, "csrc", "lsa.cu"), os.path.join(_PKG, "csrc", "groupnorm.cu"),
           os.path.join(_PKG, "csrc", "ema.cu")]
INCLUDE_DIR = os.path.join(_ROOT, "include")
LIB_PATH = os.path.join(_PKG, "libdatr_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

EXPORTS = ("datr_msda_forward", "datr_msda_backward", "datr_msda_fused_forward", "datr_msda_fused_backward", "datr_last_error", "datr_abi_version",
           "datr_msda_backward_hs", "datr_msda_fused_backward_hs", "datr_msda_pack_value_pairs", "datr_msda_fused_forward_pairs", "datr_msda_set_backward_stages", "datr_msda_get_backward_stages",
           "datr_launch_count", "datr_linear_tf32", "datr_linear_tf32_bt", "datr_linear_tf32_bt_masked", "datr_linear_bf16", "datr_linear_wgrad_bf16", "datr_linear_last_error", "datr_linear_launch_count",
           "datr_layernorm256_forward", "datr_layernorm256_backward", "datr_layernorm_last_error", "datr_layernorm_launch_count",
           "datr_colsum", "datr_relu_bwd_colsum", "datr_colsum_last_error", "datr_colsum_launch_count",
           "datr_conv3x3_nhwc_tf32", "datr_conv3x3_wgrad_nhwc_tf32", "datr_conv_last_error", "datr_conv_launch_count",
           "datr_linear_wgrad_tf32", "datr_linear_wgrad_tf32_acc", "datr_linear_wgrad_bf16_acc", "datr_linear_wgrad_last_error", "datr_linear_wgrad_launch_count",
           "datr_zero_masked_rows", "datr_rowmask_last_error", "datr_rowmask_launch_count",
           "datr_attn_softmax_forward", "datr_attn_softmax_backward", "datr_attn_last_error", "datr_attn_launch_count",
           "datr_attn_mask_words", "datr_attn_pack_mask", "datr_attn_fused_forward", "datr_attn_fused_backward", "datr_attn_fused_last_error",
           "datr_attn_fused_launch_count", "datr_sine_embed", "datr_pos_embed_hw", "datr_bn_relu_maxpool_nhwc", "datr_decoder_ops_last_error", "datr_decoder_ops_launch_count",
           "datr_adamw_step", "datr_adamw_last_error", "datr_adamw_launch_count",
           "datr_lsa_solve", "datr_lsa_last_error", "datr_lsa_launch_count",
           "datr_groupnorm_nhwc_forward", "datr_groupnorm_nhwc_backward", "datr_groupnorm_last_error", "datr_groupnorm_launch_count",
           "datr_ema_update", "datr_ema_last_error", "datr_ema_launch_count")

_lock = threading.Lock()
_lib = None


class NativeLibraryError(RuntimeError):
    pass


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = SOURCES + [os.path.join(INCLUDE_DIR, f) for f in os.listdir(INCLUDE_DIR)] \
        + [os.path.join(_PKG, "csrc", f) for f in os.listdir(os.path.join(_PKG, "csrc")) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libdatr_b200.so in-tree (cross-compiles without a GPU).  Safe under one-process-per-GPU launches: the
    compile runs under an inter-process file lock, writes to a temporary file and renames it into place, so no rank can
    load a half-written library and only the first rank to get the lock compiles."""
    if not (force or _stale()):
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale():          # another process may have built it while we waited
                tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
                cmd = ["nvcc", *NVCC_FLAGS, f"-I{INCLUDE_DIR}", "-o", tmp, *SOURCES]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                res = subprocess.run(cmd, capture_output=True, text=True)
                if res.returncode != 0:
                    if os.path.exists(tmp):
                        os.unlink(tmp)
                    raise NativeLibraryError("nvcc failed:\n" + res.stdout + res.stderr)
                os.replace(tmp, LIB_PATH)
                if verbose:
                    print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; builds it first if the sources are newer.  Raises on any failure."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        try:
            path = build()
        except FileNotFoundError as e:  # nvcc missing: accept a prebuilt library, else fail loudly
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryError(f"libdatr_b200.so is missing and nvcc is unavailable: {e}") from e
            path = LIB_PATH
        try:
            L = ctypes.CDLL(path)
        except OSError as e:
            raise NativeLibraryError(f"cannot load {path}: {e}") from e
        vp, i, i64p = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
        L.datr_msda_forward.restype = i
        L.datr_msda_forward.argtypes = [vp, i64p, i64p, vp, vp, i, i, i, i, i, i, i, i, vp, vp]
        L.datr_msda_backward.restype = i
        L.datr_msda_backward.argtypes = [vp, i64p, i64p, vp, vp, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp]
        L.datr_msda_fused_forward.restype = i
        ll = ctypes.c_longlong
        L.datr_msda_fused_forward.argtypes = [vp, i64p, i64p, vp, ll, vp, ll, vp, i, i, i, i, i, i, i, i, i, vp, vp]
        L.datr_msda_fused_backward.restype = i
        L.datr_msda_fused_backward.argtypes = [vp, i64p, i64p, vp, ll, vp, ll, vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp]
        L.datr_msda_backward_hs.restype = i
        L.datr_msda_backward_hs.argtypes = [vp, i64p, i64p, i64p, i64p, vp, vp, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp]
        L.datr_msda_fused_backward_hs.restype = i
        L.datr_msda_fused_backward_hs.argtypes = [vp, i64p, i64p, i64p, i64p, vp, ll, vp, ll, vp, i, vp, i, i, i, i, i, i, i, i,
                                                  vp, vp, vp, vp]
        L.datr_msda_pack_value_pairs.restype = i
        L.datr_msda_pack_value_pairs.argtypes = [vp, i64p, i64p, i, i, i, i, i, vp, vp]
        L.datr_msda_fused_forward_pairs.restype = i
        L.datr_msda_fused_forward_pairs.argtypes = [vp, i, i64p, i64p, vp, ll, vp, ll, vp, i, i, i, i, i, i, i, i, vp, vp]
        L.datr_msda_set_backward_stages.restype = None
        L.datr_msda_set_backward_stages.argtypes = [i]
        L.datr_msda_get_backward_stages.restype = i
        L.datr_last_error.restype = ctypes.c_char_p
        L.datr_last_error.argtypes = []
        L.datr_abi_version.restype = i
        L.datr_launch_count.restype = ctypes.c_uint64
        L.datr_linear_tf32.restype = i
        L.datr_linear_tf32.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, vp]
        L.datr_linear_tf32_bt.restype = i
        L.datr_linear_tf32_bt.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, vp]
        L.datr_linear_tf32_bt_masked.restype = i
        L.datr_linear_tf32_bt_masked.argtypes = [vp, vp, vp, vp, vp, i, i, i, vp]
        L.datr_linear_bf16.restype = i
        L.datr_linear_bf16.argtypes = [vp, vp, vp, vp, i, vp, i, i, i, i, i, vp]
        L.datr_linear_wgrad_bf16.restype = i
        L.datr_linear_wgrad_bf16.argtypes = [vp, vp, vp, vp, i, i, i, vp]
        L.datr_linear_last_error.restype = ctypes.c_char_p
        L.datr_linear_launch_count.restype = ctypes.c_uint64
        L.datr_layernorm256_forward.restype = i
        L.datr_layernorm256_forward.argtypes = [vp, vp, vp, ctypes.c_float, vp, vp, vp, i, vp]
        L.datr_layernorm256_backward.restype = i
        L.datr_layernorm256_backward.argtypes = [vp] * 9 + [i, vp]
        L.datr_layernorm_last_error.restype = ctypes.c_char_p
        L.datr_layernorm_launch_count.restype = ctypes.c_uint64
        L.datr_colsum.restype = i
        L.datr_colsum.argtypes = [vp, vp, i, i, vp]
        L.datr_relu_bwd_colsum.restype = i
        L.datr_relu_bwd_colsum.argtypes = [vp, vp, vp, vp, i, i, vp]
        L.datr_colsum_last_error.restype = ctypes.c_char_p
        L.datr_colsum_launch_count.restype = ctypes.c_uint64
        L.datr_conv3x3_nhwc_tf32.restype = i
        L.datr_conv3x3_nhwc_tf32.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, vp]
        L.datr_conv3x3_wgrad_nhwc_tf32.restype = i
        L.datr_conv3x3_wgrad_nhwc_tf32.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, vp]
        L.datr_conv_last_error.restype = ctypes.c_char_p
        L.datr_conv_launch_count.restype = ctypes.c_uint64
        L.datr_linear_wgrad_tf32.restype = i
        L.datr_linear_wgrad_tf32.argtypes = [vp, vp, vp, vp, i, i, i, vp]
        for fn in (L.datr_linear_wgrad_tf32_acc, L.datr_linear_wgrad_bf16_acc):
            fn.restype = i
            fn.argtypes = [vp, vp, vp, vp, i, i, i, vp]
        L.datr_linear_wgrad_last_error.restype = ctypes.c_char_p
        L.datr_linear_wgrad_launch_count.restype = ctypes.c_uint64
        L.datr_zero_masked_rows.restype = i
        L.datr_zero_masked_rows.argtypes = [vp, vp, ctypes.c_longlong, i, vp]
        L.datr_rowmask_last_error.restype = ctypes.c_char_p
        L.datr_rowmask_launch_count.restype = ctypes.c_uint64
        L.datr_attn_softmax_forward.restype = i
        L.datr_attn_softmax_forward.argtypes = [vp, vp, ctypes.c_float, ctypes.c_longlong, i, i, vp]
        L.datr_attn_softmax_backward.restype = i
        L.datr_attn_softmax_backward.argtypes = [vp, vp, ctypes.c_float, ctypes.c_longlong, i, vp]
        L.datr_attn_last_error.restype = ctypes.c_char_p
        L.datr_attn_launch_count.restype = ctypes.c_uint64
        L.datr_attn_mask_words.restype = i
        L.datr_attn_mask_words.argtypes = [i]
        L.datr_attn_pack_mask.restype = i
        L.datr_attn_pack_mask.argtypes = [vp, i, vp, vp, vp]
        L.datr_attn_fused_backward.restype = i
        L.datr_attn_fused_backward.argtypes = [vp, ll, vp, ll, vp, ll, vp, vp, i, i, i, ctypes.c_float, vp, vp, vp, vp,
                                               vp, ll, vp, ll, vp, ll, vp]
        L.datr_attn_fused_forward.restype = i
        L.datr_attn_fused_forward.argtypes = [vp, ll, vp, ll, vp, ll, vp, i, i, i, ctypes.c_float, vp, vp, vp, vp]
        L.datr_attn_fused_last_error.restype = ctypes.c_char_p
        L.datr_attn_fused_launch_count.restype = ctypes.c_uint64
        L.datr_sine_embed.restype = i
        L.datr_sine_embed.argtypes = [vp, vp, ll, i, vp, vp]
        L.datr_pos_embed_hw.restype = i
        L.datr_pos_embed_hw.argtypes = [vp, vp, vp, vp, ll, i, vp, vp]
        L.datr_bn_relu_maxpool_nhwc.restype = i
        L.datr_bn_relu_maxpool_nhwc.argtypes = [vp, vp, vp, i, i, i, i, vp, vp]
        L.datr_decoder_ops_last_error.restype = ctypes.c_char_p
        L.datr_decoder_ops_launch_count.restype = ctypes.c_uint64
        L.datr_adamw_step.restype = i
        fl = ctypes.c_float
        L.datr_adamw_step.argtypes = [vp, vp, i, vp, fl, fl, fl, fl, fl, vp]
        L.datr_adamw_last_error.restype = ctypes.c_char_p
        L.datr_adamw_launch_count.restype = ctypes.c_uint64
        L.datr_lsa_solve.restype = i
        L.datr_lsa_solve.argtypes = [vp, vp, i, i, i, vp, vp]
        L.datr_lsa_last_error.restype = ctypes.c_char_p
        L.datr_lsa_launch_count.restype = ctypes.c_uint64
        L.datr_groupnorm_nhwc_forward.restype = i
        L.datr_groupnorm_nhwc_forward.argtypes = [vp, vp, vp, i, ll, i, i, ctypes.c_float, vp, vp, vp, vp, vp]
        L.datr_groupnorm_nhwc_backward.restype = i
        L.datr_groupnorm_nhwc_backward.argtypes = [vp, vp, vp, vp, vp, i, ll, i, i, vp, vp, vp, vp, vp]
        L.datr_groupnorm_last_error.restype = ctypes.c_char_p
        L.datr_groupnorm_launch_count.restype = ctypes.c_uint64
        L.datr_ema_update.restype = i
        L.datr_ema_update.argtypes = [vp, vp, i, ctypes.c_float, ctypes.c_float, vp]
        L.datr_ema_last_error.restype = ctypes.c_char_p
        L.datr_ema_launch_count.restype = ctypes.c_uint64
        if L.datr_abi_version() != 2:
            raise NativeLibraryError("libdatr_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def launch_count() -> int:
    """MSDeformAttn kernel launches issued through the library by this process."""
    return int(lib().datr_launch_count())


def colsum_launch_count() -> int:
    """Bias-gradient (column-sum) kernel launches issued through the library by this process."""
    return int(lib().datr_colsum_launch_count())


def all_launch_count() -> int:
    """Every hand-written kernel launch issued through the library by this process."""
    return (launch_count() + linear_launch_count() + layernorm_launch_count() + colsum_launch_count()
            + conv_launch_count() + wgrad_launch_count() + rowmask_launch_count() + attn_launch_count() + ema_launch_count()
            + int(lib().datr_decoder_ops_launch_count()) + int(lib().datr_adamw_launch_count())
            + int(lib().datr_lsa_launch_count()) + int(lib().datr_groupnorm_launch_count()))


def ema_launch_count() -> int:
    """Multi-tensor EMA kernel launches issued through the library by this process."""
    return int(lib().datr_ema_launch_count())


def attn_launch_count() -> int:
    """Attention softmax (forward / backward) kernel launches issued through the library by this process."""
    return int(lib().datr_attn_launch_count()) + int(lib().datr_attn_fused_launch_count())


def rowmask_launch_count() -> int:
    """Padding-mask (zero masked rows) kernel launches issued through the library by this process."""
    return int(lib().datr_rowmask_launch_count())


def wgrad_launch_count() -> int:
    """tcgen05 weight-gradient kernel launches issued through the library by this process."""
    return int(lib().datr_linear_wgrad_launch_count())


def conv_launch_count() -> int:
    """Implicit-GEMM 3x3 convolution kernel launches issued through the library by this process."""
    return int(lib().datr_conv_launch_count())


def layernorm_launch_count() -> int:
    """LayerNorm backward kernel launches issued through the library by this process."""
    return int(lib().datr_layernorm_launch_count())


def linear_launch_count() -> int:
    """tcgen05 linear kernel launches issued through the library by this process."""
    return int(lib().datr_linear_launch_count())
