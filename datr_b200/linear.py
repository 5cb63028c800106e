"""Tensor-core linear layers of the DINO hot path (host side of include/datr_linear.h).

`linear(x, weight, bias, relu=False, residual=None)` computes act(x @ weight.T + bias) + residual with the
tcgen05 / TMEM / TMA kernel of csrc/linear_tf32.cu (TF32 products, fp32 accumulation) and is differentiable:
  forward   y  = act(x W^T + b) + r                      one fused kernel
  backward  dz = dy * (y - r > 0)   (ReLU only)           elementwise
            dx = dz W                                     the same kernel on W^T (a [K, N] copy of the small weight)
            dW = dz^T x, db = sum_rows dz                 cuBLAS TF32 GEMM / reduction (plain library GEMM)
            dr = dy
It replaces torch.nn.functional.linear at the nn.Linear call sites of the reference's MSDeformAttn module
(models/dino/ops/modules/ms_deform_attn.py:94-125) and FFN (models/dino/deformable_transformer.py:784-805, :941-947).

The mode switch keeps the two numerics classes apart: "fp32" (default; SIMT fp32 library GEMMs, used by the strict
1e-3 parity tests) and "tf32" (tensor cores; the benchmark configuration, checked against a 1e-2 bar).
There is no fallback inside the "tf32" path: if the CUDA library is missing the call raises."""
from __future__ import annotations

import os as _os

import torch
import torch.nn.functional as F

from . import fallbacks, native

_MODE = "fp32"
_WGRAD = True          # weight/bias gradients on the tcgen05 kernel (False: cuBLAS GEMM + column-sum kernels)

# Optional launch timers: bench.py sets this to a list and each launch appends
# ("linear", (M, N, K, has_residual), start_event, end_event) recorded on the launching stream.
_timers = None


def set_mode(mode: str) -> None:
    """'fp32': every Linear runs as torch.nn.functional.linear (cuBLAS, TF32 off unless the caller enables it);
    'tf32': eligible Linears run on the tcgen05 kernel."""
    global _MODE
    if mode not in ("fp32", "tf32"):
        raise ValueError(mode)
    _MODE = mode


def get_mode() -> str:
    return _MODE


class fp32_products:
    """Context: library matmuls inside run with true fp32 products (cuBLAS SIMT), whatever the surrounding mode.  Used
    for the class-logit heads whose outputs feed index work -- the two-stage top-k over all encoder tokens
    (deformable_transformer.py:342) and the Hungarian cost matrix (matcher.py:47-95): given the same inputs these indices
    are then bit-identical to an fp32 run; 91-wide layers are a rounding error of the step's FLOPs."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


def eligible(x: torch.Tensor, weight: torch.Tensor) -> bool:
    N, K = weight.shape
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and K % 32 == 0 and N % 4 == 0
            and N >= 32 and x.numel() // K >= 1)


def _launch(x2, w, bias, residual2, relu):
    M, K = x2.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x2.device)
    lib = native.lib()
    with torch.cuda.device(x2.device):
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        rc = lib.datr_linear_tf32(x2.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                  residual2.data_ptr() if residual2 is not None else None, y.data_ptr(), M, N, K,
                                  int(relu), stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("linear", (M, N, K, residual2 is not None), e0, e1))
    if rc != 0:
        raise RuntimeError(f"datr_linear_tf32 failed (code {rc}): {lib.datr_linear_last_error().decode()}")
    return y


def _launch_bt(x2, w_t, bias, residual2, relu):
    """act(x2 @ w_t + bias) + residual for a weight given transposed, w_t [K, N] (datr_linear_tf32_bt): the input gradient
    of a Linear straight from its [N_fwd, K_fwd] weight, no transposed copy."""
    M, K = x2.shape
    N = w_t.shape[1]
    y = torch.empty((M, N), dtype=torch.float32, device=x2.device)
    lib = native.lib()
    with torch.cuda.device(x2.device):
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        rc = lib.datr_linear_tf32_bt(x2.data_ptr(), w_t.data_ptr(), bias.data_ptr() if bias is not None else None,
                                     residual2.data_ptr() if residual2 is not None else None, y.data_ptr(), M, N, K,
                                     int(relu), stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("linear", (M, N, K, residual2 is not None), e0, e1))
    if rc != 0:
        raise RuntimeError(f"datr_linear_tf32_bt failed (code {rc}): {lib.datr_linear_last_error().decode()}")
    return y


def _launch_bt_masked(x2, w_t, residual2, mask2):
    """(x2 @ w_t + residual2) where mask2 > 0, else 0 (datr_linear_tf32_bt_masked): input gradient + the gradient of a
    parallel branch + the ReLU backward of the tensor both branches read, in one kernel."""
    M, K = x2.shape
    N = w_t.shape[1]
    y = torch.empty((M, N), dtype=torch.float32, device=x2.device)
    lib = native.lib()
    with torch.cuda.device(x2.device):
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        rc = lib.datr_linear_tf32_bt_masked(x2.data_ptr(), w_t.data_ptr(), residual2.data_ptr() if residual2 is not None else None,
                                            mask2.data_ptr(), y.data_ptr(), M, N, K, stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("linear", (M, N, K, residual2 is not None), e0, e1))
    if rc != 0:
        raise RuntimeError(f"datr_linear_tf32_bt_masked failed (code {rc}): {lib.datr_linear_last_error().decode()}")
    return y


class GradCarrier:
    """Hands the gradient of a skip connection from the layer that closes it to the layer that opened it, inside one
    backward pass: y = relu(conv3(..conv1(x)..) + x) (a ResNet bottleneck).  The closing layer's backward deposits the
    gradient of its `residual` argument here instead of returning it to autograd; the opening layer's backward picks it up
    as the residual of its input-gradient GEMM, so `dx = dgrad + skip gradient` costs no pass of its own."""
    __slots__ = ("g",)

    def __init__(self):
        self.g = None


class GradChain:
    """Running sum of the input gradients of `n` layers that read the SAME tensor and whose backward passes run one after
    the other (decoder: the six value projections of the encoder memory, deformable_transformer.py:1011-1016 in the
    reference).  Each layer's input-gradient GEMM takes the sum so far as its residual; the layer that completes the count
    returns the total to autograd, the others return nothing -- n - 1 accumulation passes over the tensor disappear, and
    since none of these gradients is needed before the end of the backward pass, a captured backward runs the whole
    backward of such a layer on the parallel branch (SideWgrad)."""
    __slots__ = ("n", "seen", "g")

    def __init__(self, n):
        self.n, self.seen, self.g = n, 0, None


def _c(t):
    return t if t.is_contiguous() and t.data_ptr() % 16 == 0 else t.contiguous()


def _colsum(g2, y_act=None):
    """Bias gradient sum_rows g2 (csrc/colsum.cu).  With `y_act` the ReLU mask (y_act > 0) is applied first and the
    masked gradient is returned too: (dz, db)."""
    M, N = g2.shape
    lib = native.lib()
    db = torch.empty(N, dtype=torch.float32, device=g2.device)
    with torch.cuda.device(g2.device):
        stream = torch.cuda.current_stream().cuda_stream
        if y_act is None:
            rc = lib.datr_colsum(g2.data_ptr(), db.data_ptr(), M, N, stream)
            dz = g2
        else:
            dz = torch.empty_like(g2)
            rc = lib.datr_relu_bwd_colsum(g2.data_ptr(), y_act.data_ptr(), dz.data_ptr(), db.data_ptr(), M, N, stream)
    if rc != 0:
        raise RuntimeError(f"datr_colsum failed (code {rc}): {lib.datr_colsum_last_error().decode()}")
    return dz, db


def _dw_db(N, K, want_db, device):
    """dW [N, K] and db [N] carved from one buffer, db right behind dW: the library then zero-fills both with one memset."""
    if not want_db:
        return torch.empty((N, K), dtype=torch.float32, device=device), None
    buf = torch.empty(N * K + N, dtype=torch.float32, device=device)
    return buf[:N * K].view(N, K), buf[N * K:]


def _wgrad(g2, x2, want_db):
    """dW = g2^T x2 (and db = column sums of g2) on the tcgen05 weight-gradient kernel (csrc/wgrad_tf32.cu)."""
    M, N = g2.shape
    K = x2.shape[1]
    lib = native.lib()
    dw, db = _dw_db(N, K, want_db, g2.device)
    with torch.cuda.device(g2.device):
        rc = lib.datr_linear_wgrad_tf32(g2.data_ptr(), x2.data_ptr(), dw.data_ptr(), db.data_ptr() if want_db else None,
                                        M, N, K, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"datr_linear_wgrad_tf32 failed (code {rc}): {lib.datr_linear_wgrad_last_error().decode()}")
    return dw, db


class GradSinks:
    """Set as `linear._SINKS` by graphs._TrainingGraph while it captures a segment's backward: `map` is id(parameter) ->
    persistent gradient buffer (a FlatGradients view, zeroed once per step).  A Linear whose weight (and bias) ARE such
    parameters -- not slices, concatenations or products of them -- lets its weight-gradient kernel reduce straight into
    the buffers (datr_linear_wgrad_*_acc) and returns no gradient to autograd: no temporary, no zero fill, no accumulation
    pass, and nothing downstream that would have to wait for the kernel (it may run on the parallel branch)."""

    def __init__(self, mapping):
        self.map, self.used = mapping, []

    def lookup(self, weight, bias, want_gb):
        if not isinstance(weight, torch.nn.Parameter):
            return None
        sw = self.map.get(id(weight))
        if sw is None or sw.shape != weight.shape or not sw.is_contiguous() or sw.data_ptr() % 16:
            return None
        sb = None
        if want_gb:
            if not isinstance(bias, torch.nn.Parameter):
                return None
            sb = self.map.get(id(bias))
            if sb is None or sb.shape != bias.shape or not sb.is_contiguous() or sb.data_ptr() % 4:
                return None
        return sw, sb


_SINKS = None


def _wgrad_into(sinks, pairs, g2, x2, bf16=False):
    """dW (+ db) of one Linear reduced directly into its gradient buffers `pairs` = (sink_w, sink_b | None)."""
    sw, sb = pairs
    M, N = g2.shape
    K = x2.shape[1]
    lib = native.lib()
    fn = lib.datr_linear_wgrad_bf16_acc if bf16 else lib.datr_linear_wgrad_tf32_acc
    with torch.cuda.device(g2.device):
        rc = fn(g2.data_ptr(), x2.data_ptr(), sw.data_ptr(), sb.data_ptr() if sb is not None else None, M, N, K,
                torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"datr_linear_wgrad_*_acc failed (code {rc}): {lib.datr_linear_wgrad_last_error().decode()}")


def _wgrad_param(fn, g2, x2, want_db, weight, bias, bf16=False):
    """Weight / bias gradient of a Linear: into the parameters' gradient buffers when they are registered (returns
    (None, None): autograd gets nothing) -- then also eligible for the parallel branch --, else `fn` on the main branch."""
    sinks = _SINKS
    pairs = sinks.lookup(weight, bias, want_db) if sinks is not None and g2.is_cuda else None
    if pairs is None:
        return fn(g2, x2, want_db)
    side = _SIDE
    if side is not None and g2.shape[0] <= side.max_rows and torch.cuda.current_stream() != side.stream:
        side.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side.stream):
            _wgrad_into(sinks, pairs, g2, x2, bf16)
        # both operands were allocated on the main branch: their memory must not be handed out again while this branch reads it
        g2.record_stream(side.stream)
        x2.record_stream(side.stream)
        side.used += 1
    else:
        _wgrad_into(sinks, pairs, g2, x2, bf16)
    sinks.used.append((weight, pairs[0]))
    if pairs[1] is not None:
        sinks.used.append((bias, pairs[1]))
    return None, None


class SideWgrad:
    """Set as `linear._SIDE` by graphs._TrainingGraph while it captures a segment's backward: weight-gradient launches of
    at most `max_rows` rows that reduce straight into a parameter's gradient buffer (GradSinks) go to `stream`, a parallel
    branch of the captured graph.  At decoder sizes (4 400 rows) a weight-gradient kernel occupies a few dozen SMs and
    nothing in the rest of the pass reads its result; on one stream it still serialises with the input-gradient chain.
    Gradients that autograd post-processes (slices / concatenations / products of parameters, shared parameters summed by
    the engine) stay on the main branch: the engine runs those operations there without waiting for this stream.  The
    capture joins the branch at its end."""

    def __init__(self, stream, max_rows):
        self.stream, self.max_rows, self.used = stream, max_rows, 0


_SIDE = None
SIDE_WGRAD_MAX_ROWS = int(_os.environ.get("DATR_WGRAD_SIDE_ROWS", "16384"))      # 0: every launch stays on the main branch


def _zero_rows(t2, mask):
    from .rowmask import _zero
    _zero(t2, mask)


class _LinearTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, relu, zero_rows=None, skip_out=None, skip_in=None, mask_input_grad=False,
                grad_premasked=False, chain=None):
        """skip_out / skip_in (GradCarrier): this layer closes / opened a skip connection whose gradient travels through
        the carrier.  mask_input_grad: x is a ReLU output and the gradient returned for it is already multiplied by
        (x > 0) -- valid when EVERY consumer of x does so and the ReLU's own layer is told `grad_premasked` (relu == 2
        layers only), which then skips its mask pass."""
        ctx.skip_out, ctx.skip_in = skip_out if residual is not None else None, skip_in
        ctx.mask_input_grad, ctx.grad_premasked = bool(mask_input_grad), bool(grad_premasked)
        ctx.chain = chain               # GradChain: x's gradient is accumulated across the layers sharing the chain
        ctx.params = (weight, bias)     # the objects themselves: their registered gradient buffers are looked up by identity
        K = weight.shape[1]
        x2 = _c(x.reshape(-1, K))
        w = _c(weight)
        r2 = _c(residual.reshape(-1, weight.shape[0])) if residual is not None else None
        y = _launch(x2, w, _c(bias) if bias is not None else None, r2, relu)
        ctx.zero_rows = zero_rows
        if zero_rows is not None:      # padding rows of the fresh output -> 0 (csrc/rowmask.cu)
            _zero_rows(y, zero_rows)
        ctx.relu, ctx.has_bias, ctx.has_res = relu, bias is not None, residual is not None
        ctx.xshape = x.shape
        ctx.rshape = residual.shape if residual is not None else None
        ctx.save_for_backward(x2, w, y if relu else None, r2 if relu == 1 else None)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        side = _SIDE
        if (ctx.chain is None or side is None or not ctx.needs_input_grad[0] or not gy.is_cuda or _SINKS is None
                or _SINKS.lookup(ctx.params[0], ctx.params[1], ctx.has_bias and ctx.needs_input_grad[2]) is None):
            return _LinearTF32._backward(ctx, gy)
        # nothing in the rest of this backward pass waits for a chained layer: its whole backward joins the parallel branch
        main = torch.cuda.current_stream()
        side.stream.wait_stream(main)
        with torch.cuda.stream(side.stream):
            out = _LinearTF32._backward(ctx, gy)
        for t in (gy,) + tuple(ctx.saved_tensors) + (ctx.zero_rows,):
            if t is not None:
                t.record_stream(side.stream)
        side.used += 1
        return out

    @staticmethod
    def _backward(ctx, gy):
        x2, w, y, r2 = ctx.saved_tensors
        N, K = w.shape
        g2 = _c(gy.reshape(-1, N))
        if ctx.zero_rows is not None:
            # the masked output has one consumer (the MSDeformAttn op), whose backward allocates this gradient:
            # nobody else reads it, so its padding rows are zeroed in place
            _zero_rows(g2, ctx.zero_rows)
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        want_gw = ctx.needs_input_grad[1]
        fused = _WGRAD and want_gw and N % 4 == 0 and K % 4 == 0      # dW and db from one tensor-core kernel
        gb = None
        if ctx.relu == 2 and not ctx.grad_premasked:     # ReLU after the residual add: the mask applies to both branches
            if want_gb and not fused:
                g2, gb = _colsum(g2, y)
            else:
                g2 = torch.ops.aten.threshold_backward(g2, y, 0.0)
        gres = g2.view(ctx.rshape) if ctx.has_res else None
        if ctx.skip_out is not None:
            ctx.skip_out.g, gres = g2, None
        if ctx.relu == 1:     # ReLU mask (+ bias gradient) in one pass
            act = y if r2 is None else y - r2
            if want_gb and not fused:
                g2, gb = _colsum(g2, act)
            else:
                g2 = torch.ops.aten.threshold_backward(g2, act, 0.0)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            skip = None
            if ctx.skip_in is not None:
                skip, ctx.skip_in.g = ctx.skip_in.g, None
            if ctx.chain is not None:
                assert skip is None
                skip = ctx.chain.g
            if N % 32 == 0 and K % 4 == 0:
                if ctx.mask_input_grad:
                    gx = _launch_bt_masked(g2, w, skip, x2)
                else:
                    gx = _launch_bt(g2, w, None, skip, 0)
            else:
                fallbacks.note(f"torch.matmul (cuBLAS) dgrad N={N} K={K}")
                gx = g2 @ w
                if skip is not None:
                    gx = gx + skip
                if ctx.mask_input_grad:
                    gx = torch.ops.aten.threshold_backward(gx, x2, 0.0)
            gx = gx.view(ctx.xshape)
            if ctx.chain is not None:
                ch = ctx.chain
                ch.seen += 1
                if ch.seen == ch.n:     # the last layer of the chain returns the total
                    ch.seen, ch.g = 0, None
                else:
                    ch.g, gx = gx.view(-1, K), None
        if fused:
            gw, gb = _wgrad_param(_wgrad, g2, x2, want_gb, *ctx.params)
        else:
            if want_gb and gb is None:
                gb = _colsum(g2)[1]
            if want_gw:
                fallbacks.note(f"torch.matmul (cuBLAS) wgrad N={N} K={K}")
                gw = g2.t() @ x2
        return gx, gw, gb, gres, None, None, None, None, None, None, None


class _FFNTF32(torch.autograd.Function):
    """y = relu(x W1^T + b1) W2^T + b2 + x, the transformer FFN with its residual (reference
    deformable_transformer.py:801-805, :941-947), as ONE autograd node so that the backward fuses what autograd would
    run as separate passes over the [M, d_ffn] activation:
        dW2, db2 = wgrad(gy, h)                      dz1 = (gy W2) * (h > 0)      ReLU mask in the dgrad epilogue
        dW1, db1 = wgrad(dz1, x)                     dx  = dz1 W1 + gy            residual-branch add in the dgrad epilogue"""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        K = w1.shape[1]
        x2 = _c(x.reshape(-1, K))
        w1c, w2c = _c(w1), _c(w2)
        h = _launch(x2, w1c, _c(b1), None, 1)
        y = _launch(h, w2c, _c(b2), x2, 0)
        ctx.xshape = x.shape
        ctx.params = (w1, b1, w2, b2)
        ctx.save_for_backward(x2, w1c, w2c, h)
        return y.view(x.shape)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x2, w1, w2, h = ctx.saved_tensors
        g2 = _c(gy.reshape(-1, w2.shape[0]))
        w1p, b1p, w2p, b2p = ctx.params
        gw2, gb2 = _wgrad_param(_wgrad, g2, h, True, w2p, b2p)
        dz1 = _launch_bt(g2, w2, None, h, 3)
        gw1, gb1 = _wgrad_param(_wgrad, dz1, x2, True, w1p, b1p)
        gx = _launch_bt(dz1, w1, None, g2, 0).view(ctx.xshape)
        return gx, gw1, gb1, gw2, gb2


def _launch_bf16(xb, wb, bias, residual, relu, out_bf16, residual_bf16=False):
    """act(xb wb^T + bias) (+ residual) on the bf16-operand tensor-core kernel (datr_linear_bf16)."""
    M, K = xb.shape
    N = wb.shape[0]
    y = torch.empty((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=xb.device)
    lib = native.lib()
    with torch.cuda.device(xb.device):
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        rc = lib.datr_linear_bf16(xb.data_ptr(), wb.data_ptr(), bias.data_ptr() if bias is not None else None,
                                  residual.data_ptr() if residual is not None else None, int(residual_bf16), y.data_ptr(),
                                  int(out_bf16), M, N, K, int(relu), stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("linear_bf16", (M, N, K, residual is not None), e0, e1))
    if rc != 0:
        raise RuntimeError(f"datr_linear_bf16 failed (code {rc}): {lib.datr_linear_last_error().decode()}")
    return y


def _wgrad_bf16(gb, xb, want_db):
    """dW = gb^T xb (and db = column sums of gb), bf16 operands, fp32 results (datr_linear_wgrad_bf16)."""
    M, N = gb.shape
    K = xb.shape[1]
    lib = native.lib()
    dw, db = _dw_db(N, K, want_db, gb.device)
    with torch.cuda.device(gb.device):
        rc = lib.datr_linear_wgrad_bf16(gb.data_ptr(), xb.data_ptr(), dw.data_ptr(), db.data_ptr() if want_db else None,
                                        M, N, K, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"datr_linear_wgrad_bf16 failed (code {rc}): {lib.datr_linear_wgrad_last_error().decode()}")
    return dw, db


class _FFNBF16(torch.autograd.Function):
    """_FFNTF32 with bf16 operands: the six GEMMs of the block (two forward, two input-gradient, two weight-gradient) read
    bf16 activations / weights and accumulate in fp32; the [M, d_ffn] hidden activation and its gradient exist only as
    bf16 (half the HBM traffic of the block); x, y, the residual branch and every parameter gradient stay fp32.
    Precision class: BASELINE.json's bf16 bar (1e-2)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        K = w1.shape[1]
        x2 = _c(x.reshape(-1, K))
        xb = x2.to(torch.bfloat16)
        w1b, w2b = w1.to(torch.bfloat16), w2.to(torch.bfloat16)
        h = _launch_bf16(xb, w1b, _c(b1), None, 1, True)
        y = _launch_bf16(h, w2b, _c(b2), x2, 0, False)
        ctx.xshape = x.shape
        ctx.params = (w1, b1, w2, b2)
        ctx.save_for_backward(xb, w1b, w2b, h)
        return y.view(x.shape)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        xb, w1b, w2b, h = ctx.saved_tensors
        g2 = _c(gy.reshape(-1, w2b.shape[0]))
        gb = g2.to(torch.bfloat16)
        w1p, b1p, w2p, b2p = ctx.params
        gw2, gb2 = _wgrad_param(_wgrad_bf16, gb, h, True, w2p, b2p, bf16=True)
        dz1 = _launch_bf16(gb, w2b.t().contiguous(), None, h, 3, True, residual_bf16=True)
        gw1, gb1 = _wgrad_param(_wgrad_bf16, dz1, xb, True, w1p, b1p, bf16=True)
        gx = _launch_bf16(dz1, w1b.t().contiguous(), None, g2, 0, False).view(ctx.xshape)
        return gx, gw1, gb1, gw2, gb2


# Operand precision of the FFN blocks in 'tf32' mode: "bf16" (default) or "tf32" (DATR_FFN=tf32 / set_ffn_precision)
_FFN = _os.environ.get("DATR_FFN", "bf16")
_FFN_BF16_MIN_ROWS = 8192


def set_ffn_precision(kind: str) -> None:
    global _FFN
    if kind not in ("bf16", "tf32"):
        raise ValueError(kind)
    _FFN = kind


def ffn(x, w1, b1, w2, b2):
    """relu(x @ w1.T + b1) @ w2.T + b2 + x  (pre-norm output of the transformer FFN block)."""
    d_ffn, d = w1.shape
    if (_MODE == "tf32" and x.is_cuda and x.dtype == torch.float32 and w1.dtype == torch.float32 and b1 is not None
            and b2 is not None and tuple(w2.shape) == (d, d_ffn) and d % 32 == 0 and d_ffn % 32 == 0
            and torch.is_grad_enabled()):
        # bf16 operands pay off on long token sequences (encoder: 1.8x on the block at 89 k rows); at decoder sizes the
        # extra casts of the block outweigh the faster GEMMs (tools/bench_ffn.py)
        if _FFN == "bf16" and d % 128 == 0 and d_ffn % 256 == 0 and x.numel() // d >= _FFN_BF16_MIN_ROWS:
            return _FFNBF16.apply(x, w1, b1, w2, b2)
        return _FFNTF32.apply(x, w1, b1, w2, b2)
    return linear(linear(x, w1, b1, relu=True), w2, b2, residual=x)


def linear(x, weight, bias=None, relu=False, residual=None, zero_rows=None, skip_out=None, skip_in=None,
           mask_input_grad=False, grad_premasked=False, chain=None):
    """relu False/0: x @ weight.T + bias + residual;  True/1: relu(x @ weight.T + bias) + residual;
    2: relu(x @ weight.T + bias + residual).  Tensor-core kernel in 'tf32' mode for eligible shapes, else torch.
    `zero_rows` (bool [*x.shape[:-1]], True = padding; plain Linear only): those rows of the result are set to zero,
    as `masked_fill(mask[..., None], 0)` after the Linear would.
    skip_out / skip_in / mask_input_grad / grad_premasked / chain: backward-pass fusion of skip connections, ReLU masks and
    shared-input gradient sums (_LinearTF32.forward, GradChain); tensor-core path only -- the caller checks `eligible`."""
    relu = int(relu)
    if skip_out is not None or skip_in is not None or mask_input_grad or grad_premasked or chain is not None:
        if not (_MODE == "tf32" and eligible(x, weight)):
            raise ValueError("backward-pass fusion flags need the tensor-core path")
        if zero_rows is not None:
            if relu or residual is not None or zero_rows.dtype != torch.bool or zero_rows.shape != x.shape[:-1]:
                raise ValueError("zero_rows applies to a plain Linear")
            zero_rows = zero_rows.contiguous()
        return _LinearTF32.apply(x, weight, bias, residual, relu, zero_rows, skip_out, skip_in, mask_input_grad, grad_premasked,
                                 chain)
    if zero_rows is not None:
        if relu or residual is not None:
            raise ValueError("zero_rows applies to a plain Linear")
        if _MODE == "tf32" and eligible(x, weight) and zero_rows.dtype == torch.bool and zero_rows.shape == x.shape[:-1]:
            return _LinearTF32.apply(x, weight, bias, None, 0, zero_rows.contiguous())
        from .rowmask import zero_masked_rows
        return zero_masked_rows(F.linear(x, weight, bias), zero_rows)
    if _MODE == "tf32" and eligible(x, weight):
        return _LinearTF32.apply(x, weight, bias, residual, relu)
    if _MODE == "tf32" and x.is_cuda:
        fallbacks.note(f"F.linear (cuBLAS) N={weight.shape[0]} K={weight.shape[1]}: outside the tcgen05 kernel's shapes (N % 4, N >= 32, K % 32)")
    y = F.linear(x, weight, bias)
    if relu == 1:
        y = F.relu(y)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu == 2 else y
