"""CUDA graphs for the static segments of the DINO training step.

Launched kernel by kernel the step is bound by the host (tools/segment_times.py: ~10 k launches, 47 ms of pure enqueue time
for the forward alone).  The ResNet body, the input projections, the level flattening, the deformable encoder, the two-stage
query selection, the decoder, the prediction heads, the image-level discriminator and the criterion's losses are pure tensor
functions with static shapes for a fixed image size / number of de-noising queries / number of boxes, so each is captured ONCE
-- forward and backward, following the protocol of torch.cuda.make_graphed_callables (_TrainingGraph below) -- and replayed
afterwards: no Python, no per-kernel launch cost.  What stays eager in between depends on the targets (de-noising query
construction, the matcher's cost matrices and the GPU Hungarian solver, prototype bookkeeping).

What the captured backward does beyond torch's version: parameter gradients go INTO the step's flat gradient
buffer inside the graph (`set_grad_sinks`), small weight-gradient GEMMs fork onto a parallel branch, and a segment can be
replayed on a second stream beside its neighbours (`call(..., side=True)` / `join_side()`).

Precondition of the first (capturing) call, as for torch's own make_graphed_callables: no autograd graph of an EARLIER eager
pass over the same parameters may still be alive (a kept loss tensor is enough) -- its AccumulateGrad nodes belong to the
default stream, and the capture fails loudly with cudaErrorStreamCaptureImplicit rather than make that stream wait.

Usage (what bench_dino.DinoStep does):
    graphs.ACTIVE = graphs.StepGraphs()      # opt in
    ...each step:  graphs.ACTIVE.begin_step(); loss = criterion(model(...)); loss.backward()

A segment is keyed by (name, index of the call inside the step, shapes/dtypes of its tensor arguments): the two
transformer passes of a step get separate graphs because the activations of the first must survive until its
backward.  A new shape signature simply captures another graph.  Capture failures are raised, never hidden.

Our own kernels are capture-safe: they run on the current stream, allocate through torch's caching allocator (the
graph's private pool during capture), pass tensor maps by value and use cudaMemsetAsync for the zero fill.
"""
from __future__ import annotations

import torch
from torch import nn

from . import native

ACTIVE = None          # a StepGraphs instance when graph replay is enabled


def _native_launches():
    return native.all_launch_count()


class StepGraphs:
    def __init__(self, warmup_iters: int = 3):
        self.cache = {}
        self.calls = {}
        self.warmup_iters = warmup_iters
        self.replayed_native_launches = 0      # hand-written kernel launches executed by graph replays
        self.captures = 0
        self.per_segment = {}                  # (name, index) -> number of signatures captured so far
        self.max_signatures = 4                # beyond that a segment runs eagerly for unseen signatures
        self.capture_inference = True          # no-grad passes get forward-only graphs (False: they run eagerly)
        import os
        self.skip = set(filter(None, os.environ.get("DATR_GRAPH_SKIP", "").split(",")))   # segment names kept eager
        self.alias_inputs = os.environ.get("DATR_GRAPH_ALIAS", "1") != "0"
        self.stable_storages = set()           # storage addresses of the static outputs of the captured training segments
        # id(parameter) -> persistent gradient buffer (a FlatGradients view): the captured backward of a segment adds the
        # parameter gradients into these buffers itself (one multi-tensor launch inside the graph) and hands autograd
        # nothing, instead of one eager `grad += new` kernel per parameter after every replay (DATR_GRAPH_SINKS=0: off)
        self.grad_sinks = {}
        self.side_segments = os.environ.get("DATR_SIDE_SEGMENTS", "1") != "0"
        self._side_stream, self._side_pending = None, False
        self.sink_grads = os.environ.get("DATR_GRAPH_SINKS", "1") != "0"

    def begin_step(self):
        self.calls.clear()
        self.join_side()

    def set_grad_sinks(self, flat):
        """Register the gradient buffers of `flat` (datr_b200.parallel.FlatGradients, gather=False: every .grad is a
        persistent view, zeroed once per step) as the destinations of the in-graph gradient accumulation.  Call it before
        the first training step; segments captured earlier keep handing their gradients to autograd."""
        if getattr(flat, "gather", False):
            return
        moved = any(id(p) in self.grad_sinks and self.grad_sinks[id(p)].data_ptr() != v.data_ptr()
                    for p, v in zip(flat.params, flat.views))
        if moved:
            # a new gradient buffer for parameters whose segments were captured against the old one: those captures add
            # into memory nobody reads any more, so the training graphs are dropped and captured again on next use
            self.grad_sinks.clear()
            for k in [k for k in self.cache if k[2]]:
                del self.cache[k]
                self.per_segment[k[:2]] = max(0, self.per_segment.get(k[:2], 1) - 1)
        for p, v in zip(flat.params, flat.views):
            self.grad_sinks[id(p)] = v

    def call(self, name, owner, fn, *args, side=False):
        """Run `fn(*args)` (args: any pytree of tensors and hashable constants; `owner`: the nn.Module whose
        parameters fn uses, or None) as a graph segment; returns fn's output pytree.
        `side`: replay the segment on a second stream, beside what the caller enqueues next; its results are valid on the
        caller's stream after join_side().  Autograd runs the segment's backward on that stream too -- beside the backward of
        whatever the caller ran in between -- and orders it against producers and consumers of its gradients itself."""
        if name in self.skip:
            return fn(*args)
        return _call_segment(self, name, owner, fn, args, side and self.side_segments)

    def join_side(self):
        """Make the results of every `side=True` segment of this step visible to the current stream."""
        if self._side_pending:
            torch.cuda.current_stream().wait_stream(self._side_stream)
            self._side_pending = False

    def run(self, name, make_module, args, key_extra=(), want_module=False, owner=None, side=False):
        """Run segment `name` on tensor arguments `args` through its graph (capturing it on first use).  `owner`
        (optional) is the module the segment belongs to: two models that run the same segment on the same shapes (student
        and EMA teacher) get separate graphs."""
        idx = self.calls.get(name, 0)
        self.calls[name] = idx + 1
        grad = torch.is_grad_enabled()
        key = (name, idx, grad, id(owner) if owner is not None else None, key_extra) + \
            tuple((tuple(a.shape), a.dtype, a.requires_grad) for a in args)
        entry = self.cache.get(key)
        if entry is None:
            # a segment whose signature keeps changing (e.g. the criterion with real data: its index tensors are sized by
            # the number of boxes in the batch) stays eager: capturing costs far more than one eager pass and every
            # capture keeps its private memory pool
            if self.per_segment.get((name, idx), 0) >= self.max_signatures or (not grad and not self.capture_inference):
                module = make_module()
                out = module(*args)
                return (out, module) if want_module else out
            self.per_segment[(name, idx)] = self.per_segment.get((name, idx), 0) + 1
            module = make_module()
            n0 = _native_launches()
            if grad:
                # static inputs: private clones -- except for arguments that live in the static output buffers of a segment
                # captured earlier (same address at every replay): those are captured in place, so a replay finds
                # `static_input.data_ptr() == arg.data_ptr()` and skips the device copy + its launch (the criterion segment
                # alone has ~50 such inputs; DATR_GRAPH_ALIAS=0 restores the clones)
                sample = tuple((a.detach() if self._stable(a) else a.detach().clone()).requires_grad_(a.requires_grad)
                               for a in args)
                graphed = _TrainingGraph(module, sample, self.warmup_iters, self.grad_sinks if self.sink_grads else {})
                # warm-up iterations + one capture each ran forward and backward eagerly/under capture once
                per_pair = (_native_launches() - n0) // (self.warmup_iters + 1)
            else:
                graphed = _InferenceGraph(module, args, self.warmup_iters)
                per_pair = (_native_launches() - n0) // (self.warmup_iters + 1)
            entry = self.cache[key] = (graphed, per_pair)
            self.captures += 1
        graphed, per_pair = entry
        self.replayed_native_launches += per_pair
        if side and grad:
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream()
            self._side_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side_stream):
                out = graphed(*args)
            self._side_pending = True
        else:
            out = graphed(*args)
        if self.alias_inputs and grad:
            for o in (out if isinstance(out, (tuple, list)) else (out,)):
                if isinstance(o, torch.Tensor) and o.is_cuda:
                    self.stable_storages.add(o.untyped_storage().data_ptr())
        return (out, graphed) if want_module else out

    def _stable(self, a):
        """True if `a` aliases the static output storage of a segment captured earlier in this process."""
        return (self.alias_inputs and isinstance(a, torch.Tensor) and a.is_cuda
                and a.untyped_storage().data_ptr() in self.stable_storages)


class _TrainingGraph:
    """Forward and backward CUDA graphs of one module -- the capture protocol of torch.cuda.make_graphed_callables (side-
    stream warm-up, forward capture, backward capture of torch.autograd.grad into static gradient tensors sharing one
    memory pool, replay from a torch.autograd.Function) with one difference: parameters listed in `sinks`
    (id(parameter) -> persistent gradient buffer) get their gradient ADDED INTO that buffer inside the captured backward --
    by the weight-gradient kernel itself where a Linear owns the parameter (linear.GradSinks), by multi-tensor launches at
    the end of the capture for the rest -- and the autograd node returns None for them.  torch's version hands the static gradient of
    every parameter to AccumulateGrad, i.e. one eager `grad += new` kernel per parameter per replay (640 per DINO step)."""

    def __init__(self, module, sample_args, warmup_iters, sinks):
        from torch.utils import _pytree as pytree
        self.module = module
        self.training = module.training
        n_args = len(sample_args)
        params = tuple(module.parameters())
        surface = tuple(sample_args) + params
        req = tuple(t for t in surface if t.requires_grad)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup_iters):
                outs = pytree.tree_leaves(module(*sample_args))
                outs = tuple(o for o in outs if isinstance(o, torch.Tensor) and o.requires_grad)
                if outs and req:
                    torch.autograd.grad(outs, req, tuple(torch.empty_like(o) for o in outs), only_inputs=True,
                                        allow_unused=True)
                del outs
        cur.wait_stream(side)
        pool = torch.cuda.graph_pool_handle()
        fwd, bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(fwd, pool=pool):
            out = module(*sample_args)
        static_outputs, out_spec = pytree.tree_flatten(out)
        static_outputs = tuple(static_outputs)
        static_grad_outputs = tuple(torch.empty_like(o) if o.requires_grad else None for o in static_outputs)
        outs_req = tuple(o for o in static_outputs if o.requires_grad)
        n_sunk = 0
        grads = ()
        probe = None
        self.side_launches = 0
        if outs_req and req:
            from . import linear as dl
            branch = dl.SideWgrad(torch.cuda.Stream(), dl.SIDE_WGRAD_MAX_ROWS) if dl.SIDE_WGRAD_MAX_ROWS > 0 else None
            with torch.cuda.graph(bwd, pool=pool):
                dl._SIDE = branch        # small weight-gradient launches become a parallel branch of this graph
                direct = dl._SINKS = dl.GradSinks(sinks) if sinks else None     # Linears reduce dW / db straight into the sinks
                try:
                    grads = torch.autograd.grad(outs_req, req, tuple(g for g in static_grad_outputs if g is not None),
                                                only_inputs=True, allow_unused=True)
                finally:
                    dl._SIDE = dl._SINKS = None
                if branch is not None and branch.used:
                    torch.cuda.current_stream().wait_stream(branch.stream)
                    self.side_launches = branch.used
                dst, src = [], []
                kept = []
                for t, g in zip(req, grads):
                    sink = sinks.get(id(t)) if g is not None else None
                    if sink is not None and sink.shape == g.shape and sink.dtype == g.dtype:
                        if not dst:
                            probe = (t, sink)
                        dst.append(sink)
                        src.append(g)
                        kept.append(None)
                    else:
                        kept.append(g)
                if dst:
                    torch._foreach_add_(dst, src)
                n_sunk = len(dst)
                if direct is not None and direct.used:
                    n_sunk += len({id(p) for p, _ in direct.used})
                    if probe is None:
                        probe = direct.used[0]
                grads = tuple(kept)
                del dst, src, kept
        it = iter(grads)
        static_grad_inputs = tuple(next(it) if t.requires_grad else None for t in surface) if grads else (None,) * len(surface)
        self.n_sunk = n_sunk
        self._keep = (fwd, bwd, static_outputs, static_grad_outputs, static_grad_inputs, surface)
        has_bwd = bool(outs_req and req)

        class Graphed(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *inputs):
                for i in range(n_args):
                    if surface[i].data_ptr() != inputs[i].data_ptr():
                        surface[i].copy_(inputs[i])
                fwd.replay()
                return tuple(o.detach() for o in static_outputs)

            @staticmethod
            @torch.autograd.function.once_differentiable
            def backward(ctx, *gouts):
                if not has_bwd:
                    return (None,) * len(surface)
                for s, g in zip(static_grad_outputs, gouts):
                    if s is not None and s.data_ptr() != g.data_ptr():
                        s.copy_(g)
                if probe is not None and (probe[0].grad is None or probe[0].grad.data_ptr() != probe[1].data_ptr()):
                    raise RuntimeError("datr_b200.graphs: this segment's backward adds its parameter gradients into the "
                                       "FlatGradients buffer registered at capture, but .grad no longer points there "
                                       "(zero_grad(set_to_none=True) or `p.grad = None`?); keep the views in place "
                                       "(FlatGradients.zero()) or set DATR_GRAPH_SINKS=0")
                bwd.replay()
                return tuple(g.detach() if g is not None else None for g in static_grad_inputs)

        self._fn, self._params, self._out_spec = Graphed, params, out_spec

    def __getattr__(self, name):            # FnSegment bookkeeping (out_spec, out_consts, out_n) lives on the module
        return getattr(self.__dict__["module"], name)

    def __call__(self, *args):
        from torch.utils import _pytree as pytree
        if self.module.training != self.training:
            return self.module(*args)
        return pytree.tree_unflatten(list(self._fn.apply(*(tuple(args) + self._params))), self._out_spec)


class _InferenceGraph:
    """Forward-only CUDA graph of a segment (no-grad passes: the EMA teacher of the self-training step, evaluation).
    Inputs are copied into static buffers, the captured forward is replayed, and the STATIC output tensors are returned:
    they are valid until the next replay of the same segment, i.e. until the same point of the next step."""

    def __init__(self, module, args, warmup_iters):
        self.module = module
        self.static_in = tuple(a.detach().clone() for a in args)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup_iters):
                module(*self.static_in)
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = module(*self.static_in)

    def __getattr__(self, name):            # FnSegment bookkeeping (out_spec, out_consts, out_n) lives on the module
        return getattr(self.__dict__["module"], name)

    def __call__(self, *args):
        for s, a in zip(self.static_in, args):
            if s.data_ptr() != a.data_ptr():
                s.copy_(a)
        self.graph.replay()
        return self.static_out


class FnSegment(nn.Module):
    """A pure function of a pytree of tensors (plus constants) run as a graph segment.  `owner` supplies the
    parameters the function may touch (their gradients come back from the graphed backward); non-tensor leaves of the
    arguments are baked into the capture and are part of the cache key."""

    def __init__(self, owner, fn, spec, consts, n_leaves):
        super().__init__()
        self.owner = owner                      # registers the owner's parameters with this segment
        self._fn, self._spec, self._consts, self._n = fn, spec, dict(consts), n_leaves
        self.out_spec, self.out_consts = None, None
        if owner is not None:
            self.train(owner.training)

    def forward(self, *tensors):
        from torch.utils import _pytree as pytree
        it = iter(tensors)
        leaves = [self._consts[i] if i in self._consts else next(it) for i in range(self._n)]
        out = self._fn(*pytree.tree_unflatten(leaves, self._spec))
        flat, spec = pytree.tree_flatten(out)
        self.out_spec = spec
        self.out_consts = {i: l for i, l in enumerate(flat) if not isinstance(l, torch.Tensor)}
        self.out_n = len(flat)
        return tuple(l for l in flat if isinstance(l, torch.Tensor))


def _call_segment(sg: "StepGraphs", name, owner, fn, args, side=False):
    from torch.utils import _pytree as pytree
    leaves, spec = pytree.tree_flatten(args)
    consts = tuple((i, l) for i, l in enumerate(leaves) if not isinstance(l, torch.Tensor))
    tensors = tuple(l for l in leaves if isinstance(l, torch.Tensor))
    holder = {}

    def make():
        holder["m"] = FnSegment(owner, fn, spec, consts, len(leaves))
        return holder["m"]

    key_extra = (repr(spec), repr(consts))
    out_tensors, module = sg.run(name, make, tensors, key_extra=key_extra, want_module=True, owner=owner, side=side)
    if not isinstance(out_tensors, tuple):
        out_tensors = (out_tensors,)
    it = iter(out_tensors)
    flat = [module.out_consts[i] if i in module.out_consts else next(it) for i in range(module.out_n)]
    return pytree.tree_unflatten(flat, module.out_spec)


class BodySegment(nn.Module):
    """ResNet body: images [B,3,H,W] -> tuple of the requested stage outputs."""

    def __init__(self, body):
        super().__init__()
        self.body = body
        self.train(body.training)

    def forward(self, x):
        return tuple(self.body(x).values())


class EncoderSegment(nn.Module):
    def __init__(self, encoder, shapes_list):
        super().__init__()
        self.encoder, self.shapes_list = encoder, list(shapes_list)
        self.train(encoder.training)

    def forward(self, src, pos, spatial_shapes, level_start_index, valid_ratios, key_padding_mask):
        return self.encoder(src, pos=pos, spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                            valid_ratios=valid_ratios, key_padding_mask=key_padding_mask,
                            shapes_list=self.shapes_list)[0]


class DecoderSegment(nn.Module):
    """Decoder stack: returns the per-layer outputs followed by the reference boxes as one flat tuple."""

    def __init__(self, decoder, has_mask):
        super().__init__()
        self.decoder, self.has_mask = decoder, has_mask
        self.train(decoder.training)

    def forward(self, tgt, memory, memory_key_padding_mask, pos, refpoints_unsigmoid, level_start_index,
                spatial_shapes, valid_ratios, *tgt_mask):
        hs, refs = self.decoder(tgt=tgt, memory=memory, memory_key_padding_mask=memory_key_padding_mask, pos=pos,
                                refpoints_unsigmoid=refpoints_unsigmoid, level_start_index=level_start_index,
                                spatial_shapes=spatial_shapes, valid_ratios=valid_ratios,
                                tgt_mask=tgt_mask[0] if self.has_mask else None)
        return tuple(hs) + tuple(refs)
