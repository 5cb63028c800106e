"""Hungarian assignment between predictions and ground-truth boxes.

Mirrors HungarianMatcher / SimpleMinsumMatcher / build_matcher of the reference's
models/dino/matcher.py (:23-95, :98-169, :172-190): cost = cost_class * focal-style class cost
+ cost_bbox * L1 + cost_giou * (-GIoU), solved per image by scipy's linear_sum_assignment on the host
(the assignment must stay bit-exact, so the solver is the same library call as the reference's).
One D2H copy per call, on the current stream, through pinned memory.
"""
import torch
from scipy.optimize import linear_sum_assignment
from torch import nn

from datr_b200.util.box_ops import box_cxcywh_to_xyxy, boxes_well_formed, generalized_box_iou
from datr_b200.util.misc import upload


# Solve the assignments of a step on the GPU (csrc/lsa.cu) instead of scipy on a host copy of the cost matrix: same
# algorithm, identical assignments, no host synchronisation.  DATR_MATCHER=host keeps scipy.
import os as _os
DEVICE_SOLVER = _os.environ.get("DATR_MATCHER", "device") != "host"


def _cost_matrix(outputs, targets, w_class, w_bbox, w_giou, alpha, gamma=2.0, deferred_check=None):
    """deferred_check: a list that receives the device-side result of the degenerate-box asserts instead of
    synchronising on them here (BatchedMatch checks it on the host together with the cost matrix)."""
    bs, nq = outputs["pred_logits"].shape[:2]
    prob = outputs["pred_logits"].flatten(0, 1).sigmoid()
    boxes = outputs["pred_boxes"].flatten(0, 1)
    tgt_ids = torch.cat([t["labels"] for t in targets])
    tgt_box = torch.cat([t["boxes"] for t in targets])
    neg = (1 - alpha) * (prob ** gamma) * (-(1 - prob + 1e-8).log())
    pos = alpha * ((1 - prob) ** gamma) * (-(prob + 1e-8).log())
    c_class = pos[:, tgt_ids] - neg[:, tgt_ids]
    c_bbox = torch.cdist(boxes, tgt_box, p=1)
    b1, b2 = box_cxcywh_to_xyxy(boxes), box_cxcywh_to_xyxy(tgt_box)
    if deferred_check is not None:
        deferred_check.append(boxes_well_formed(b1, b2))
    c_giou = -generalized_box_iou(b1, b2, check=deferred_check is None)
    C = w_bbox * c_bbox + w_class * c_class + w_giou * c_giou
    return C.view(bs, nq, -1), [len(t["boxes"]) for t in targets]


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal_alpha=0.25):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha = cost_class, cost_bbox, cost_giou, focal_alpha

    @torch.no_grad()
    def forward(self, outputs, targets):
        """-> [(pred_idx int64[K_i], tgt_idx int64[K_i])] per image, K_i = min(num_queries, num_targets_i)."""
        C, sizes = _cost_matrix(outputs, targets, self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha)
        C = C.cpu()
        pairs = [linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))]
        return [(torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)) for i, j in pairs]


class BatchedMatch:
    """The assignments of several prediction sets against the same targets, in two phases so that the host-side part
    (scipy) can overlap GPU work enqueued in between:

      begin (constructor): ONE cost-matrix pass over the stacked sets and ONE asynchronous device->host copy into
          pinned memory, closed by an event; with torch.distributed initialised the num_boxes all-reduce of
          SetCriterion.forward (reference dino.py:767-770) rides along.  Nothing blocks.
      result(): waits for the event only (not for work enqueued after it), runs linear_sum_assignment per image and
          returns (one index list per prediction set, num_boxes) -- the index vectors go back in one pinned,
          asynchronous copy.

    Stacking along the batch axis leaves every row of the cost matrix -- and so every assignment -- bit-identical to
    matching the sets one by one (the reference does 7 matchings with 7 syncs per step, dino.py:723-933 /
    matcher.py:91)."""

    _problem_tables = {}      # (device, n_sets, bs, nq, sizes) -> (int64 [P, 5] device table, max boxes, output length)
    _pending_checks = []      # (pinned flag, event) of earlier steps' degenerate-box checks (device mode: read one step late)

    def __init__(self, matcher, outputs_list, targets):
        self.targets, self.n_sets = targets, len(outputs_list)
        self.flat_dev = None
        self.logits = [o["pred_logits"] for o in outputs_list]
        self.bs = self.logits[0].shape[0]
        stacked = {"pred_logits": torch.cat(self.logits, 0), "pred_boxes": torch.cat([o["pred_boxes"] for o in outputs_list], 0)}
        self.device = dev = stacked["pred_logits"].device
        n_boxes = sum(len(t["labels"]) for t in targets)
        self.world, self.nb_host = 1, None
        with torch.no_grad():
            ok = []
            C, self.sizes = _cost_matrix(stacked, targets, matcher.cost_class, matcher.cost_bbox, matcher.cost_giou,
                                         matcher.focal_alpha, deferred_check=ok)
            nb = None
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                self.world = torch.distributed.get_world_size()
                nb = upload([n_boxes], dtype=torch.float, device=dev)
                torch.distributed.all_reduce(nb)
            if dev.type == "cuda" and DEVICE_SOLVER and self._solve_on_device(C, ok[0], nb, n_boxes):
                pass
            elif dev.type == "cuda":
                self.C = torch.empty(C.shape, dtype=C.dtype, pin_memory=True)
                self.C.copy_(C, non_blocking=True)
                self.ok = torch.empty(1, dtype=torch.bool, pin_memory=True)
                self.ok.copy_(ok[0].reshape(1), non_blocking=True)
                if nb is not None:
                    self.nb_host = torch.empty(1, dtype=torch.float, pin_memory=True)
                    self.nb_host.copy_(nb, non_blocking=True)
                self.event = torch.cuda.Event()
                self.event.record(torch.cuda.current_stream(dev))
            else:
                self.C, self.event, self.ok = C, None, ok[0].reshape(1)
                self.nb_host = nb
        self.n_boxes = float(n_boxes)
        self.consumed = False

    def _solve_on_device(self, C, ok, nb, n_boxes) -> bool:
        """All assignments of the step in ONE kernel launch on the cost matrix where it is (csrc/lsa.cu: scipy's algorithm,
        identical assignments), no device->host copy, no synchronisation: result() returns device index tensors at once.
        The degenerate-box assert of util/box_ops.py:48-49 is read back one step late."""
        from datr_b200 import native
        bs, nq = self.bs, C.shape[1]
        sizes = tuple(int(n) for n in self.sizes)
        if not sizes or max(sizes) > nq or max(sizes) > 2048 or sum(sizes) == 0:
            return False
        dev = self.device
        key = (str(dev), self.n_sets, bs, nq, sizes)
        hit = BatchedMatch._problem_tables.get(key)
        if hit is None:
            if torch.cuda.is_current_stream_capturing():
                return False
            T = sum(sizes)
            rows, off = [], 0
            for g in range(self.n_sets):
                col = 0
                for i, n in enumerate(sizes):
                    rows.append([((g * bs + i) * nq) * T + col, T, nq, n, off])
                    col += n
                    off += 2 * n
            if len(BatchedMatch._problem_tables) > 64:
                BatchedMatch._problem_tables.clear()
            hit = BatchedMatch._problem_tables[key] = (upload(rows, dtype=torch.int64, device=dev), max(sizes), off)
        table, max_nt, total = hit
        C = C.contiguous()
        flat = torch.empty(total, dtype=torch.int64, device=dev)
        lib = native.lib()
        with torch.cuda.device(dev):
            rc = lib.datr_lsa_solve(C.data_ptr(), table.data_ptr(), table.shape[0], nq, max_nt, flat.data_ptr(),
                                    torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"datr_lsa_solve failed (code {rc}): {lib.datr_lsa_last_error().decode()}")
        self.flat_dev, self.C, self.event = flat, None, None
        # deferred degenerate-box check: flags of earlier steps whose copy has completed are examined now
        still = []
        for flag, ev in BatchedMatch._pending_checks:
            if ev.query():
                assert bool(flag[0]), "degenerate boxes (x1 < x0 or y1 < y0) reached the matcher"   # util/box_ops.py:48-49
            else:
                still.append((flag, ev))
        flag = torch.empty(1, dtype=torch.bool, pin_memory=True)
        flag.copy_(ok.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        BatchedMatch._pending_checks = still[-8:] + [(flag, ev)]
        self.ok = None
        # num_boxes: host-known on one rank; after the all-reduce it stays on the device (a 0-dim tensor)
        self.nb_dev = torch.clamp(nb[0] / self.world, min=1.0) if nb is not None else None
        return True

    def matches(self, outputs_list, targets) -> bool:
        """True if this prefetched match was started for exactly these prediction tensors and targets (and has not
        been used yet: graph-replayed outputs are the same tensor objects every step)."""
        return (not self.consumed and targets is self.targets and len(outputs_list) == self.n_sets
                and all(o["pred_logits"] is l for o, l in zip(outputs_list, self.logits)))

    def result(self):
        self.consumed = True
        if self.flat_dev is not None:          # solved on the device: nothing to wait for
            out, off = [], 0
            for _ in range(self.n_sets):
                cur = []
                for n in self.sizes:
                    cur.append((self.flat_dev[off:off + n], self.flat_dev[off + n:off + 2 * n]))
                    off += 2 * n
                out.append(cur)
            return out, (self.nb_dev if self.nb_dev is not None else max(self.n_boxes / self.world, 1.0))
        if self.event is not None:
            self.event.synchronize()
        assert bool(self.ok[0]), "degenerate boxes (x1 < x0 or y1 < y0) reached the matcher"   # util/box_ops.py:48-49
        C, bs = self.C, self.bs
        pairs = [[linear_sum_assignment(c[g * bs + i]) for i, c in enumerate(C.split(self.sizes, -1))]
                 for g in range(self.n_sets)]
        # all index vectors travel to the device in ONE pinned, asynchronous copy (the losses index device tensors with
        # them; CPU index tensors would cost a pageable host->device copy per use, ~40 per step)
        import numpy as np
        flat = np.concatenate([np.asarray(v, dtype=np.int64) for grp in pairs for ij in grp for v in ij]) \
            if any(len(ij[0]) for grp in pairs for ij in grp) else np.zeros(0, dtype=np.int64)
        host = torch.from_numpy(flat)
        if self.device.type == "cuda":
            host = host.pin_memory()
        flat_dev = host.to(self.device, non_blocking=True)
        out, off = [], 0
        for grp in pairs:
            cur = []
            for i, j in grp:
                n = len(i)
                cur.append((flat_dev[off:off + n], flat_dev[off + n:off + 2 * n]))
                off += 2 * n
            out.append(cur)
        total = float(self.nb_host[0]) if self.nb_host is not None else self.n_boxes
        return out, max(total / self.world, 1.0)


def batchable(matcher, outputs_list) -> bool:
    return (isinstance(matcher, HungarianMatcher) and len(outputs_list) > 1
            and all(o["pred_logits"].shape == outputs_list[0]["pred_logits"].shape for o in outputs_list))


def matching_sets(outputs, key_aux="aux_outputs", key_interm="interm_outputs"):
    """The prediction sets SetCriterion matches in one step: final output, auxiliary decoder layers, intermediate."""
    head = {k: v for k, v in outputs.items() if k != key_aux}
    return [head] + list(outputs.get(key_aux, [])) + ([outputs[key_interm]] if key_interm in outputs else [])


def prefetch(matcher, outputs, targets):
    """Called by DINO.forward right after the source-domain heads: starts the step's matchings so that their
    device->host copy and (later, in SetCriterion.forward) the scipy solve overlap the target-domain transformer pass
    instead of idling the GPU.  The handle travels on the prediction tensor; SetCriterion picks it up."""
    if matcher is None or not targets or not outputs["pred_logits"].is_cuda:
        return
    sets = matching_sets(outputs)
    if batchable(matcher, sets):
        outputs["pred_logits"]._datr_match = BatchedMatch(matcher, sets, targets)


def take_prefetched(outputs, sets, targets):
    handle = getattr(outputs["pred_logits"], "_datr_match", None)
    if handle is None:
        return None
    del outputs["pred_logits"]._datr_match
    return handle if handle.matches(sets, targets) else None


def match_many(matcher, outputs_list, targets):
    """Assignments for several prediction sets against the same targets (see BatchedMatch); one index list per set."""
    if not batchable(matcher, outputs_list):
        return [matcher(o, targets) for o in outputs_list]
    return BatchedMatch(matcher, outputs_list, targets).result()[0]


class SimpleMinsumMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal_alpha=0.25):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha = cost_class, cost_bbox, cost_giou, focal_alpha

    @torch.no_grad()
    def forward(self, outputs, targets):
        C, sizes = _cost_matrix(outputs, targets, self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha)
        out = []
        for i, (c, n) in enumerate(zip(C.split(sizes, -1), sizes)):
            out.append((c[i].min(0)[1].to(torch.int64), torch.arange(n, device=C.device, dtype=torch.int64)))
        return out


def build_matcher(args):
    kinds = {"HungarianMatcher": HungarianMatcher, "SimpleMinsumMatcher": SimpleMinsumMatcher}
    assert args.matcher_type in kinds, f"Unknown args.matcher_type: {args.matcher_type}"
    return kinds[args.matcher_type](cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox,
                                    cost_giou=args.set_cost_giou, focal_alpha=args.focal_alpha)
