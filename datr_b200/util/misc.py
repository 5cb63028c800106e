"""NestedTensor batching, inverse_sigmoid and distributed probes.

Mirrors the pieces of the reference's util/misc.py that models/dino imports:
NestedTensor (:313-372), nested_tensor_from_tensor_list (:387-409), inverse_sigmoid (:587-591, eps 1e-3),
accuracy (:534-549), get_world_size / is_dist_avail_and_initialized (:440-452).
"""
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


class NestedTensor:
    """A zero-padded image batch plus its padding mask (True on padded pixels)."""

    def __init__(self, tensors: Tensor, mask: Optional[Tensor]):
        self.tensors = tensors
        if isinstance(mask, str):
            if mask != "auto":
                raise ValueError(f"unknown mask spec {mask!r}")
            if tensors.dim() not in (3, 4):
                raise ValueError(f"tensors dim must be 3 or 4 but {tensors.dim()}({tensors.shape})")
            shape = tensors.shape[-2:] if tensors.dim() == 3 else (tensors.shape[0], *tensors.shape[-2:])
            mask = torch.zeros(shape, dtype=torch.bool, device=tensors.device)
        self.mask = mask

    def to(self, device):
        return NestedTensor(self.tensors.to(device), None if self.mask is None else self.mask.to(device))

    def decompose(self):
        return self.tensors, self.mask

    def imgsize(self):
        return [torch.Tensor([(~m).sum(0).max(), (~m).sum(1).max()]) for m in self.mask]

    @property
    def shape(self):
        return {"tensors.shape": self.tensors.shape, "mask.shape": None if self.mask is None else self.mask.shape}

    def __repr__(self):
        return f"NestedTensor({self.shape})"


def nested_tensor_from_tensor_list(tensor_list: List[Tensor]) -> NestedTensor:
    """Pad [C,H_i,W_i] images to the batch maximum (top-left aligned) and build the mask."""
    if isinstance(tensor_list, Tensor):
        tensor_list = list(tensor_list)
    if tensor_list[0].dim() != 3:
        raise ValueError("not supported")
    c = tensor_list[0].shape[0]
    h = max(t.shape[1] for t in tensor_list)
    w = max(t.shape[2] for t in tensor_list)
    batch = tensor_list[0].new_zeros((len(tensor_list), c, h, w))
    mask = torch.ones((len(tensor_list), h, w), dtype=torch.bool, device=batch.device)
    for i, img in enumerate(tensor_list):
        batch[i, :, :img.shape[1], :img.shape[2]].copy_(img)
        mask[i, :img.shape[1], :img.shape[2]] = False
    return NestedTensor(batch, mask)


def inverse_sigmoid(x, eps=1e-3):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


@torch.no_grad()
def accuracy(output, target, topk=(1,)):
    """precision@k in percent (list, one 0-dim tensor per k)."""
    if target.numel() == 0:
        return [torch.zeros([], device=output.device)]
    pred = output.topk(max(topk), 1, True, True)[1].t()
    hit = pred.eq(target.view(1, -1).expand_as(pred))
    return [hit[:k].reshape(-1).float().sum(0) * (100.0 / target.size(0)) for k in topk]


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def upload(data, dtype=None, device=None):
    """Small host data (lists of ints / floats) as a device tensor WITHOUT stalling the host: a pageable host->device copy
    (torch.as_tensor(list, device="cuda")) is ordered behind everything already enqueued on the stream and blocks the
    calling thread until then -- with per-step target counts that is one pipeline drain per call.  Staging through pinned
    memory makes the copy asynchronous (the caching host allocator keeps the staging buffer alive until it has run)."""
    t = torch.as_tensor(data, dtype=dtype)
    device = torch.device(device) if device is not None else t.device
    if device.type != "cuda":
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)
