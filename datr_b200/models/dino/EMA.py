"""Teacher / best-model exponential moving averages -- host-side mirror of the reference's models/dino/EMA.py
(ModelEMA :21-54, SemiSupModelEMA :56-88, CosineEMA :90-129, is_parallel :7-9, copy_attr :11-17), SURVEY 8 f3.

Same class names, constructor arguments, attributes (`ema`, `updates`, `decay`, ...) and update rules; the per-tensor
Python loop of `update()` (1 280 tiny kernels for DINO-4scale) is ONE multi-tensor CUDA kernel with the reference's
arithmetic (datr_b200.ema.StateDictEMA, include/datr_ema.h).  `main.py:24` / `main_teacher.py:292,374` import this
module as `models.dino.EMA`."""
import math
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from datr_b200.ema import StateDictEMA


def is_parallel(model):
    """True for DataParallel / DistributedDataParallel wrappers."""
    return type(model) in (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)


def copy_attr(a, b, include=(), exclude=()):
    """Copy the public attributes of b to a, optionally restricted to `include` / without `exclude`."""
    for k, v in b.__dict__.items():
        if (len(include) and k not in include) or k.startswith("_") or k in exclude:
            continue
        setattr(a, k, v)


class _EMABase:
    def _make(self, model):
        self.ema = deepcopy(model.module if is_parallel(model) else model).eval()     # fp32 copy
        for p in self.ema.parameters():
            p.requires_grad_(False)
        self._pairing = None

    def _apply(self, model, d):
        """ema_v = ema_v * d + (1 - d) * model_v for every floating-point entry of the state dict."""
        src = model.module if is_parallel(model) else model
        key = id(src)
        if self._pairing is None or self._pairing[0] != key or self._pairing[1].stale():
            msd = src.state_dict()
            esd = self.ema.state_dict()
            self._pairing = (key, StateDictEMA([v for v in esd.values()], [msd[k] for k in esd]))
        self._pairing[1].update(d)

    def update_attr(self, model, include=(), exclude=("process_group", "reducer")):
        copy_attr(self.ema, model, include, exclude)


class ModelEMA(_EMABase):
    """EMA of everything in the model's state dict with the exponential decay ramp d(x) = decay * (1 - exp(-x / 2000))."""

    def __init__(self, model, decay=0.9999, updates=0):
        self._make(model)
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000))

    def update(self, model):
        self.updates += 1
        self._apply(model, self.decay(self.updates))


class SemiSupModelEMA(_EMABase):
    """Constant-decay variant.  The reference's update() calls `self.decay(self.updates)` on the float it stores
    (EMA.py:78) and therefore raises TypeError; the constant decay of its commented-out line (:79) is what is meant
    and what this class applies."""

    def __init__(self, model, decay=0.99, updates=0):
        self._make(model)
        self.updates = updates
        self.decay = decay

    def update(self, model):
        self.updates += 1
        self._apply(model, self.decay(self.updates) if callable(self.decay) else self.decay)


class CosineEMA(_EMABase):
    """EMA whose decay follows a cosine schedule from decay_start to decay_end over total_epoch epochs."""

    def __init__(self, model, decay_start=0.99, decay_end=0.9999, total_epoch=0):
        self._make(model)
        self.total_epoch = total_epoch
        self.decay_start, self.decay_end = decay_start, decay_end
        self.decay = decay_start
        self.updates = 0

    def update(self, model):
        self._apply(model, self.decay)

    def update_decay(self, cur_epoch):
        self.decay = self.decay_end - (self.decay_end - self.decay_start) * (np.cos(np.pi * cur_epoch / self.total_epoch) + 1) / 2
