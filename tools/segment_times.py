"""Wall time of the segments of one DINO DA training step, each bracketed by a device synchronise (GPU box only).
Also reports the host-only enqueue time of the forward (no sync) to show how launch-bound the step is."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino
from datr_b200.util.misc import NestedTensor

wl = bench_dino.DinoStep(torch.device("cuda", 0))
for _ in range(3):
    wl.step()
torch.cuda.synchronize()


def seg(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return r, (t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3


acc = {}
for it in range(3):
    if wl.graphs is not None:
        wl.graphs.begin_step()
    _, h, w = seg(wl.grads.zero); acc.setdefault("zero", []).append((h, w))
    out, h, w = seg(lambda: wl.model(NestedTensor(wl.images, wl.mask), wl.targets)); acc.setdefault("forward", []).append((h, w))
    losses, h, w = seg(lambda: wl.criterion(out, wl.targets)); acc.setdefault("criterion", []).append((h, w))
    wd = wl.criterion.weight_dict
    loss, h, w = seg(lambda: sum(losses[k] * wd[k] for k in losses if k in wd)); acc.setdefault("weighted_sum", []).append((h, w))
    _, h, w = seg(loss.backward); acc.setdefault("backward", []).append((h, w))
    _, h, w = seg(lambda: (wl.grads.all_reduce(), wl.grads.clip_(0.1), wl.opt.step())); acc.setdefault("clip+adamw", []).append((h, w))
tot_h = tot_w = 0
for k, v in acc.items():
    h = min(x[0] for x in v); w = min(x[1] for x in v); tot_h += h; tot_w += w
    print(f"{k:14s} host-enqueue {h:8.2f} ms   with-sync {w:8.2f} ms")
print(f"{'total':14s} host-enqueue {tot_h:8.2f} ms   with-sync {tot_w:8.2f} ms")
