"""Where the host falls behind the GPU in one DINO DA training step (benchmark configuration, CUDA graphs on):
at each probe the host time of the enqueue and the device time at which the probe's event completes.
lead = device completion - host enqueue (ms): positive = the GPU still had queued work when the host got there."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino
from datr_b200.util.misc import NestedTensor

wl = bench_dino.DinoStep(torch.device("cuda", 0))
for _ in range(4):
    wl.step()
torch.cuda.synchronize()
probes = []


def probe(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    probes.append((name, time.perf_counter(), e))


for it in range(2):
    probes.clear()
    torch.cuda.synchronize()
    probe("start")
    wl.graphs.begin_step()
    wl.grads.zero()
    out = wl.model(NestedTensor(wl.images, wl.mask), wl.targets)
    probe("forward enqueued")
    losses = wl.criterion(out, wl.targets)
    probe("criterion enqueued")
    wd = wl.criterion.weight_dict
    keys = [k for k in losses if k in wd]
    loss = torch.dot(torch.stack([losses[k].reshape(()) for k in keys]), wl._loss_w)
    probe("total loss")
    loss.backward()
    probe("backward enqueued")
    wl.grads.all_reduce(); wl.grads.clip_(wl.args.clip_max_norm); wl.opt.step()
    probe("optimizer enqueued")
    torch.cuda.synchronize()
t0h, e0 = probes[0][1], probes[0][2]
with open(os.path.join(ROOT, "gpurun_out", "host_vs_gpu.txt"), "w") as f:
    for name, th, e in probes:
        tg = e0.elapsed_time(e)
        line = f"{name:22s} host {1e3 * (th - t0h):8.2f} ms   device {tg:8.2f} ms   lead {tg - 1e3 * (th - t0h):8.2f} ms"
        print(line); f.write(line + "\n")
