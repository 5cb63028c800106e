"""Kernel-time breakdown of one DINO DA training step (torch.profiler, CUDA activities); GPU box only.
Writes gpurun_out/dino_step_kernels.txt (top kernels by device time) and prints GPU-busy vs wall time."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
over = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    over[k] = eval(v)
wl = bench_dino.DinoStep(dev, **over)
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    wl.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in ev if e.device_time) / 1e3 if ev else 0
ka = prof.key_averages()
rows = sorted(ka, key=lambda e: -e.self_device_time_total)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "dino_step_kernels.txt"), "w") as f:
    f.write(f"wall ms/step {wall*1e3:.2f}\n")
    tot = sum(e.self_device_time_total for e in rows)
    f.write(f"sum self device time ms {tot/1e3:.2f}\n")
    for e in rows[:70]:
        f.write(f"{e.self_device_time_total/1e3:9.3f} ms {e.count:6d}x  {e.key[:150]}\n")
print(open(os.path.join(ROOT, "gpurun_out", "dino_step_kernels.txt")).read()[:6000])
