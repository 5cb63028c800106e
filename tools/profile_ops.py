"""Operator-level breakdown of one EAGER DINO DA training step: device time per (aten op, input shapes), so that the
elementwise / copy / reduction kernels of the step can be attributed to call sites.  GPU box only.
Writes gpurun_out/dino_step_ops.txt."""
import os, sys
os.environ["DATR_GRAPHS"] = "0"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
wl = bench_dino.DinoStep(dev)
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    wl.step()
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = [e for e in ka if e.self_device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CPU]
rows.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
out = os.path.join(ROOT, "gpurun_out", "dino_step_ops.txt")
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "w") as f:
    f.write(f"sum self device time of ops: {tot/1e3:.2f} ms\n")
    for e in rows[:160]:
        f.write(f"{e.self_device_time_total/1e3:8.3f} ms {e.count:5d}x  {e.key[:48]:48s} {str(e.input_shapes)[:170]}\n")
print(open(out).read()[:3000])

# second table: launches per operator name (where the short-kernel tail comes from)
by = {}
for e in rows:
    a = by.setdefault(e.key, [0, 0.0])
    a[0] += e.count; a[1] += e.self_device_time_total
with open(out, "a") as f:
    f.write("\nby operator name, sorted by number of calls with device work:\n")
    for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][0])[:60]:
        f.write(f"{n:6d}x {t/1e3:8.3f} ms  {k[:80]}\n")
print(open(out).read().split("by operator name")[1][:4500])
