"""GPU idle gaps of one DINO DA training step in the benchmark configuration (CUDA graphs on): kernel start / end
timestamps from torch.profiler, sorted, with the largest gaps and the kernels on both sides.  GPU box only.
Writes gpurun_out/dino_step_gaps.txt."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino
from torch.profiler import profile, ProfilerActivity

wl = bench_dino.DinoStep(torch.device("cuda", 0))
for _ in range(4):
    wl.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    wl.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy_until, gaps, busy = ev[0].time_range.start, [], 0.0
prev = ev[0]
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if s > busy_until:
        gaps.append((s - busy_until, busy_until - t0, prev.name, e.name))
        busy += t - s
    elif t > busy_until:
        busy += t - busy_until
    if t > busy_until:
        busy_until, prev = t, e
gaps.sort(key=lambda g: -g[0])
# duration histogram of the kernels, and how the short ones spread over the step (2 ms windows)
bins = [(0, 3), (3, 6), (6, 12), (12, 25), (25, 50), (50, 100), (100, 1e9)]
hist = {b: [0, 0.0] for b in bins}
win = {}
for e in ev:
    d = e.time_range.end - e.time_range.start
    for b in bins:
        if b[0] <= d < b[1]:
            hist[b][0] += 1; hist[b][1] += d
    if d < 6:
        w = int((e.time_range.start - t0) / 2000)
        c = win.setdefault(w, [0, 0.0]); c[0] += 1; c[1] += d
out = os.path.join(ROOT, "gpurun_out", "dino_step_gaps.txt")
with open(out, "w") as f:
    f.write(f"span {(t1 - t0) / 1e3:.2f} ms, GPU busy {busy / 1e3:.2f} ms, idle {sum(g[0] for g in gaps) / 1e3:.2f} ms in {len(gaps)} gaps "
            f"({sum(1 for g in gaps if g[0] > 20)} longer than 20 us = {sum(g[0] for g in gaps if g[0] > 20) / 1e3:.2f} ms)\n")
    for b in bins:
        f.write(f"kernels of {b[0]:>4}-{b[1] if b[1] < 1e9 else 'inf':>5} us: {hist[b][0]:6d}  total {hist[b][1] / 1e3:7.2f} ms\n")
    f.write("kernels shorter than 6 us per 2 ms window (start ms: count, busy ms): " +
            "  ".join(f"{2 * w}: {c[0]}, {c[1] / 1e3:.2f}" for w, c in sorted(win.items())) + "\n")
    names = {}
    for e in ev:
        d = e.time_range.end - e.time_range.start
        if d < 6:
            c = names.setdefault(e.name[:110], [0, 0.0]); c[0] += 1; c[1] += d
    f.write("kernels shorter than 6 us by name:\n")
    for n, c in sorted(names.items(), key=lambda x: -x[1][0])[:45]:
        f.write(f"   {c[0]:5d}x {c[1] / 1e3:6.2f} ms  {n}\n")
    for g in gaps[:40]:
        f.write(f"{g[0]:8.1f} us at +{g[1] / 1e3:7.2f} ms   after {g[2][:60]:60s} before {g[3][:60]}\n")
print(open(out).read()[:5000])
