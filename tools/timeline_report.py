"""Chronological device timeline of one graph-replayed DINO DA training step: the kernels of every CUDA-graph launch
(grouped by the correlation id of its cudaGraphLaunch) and the eager kernels between two launches, in order, with span,
busy time, kernel count, short-kernel count and the heaviest kernels of the group.  GPU box only.
Writes gpurun_out/dino_step_timeline.txt.  Arguments: DinoStep keyword overrides as key=value."""
import json, os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino
from torch.profiler import profile, ProfilerActivity

over = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    over[k] = eval(v)
wl = bench_dino.DinoStep(torch.device("cuda", 0), **over)
for _ in range(4):
    wl.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    wl.step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
trace = json.load(open(path))["traceEvents"]
kern = [e for e in trace if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
kern.sort(key=lambda e: e["ts"])
launch = {e["args"].get("correlation"): e["name"] for e in trace
          if e.get("cat") == "cuda_runtime" and "args" in e and "correlation" in e["args"]}
t0 = kern[0]["ts"]
groups = []          # [kind, [events]]
for e in kern:
    corr = e["args"].get("correlation")
    is_graph = "GraphLaunch" in launch.get(corr, "")
    key = ("graph", corr) if is_graph else ("eager", None)
    if groups and groups[-1][0] == key:
        groups[-1][1].append(e)
    else:
        groups.append([key, [e]])
out = os.path.join(ROOT, "gpurun_out", "dino_step_timeline.txt")
with open(out, "w") as f:
    f.write("start ms | span ms | busy ms | kernels | <6us | kind | heaviest kernels (ms, count)\n")
    for key, evs in groups:
        s = evs[0]["ts"]; t = max(e["ts"] + e["dur"] for e in evs)
        busy = sum(e["dur"] for e in evs)
        short = sum(1 for e in evs if e["dur"] < 6)
        names = {}
        for e in evs:
            c = names.setdefault(e["name"].replace("void ", "").replace("(anonymous namespace)::", "")
                                 .replace("at::native::", "")[:48], [0.0, 0])
            c[0] += e["dur"]; c[1] += 1
        top = sorted(names.items(), key=lambda x: -x[1][0])[:4]
        f.write(f"{(s - t0) / 1e3:7.2f} | {(t - s) / 1e3:7.2f} | {busy / 1e3:7.2f} | {len(evs):5d} | {short:5d} | {key[0]:5s} | "
                + "; ".join(f"{n} {c[0] / 1e3:.2f}/{c[1]}" for n, c in top) + "\n")
# ordered kernel list of every group with more than 100 kernels (what a fused kernel per chain would replace)
detail = os.path.join(ROOT, "gpurun_out", "dino_step_timeline_detail.txt")
with open(detail, "w") as f:
    for key, evs in groups:
        if len(evs) <= 100:
            continue
        f.write(f"== group at {(evs[0]['ts'] - t0) / 1e3:.2f} ms, {len(evs)} kernels\n")
        prev_end = evs[0]["ts"]
        for e in evs:
            name = e["name"].replace("void ", "").replace("(anonymous namespace)::", "").replace("at::native::", "")
            f.write(f"{e['dur']:7.1f} us  gap {max(0.0, e['ts'] - prev_end):5.1f}  {name[:150]}\n")
            prev_end = e["ts"] + e["dur"]
# host side: CPU ops / runtime calls longer than 80 us, in order (where the host makes the GPU wait)
host = os.path.join(ROOT, "gpurun_out", "dino_step_timeline_host.txt")
with open(host, "w") as f:
    cpu = [e for e in trace if e.get("cat") in ("cpu_op", "cuda_runtime", "user_annotation", "python_function") and e.get("dur", 0) > 80]
    cpu.sort(key=lambda e: e["ts"])
    for e in cpu:
        f.write(f"{(e['ts'] - t0) / 1e3:8.2f} ms  {e['dur'] / 1e3:7.2f} ms  {e.get('cat'):14s} {e['name'][:120]}\n")
print(open(out).read()[:12000])
