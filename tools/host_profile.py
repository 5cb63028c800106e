"""cProfile of the host side of the graph-replayed DINO step (10 steps): where Python spends its time between the graph
launches.  GPU box only; writes gpurun_out/dino_step_host_profile.txt."""
import cProfile, io, os, pstats, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino

wl = bench_dino.DinoStep(torch.device("cuda", 0))
for _ in range(5):
    wl.step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    wl.step()
torch.cuda.synchronize()
pr.disable()
out = io.StringIO()
st = pstats.Stats(pr, stream=out).sort_stats("cumulative")
st.print_stats(70)
text = out.getvalue()
open(os.path.join(ROOT, "gpurun_out", "dino_step_host_profile.txt"), "w").write(text)
print(text[:9000])
