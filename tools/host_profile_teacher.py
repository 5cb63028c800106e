import cProfile, io, os, pstats, sys
import torch
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from datr_b200 import bench_dino
wl = bench_dino.TeacherStep(torch.device("cuda", 0))
for _ in range(5):
    wl.step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    wl.step()
torch.cuda.synchronize()
pr.disable()
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(45)
print(out.getvalue()[:7000])
