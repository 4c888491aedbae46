"""Diagnostic: where Mapper.process_frame spends its time (torch.profiler over 3 frames of the synthetic KITTI loop)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pin_slam_b200.frame_loop import FrameLoop
from torch.profiler import profile, ProfilerActivity

loop = FrameLoop(device="cuda:0")
loop.step(0, timed=False, map_iters=100)
for f in range(1, 6):
    loop.step(f)
torch.cuda.synchronize()
mp = loop.mapper
orig = mp.process_frame
stage = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); stage[name] = stage.get(name, 0) + (time.perf_counter() - t0) * 1e3
        return r
    return w
mp.sampler.sample = timed("sampler.sample", mp.sampler.sample)
npm = loop.neural_points
npm.update = timed("npm.update (incl. reset_local_map)", npm.update)
npm.reset_local_map = timed("  reset_local_map", npm.reset_local_map)
npm.query_certainty = timed("query_certainty", npm.query_certainty)
mp.process_frame = timed("process_frame total", orig)
n = 5
for f in range(6, 6 + n):
    loop.step(f)
print({k: round(v / n, 3) for k, v in stage.items()})
mp.process_frame = orig
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for f in range(6 + n, 6 + n + 2):
        loop.step(f)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=50))
