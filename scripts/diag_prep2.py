"""Where Mapper.process_frame spends its time on the KITTI-shaped loop: stage wall times (one sync per stage boundary)
and the torch.profiler table of process_frame alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pin_slam_b200.frame_loop import FrameLoop

loop = FrameLoop(device="cuda:0")
loop.step(0, timed=False, map_iters=100)
for f in range(1, 8):
    loop.step(f)
torch.cuda.synchronize()
mp, npm = loop.mapper, loop.neural_points
stage = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); stage[name] = stage.get(name, 0) + (time.perf_counter() - t0) * 1e3
        return r
    return w
orig = dict(sample=mp.sampler.sample, update=npm.update, reset=npm.reset_local_map, cert=npm.query_certainty, pf=mp.process_frame)
mp.sampler.sample = timed("sampler.sample", orig["sample"])
npm.update = timed("npm.update (incl. reset_local_map)", orig["update"])
npm.reset_local_map = timed("  reset_local_map", orig["reset"])
npm.query_certainty = timed("query_certainty", orig["cert"])
mp.process_frame = timed("process_frame total", orig["pf"])
n = 6
for f in range(8, 8 + n):
    loop.step(f)
print({k: round(v / n, 3) for k, v in stage.items()})
mp.sampler.sample, npm.update, npm.reset_local_map, npm.query_certainty, mp.process_frame = (orig[k] for k in ("sample", "update", "reset", "cert", "pf"))
# profile process_frame only
calls = []
real_pf = mp.process_frame
def pf_prof(*a, **k):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        r = real_pf(*a, **k)
        torch.cuda.synchronize()
    calls.append(prof)
    return r
mp.process_frame = pf_prof
loop.step(8 + n)
loop.step(9 + n)
prof = calls[-1]
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=44))
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=12, max_name_column_width=60))
