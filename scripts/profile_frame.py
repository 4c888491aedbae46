"""torch.profiler view of one tracker+mapper frame (run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pin_slam_b200.frame_loop import FrameLoop

loop = FrameLoop(device="cuda")
loop.step(0, timed=False, map_iters=60)
for f in range(1, 4):
    loop.step(f)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for f in range(4, 7):
        loop.step(f)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(loop.times[-3:])
