"""torch.profiler view of one tracker+mapper frame (run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pin_slam_b200.frame_loop import FrameLoop

loop = FrameLoop(device="cuda")
loop.step(0, timed=False, map_iters=60)
for f in range(1, 10):
    loop.step(f)
torch.cuda.synchronize()
print("device ms (tracker, mapping):", [tuple(round(x, 3) for x in t) for t in loop.times[-5:]])
print("host issue ms (tracker, mapping):", [tuple(round(x, 3) for x in t) for t in loop.host_issue_times[-5:]])


def _ranged(fn, name):
    def wrapped(*a, **k):
        with torch.profiler.record_function(name):
            return fn(*a, **k)
    return wrapped


loop.mapper.mapping = _ranged(loop.mapper.mapping, "mapping")
loop.tracker.track_fixed = _ranged(loop.tracker.track_fixed, "tracker")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for f in range(10, 13):
        loop.step(f)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(loop.times[-3:])
# kernel timeline of the last profiled frame (gpurun_out/frame_timeline.json): [name, start_us, dur_us]
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/frame_trace.json")
import json
tr = json.load(open("gpurun_out/frame_trace.json"))
ev = [(e["name"], e["ts"], e["dur"]) for e in tr["traceEvents"]
      if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda x: x[1])
rng = [(e["name"], e["ts"], e["dur"]) for e in tr["traceEvents"]
       if e.get("ph") == "X" and e.get("cat") in ("user_annotation", "gpu_user_annotation")]
json.dump({"kernels": ev, "ranges": rng}, open("gpurun_out/frame_timeline.json", "w"))
os.remove("gpurun_out/frame_trace.json")
