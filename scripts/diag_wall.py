"""Wall-clock cost of the per-frame stages (synchronised before and after each stage; no profiler)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pin_slam_b200.frame_loop import FrameLoop

loop = FrameLoop(device="cuda:0")
mp, npm, trk = loop.mapper, loop.neural_points, loop.tracker
acc = {}
def timed(obj, name, label):
    fn = getattr(obj, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc.setdefault(label, []).append((time.perf_counter() - t0) * 1e3)
        return r
    setattr(obj, name, w)
timed(loop, "preprocess", "preprocess (scan synthesis + 2 voxel filters)")
timed(trk, "track_fixed", "tracker (3 GN)")
timed(mp, "process_frame", "process_frame")
timed(mp, "mapping", "mapping (5 iters)")
timed(mp.sampler, "sample", "  sampler.sample")
timed(npm, "update", "  update (incl. reset_local_map)")
timed(npm, "query_certainty", "  query_certainty")
loop.step(0, timed=False, map_iters=100)
for f in range(1, 16):
    loop.step(f)
med = lambda v: sorted(v)[len(v) // 2]
for k, v in acc.items():
    print(f"{k:50s} median {med(v[3:]):8.3f} ms   max {max(v[3:]):8.3f} ms   (n={len(v[3:])})")
