"""Diagnostic: per-call timing of Mapper.mapping(20) (BASELINE configs[4] setup, 1 GPU) -- looks for run-to-run instability."""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pin_slam_b200 import ops
from pin_slam_b200.config import HotPathConfig
from pin_slam_b200.model import Decoder
from pin_slam_b200.synthetic import build_map, surface_queries
from pin_slam_b200.utils.mapper import Mapper

dev = torch.device("cuda:0")
cfg = HotPathConfig.kitti(device=str(dev), feature_std=0.05, bs_new_sample=0, local_map_radius=1e4)
npm = build_map(cfg, n_surface=2_000_000, seed=0, extent=80.0)
torch.manual_seed(42)
dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
ds = types.SimpleNamespace(processed_frame=0, lose_track=False, stop_status=False, gt_pose_provided=False, odom_poses=None, pgo_poses=None, gt_poses=None)
mapper = Mapper(cfg, ds, npm, {"sdf": dec, "semantic": None, "color": None})
n = 2_000_000
g = torch.Generator().manual_seed(100)
coord = surface_queries(npm, n, seed=200, sigma=0.15)
mapper.global_coord_pool = coord; mapper.coord_pool = coord
mapper.sdf_label_pool = (0.15 * torch.randn(n, generator=g)).to(dev)
mapper.weight_pool = (torch.rand(n, generator=g) * 0.8 + 0.6).to(dev)
mapper.time_pool = torch.zeros(n, dtype=torch.int32, device=dev)
mapper.pool_sample_count = n
torch.manual_seed(1000)
mapper.mapping(5)
torch.cuda.synchronize()
res = []
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); mapper.mapping(20); e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    res.append((round(e0.elapsed_time(e1) / 20, 4), round((t1 - t0) * 1e3 / 20, 4)))
print("ms/iter (gpu events, host issue):", res, flush=True)
