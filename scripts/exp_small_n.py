"""K1 time vs batch size: fused query_kernel (one launch) vs split pipeline (search + wsq decode), weighted_first maps.
python scripts/exp_small_n.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pin_slam_b200 import ops
from pin_slam_b200.config import HotPathConfig
from pin_slam_b200.model import Decoder
from pin_slam_b200.synthetic import build_map, surface_queries


def timed(fn, n=30):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for name, cfg in (("replica F=8 K=6 1x64 + colour", HotPathConfig.replica(device="cuda", feature_std=0.1)),
                  ("cfg2 F=32 K=8 2x64", HotPathConfig.cfg2(device="cuda", feature_std=0.1, local_map_radius=1e4))):
    npm = build_map(cfg, n_surface=400000, seed=3, extent=20.0 if "replica" in name else 60.0)
    torch.manual_seed(1)
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    cdec = Decoder(cfg, cfg.color_mlp_hidden_dim, cfg.color_mlp_level, cfg.color_channel) if cfg.color_on else None
    for n in (2048, 4096, 8192, 16384, 32768, 65536):
        q = surface_queries(npm, n, seed=2)
        out = {}
        row = []
        for split_min in (1 << 30, 1):
            ops.set_option("split_min_queries", split_min)
            ms = timed(lambda: npm.query_sdf(q, dec, need_grad=True, color_decoder=cdec, color_grad=cdec is not None, out=out))
            row.append(ms)
        ops.set_option("split_min_queries", 0)
        print(f"{name}: n {n:6d}  fused {row[0]*1e3:7.1f} us   split+wsq {row[1]*1e3:7.1f} us", flush=True)
