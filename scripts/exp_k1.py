"""K1 timing experiment: python scripts/exp_k1.py [buffer_size ...]  (cfg2 workload of bench.py at other table sizes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pin_slam_b200.config import HotPathConfig
from pin_slam_b200.model import Decoder
from pin_slam_b200.synthetic import build_map, surface_queries

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for bs in [int(float(a)) for a in sys.argv[1:]] or [50_000_000]:
    cfg = HotPathConfig.cfg2(device=str(dev), feature_std=0.1, local_map_radius=1e4)
    cfg.buffer_size = bs
    npm = build_map(cfg, n_surface=3_000_000, seed=0, extent=80.0)
    torch.manual_seed(42)
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    q = surface_queries(npm, 200000, seed=1, sigma=0.1)
    out = {}
    ts = []
    for it in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        npm.query_sdf(q, dec, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    hot = []
    for it in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        npm.query_sdf(q, dec, out=out)
        b.record()
        torch.cuda.synchronize()
        hot.append(a.elapsed_time(b))
    print(f"buffer_size {bs}: map {npm.count()} pts, K1 cold-L2 median {sorted(ts[3:])[2]:.4f} ms, warm-L2 min {min(hot):.4f} ms", flush=True)
    del npm
    torch.cuda.empty_cache()
