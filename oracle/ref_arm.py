"""Benchmark baseline that runs the UNMODIFIED reference classes (oracle/_ref, see oracle/make_ref.py).

Test / benchmark infrastructure: `bench.py --impl reference` and the per-frame baseline leg of `bench.py` execute
this file in a process that never imports `pin_slam_b200` and never loads `libpinb200.so`; nothing under
`pin_slam_b200/` imports it.

    python oracle/ref_arm.py query  --device cpu  --steps 3 --warmup 1     # BASELINE configs[1], reference hot call
    python oracle/ref_arm.py frames --device cuda --frames 12             # BASELINE configs[2], reference Tracker/Mapper

Each mode prints one JSON line.  The hot call of the query mode is the reference's own public entry point for this
path, `Tracker.query_source_points` (utils/tracker.py:227-365): `NeuralPoints.query_feature` -> `Decoder.sdf` ->
`get_gradient` in batches of `config.infer_bs`.
"""
import argparse
import importlib.util
import json
import os
import sys
import time
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
ROOT = os.path.dirname(HERE)
N_QUERY = 200_000


def available():
    return os.path.isfile(os.path.join(REF, "model", "neural_points.py"))


def load_reference():
    """Import the reference modules from oracle/_ref (GUI / IO dependencies that the hot path never touches are
    stubbed)."""
    for m in ["open3d", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "roma", "wandb", "natsort", "skimage",
              "skimage.measure", "pypose", "gtsam", "dtyper", "pyquaternion", "laspy", "evo", "dataset",
              "dataset.slam_dataset"]:
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import torch  # noqa: F401
    from model.decoder import Decoder
    from model.neural_points import NeuralPoints
    from utils.config import Config
    from utils.mapper import Mapper
    from utils.tools import voxel_down_sample_torch
    from utils.tracker import Tracker

    return types.SimpleNamespace(Decoder=Decoder, NeuralPoints=NeuralPoints, Config=Config, Mapper=Mapper,
                                 Tracker=Tracker, voxel_down_sample_torch=voxel_down_sample_torch)


def load_synthetic():
    """pin_slam_b200/synthetic.py holds the seeded scene / scan generators (pure torch).  It is loaded BY FILE so that
    the package (and with it the CUDA extension) is never imported in this process."""
    spec = importlib.util.spec_from_file_location("_pinb_synth", os.path.join(ROOT, "pin_slam_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_config(R, kind, device):
    import torch

    cfg = R.Config()
    if kind == "kitti":
        cfg.load(os.path.join(REF, "config", "lidar_slam", "run_kitti.yaml"))
    elif kind == "replica":
        cfg.load(os.path.join(REF, "config", "rgbd_slam", "run_replica.yaml"))
    else:  # BASELINE configs[1]: F = 32, K = 8, 2 x 64 decoder (reachable through YAML, utils/config.py:394-414)
        cfg.feature_dim = 32
        cfg.query_nn_k = 8
        cfg.geo_mlp_level = 2
        cfg.voxel_size_m = 0.4
        cfg.feature_std = 0.1
        cfg.local_map_radius = 1e4
        cfg.track_on = True
    cfg.device = device
    cfg.pgo_on = False
    cfg.silence = True
    cfg.setup_dtype()
    torch.set_default_dtype(cfg.dtype)
    return cfg


def build_cfg2(R, S, device):
    """The cfg2 workload of bench.py built with the reference's own classes: the same seeded surface points grown into
    a map through `NeuralPoints.update`, the same seeded queries."""
    import torch

    cfg = make_config(R, "cfg2", device)
    dev = torch.device(device)
    npm = R.NeuralPoints(cfg)
    npm.travel_dist = torch.zeros(4, device=dev)
    pts = S.room_surface_points(3_000_000, 0, extent=80.0, device=dev)
    torch.manual_seed(1234)
    npm.update(pts, torch.tensor([0.0, 0.0, 1.0], device=dev), torch.eye(3, device=dev), 0)
    torch.manual_seed(42)
    dec = R.Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1).to(dev)
    for prm in dec.parameters():
        prm.requires_grad_(False)
    q = S.surface_queries(npm, N_QUERY, seed=1, sigma=0.1)
    tracker = R.Tracker(cfg, npm, {"sdf": dec, "semantic": None, "color": None})
    return cfg, npm, dec, q, tracker


def hot_call(tracker, cfg, q):
    """The reference's hot call: SDF, gradient, mask, certainty, std for every query."""
    return tracker.query_source_points(q, cfg.infer_bs, True, True, False, False, query_locally=True,
                                       mask_min_nn_count=cfg.track_mask_query_nn_k)


def workload_stats(npm, q):
    import torch

    with torch.no_grad():
        _, idx = npm.radius_neighborhood_search(q[:50000])
        n_occ = float((idx >= 0).sum(1).float().mean())
        _, idx_l = npm.radius_neighborhood_search(q[:50000], time_filtering=False)
        k_v = float(torch.clamp((npm.global2local[idx_l] >= 0).sum(1), max=npm.config.query_nn_k).float().mean())
    return n_occ, k_v


def run_query(args):
    import torch

    R, S = load_reference(), load_synthetic()
    dev = args.device
    cores = os.cpu_count() or 1
    cfg, npm, dec, q, tracker = build_cfg2(R, S, dev)
    n_occ, k_v = workload_stats(npm, q)
    threads_tried = {}
    if dev == "cpu":
        # the op mix is ~200 small ATen calls per batch: more threads than a few dozen only add fork/join latency
        probe = q[:20000]
        for t in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
            torch.set_num_threads(t)
            hot_call(tracker, cfg, probe[:2000])
            t0 = time.perf_counter()
            hot_call(tracker, cfg, probe)
            threads_tried[t] = time.perf_counter() - t0
        best = min(threads_tried, key=threads_tried.get)
        torch.set_num_threads(best)
    else:
        best = 0
    n = q.shape[0] if args.sample <= 0 else min(args.sample, q.shape[0])
    qq = q[:n].contiguous()
    for _ in range(args.warmup):
        hot_call(tracker, cfg, qq)
    if dev != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = hot_call(tracker, cfg, qq)
    if dev != "cpu":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    return ({
        "mode": "query", "device": dev, "n_query": n, "s_per_step": dt, "queries_per_s": n / dt, "threads": best,
        "host_cores": cores, "threads_tried_s_per_20000_queries": {str(k): round(v, 4) for k, v in threads_tried.items()},
        "n_probe": int(npm.neighbor_K), "map_points": int(npm.count()), "local_points": int(npm.local_neural_points.shape[0]),
        "occupied_probes_mean": n_occ, "valid_knn_mean": k_v, "nn_k": cfg.query_nn_k, "feature_dim": cfg.feature_dim,
        "decoder": f"{cfg.geo_mlp_level}x{cfg.geo_mlp_hidden_dim}", "weighted_first": bool(cfg.weighted_first),
        "buffer_size": int(cfg.buffer_size), "sdf_mean": float(out[0].mean()),
        "torch": torch.__version__, "kind": "reference"})


def run_frames(args):
    """BASELINE configs[2] with the reference's own Tracker / Mapper: tracker = 3 registration iterations, mapper =
    5 training iterations per frame, run_kitti.yaml, the same synthetic 64 x 1024 scans and preprocessing as
    pin_slam_b200/frame_loop.py; timed like the reference's time_table (tracking, mapping)."""
    import numpy as np
    import torch

    R, S = load_reference(), load_synthetic()
    dev = torch.device(args.device)
    cfg = make_config(R, "kitti", args.device)
    cfg.reg_iter_n = 3
    cfg.reg_term_thre_deg = 0.0
    cfg.reg_term_thre_m = 0.0
    torch.manual_seed(42)
    npm = R.NeuralPoints(cfg)
    dec = R.Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1).to(dev)
    decoders = {"sdf": dec, "semantic": None, "color": None}
    dataset = types.SimpleNamespace(processed_frame=0, odom_poses=np.zeros((0, 4, 4)), pgo_poses=None, gt_poses=None,
                                    gt_pose_provided=False, lose_track=False, stop_status=False, static_mask=None)
    tracker = R.Tracker(cfg, npm, decoders)
    mapper = R.Mapper(cfg, dataset, npm, decoders)
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    travel, poses, times, errs, preps = [0.0], [], [], [], []
    for f in range(2 + args.frames):
        gt = S.trajectory_pose(f)
        scan = S.lidar_scan(gt, seed=f, device=dev)
        scan = scan[R.voxel_down_sample_torch(scan, 0.08)]
        rng = scan.norm(dim=1)
        scan = scan[(rng > 3.0) & (rng < cfg.max_range) & (scan[:, 2] > -3.5)]
        source = scan[R.voxel_down_sample_torch(scan, 0.6)].contiguous()
        trk_s = 0.0
        if f == 0:
            pose = gt.to(dev)
        else:
            last = poses[-1]
            if len(poses) < 2:
                guess = (gt @ torch.linalg.inv(S.trajectory_pose(f - 1))).to(dev) @ last
            else:
                guess = last @ torch.linalg.inv(poses[-2]) @ last
            sync()
            t0 = time.perf_counter()
            pose, _, _, _ = tracker.tracking(source, guess, cur_ts=f)
            sync()
            trk_s = time.perf_counter() - t0
            pose = pose.clone()
        poses.append(pose)
        if f > 0:
            travel.append(travel[-1] + float((poses[-1][:3, 3] - poses[-2][:3, 3]).norm()))
        dataset.processed_frame = f
        dataset.odom_poses = torch.stack(poses).cpu().numpy()
        npm.travel_dist = torch.tensor(travel, device=dev, dtype=cfg.dtype)
        sync()
        t0 = time.perf_counter()
        mapper.process_frame(scan, None, pose, f)
        sync()
        prep_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        mapper.mapping(100 if f == 0 else 5)
        sync()
        map_s = time.perf_counter() - t0
        if f >= 2:
            times.append((trk_s, map_s))
            preps.append(prep_s)
        errs.append(float((pose[:3, 3].cpu() - gt[:3, 3]).norm()))
    trk = sorted(t for t, _ in times)[len(times) // 2]
    mp = sorted(m for _, m in times)[len(times) // 2]
    return ({
        "mode": "frames", "device": args.device, "frames": args.frames, "tracker_ms_median": trk * 1e3,
        "mapping_ms_median": mp * 1e3, "frames_per_s": 1.0 / (trk + mp),
        "prep_ms_median": sorted(preps)[len(preps) // 2] * 1e3,
        "frames_per_s_with_prep": 1.0 / (trk + mp + sorted(preps)[len(preps) // 2]), "source_points": int(source.shape[0]),
        "scan_points": int(scan.shape[0]), "local_map_points": int(npm.local_neural_points.shape[0]),
        "pool_samples": int(mapper.pool_sample_count), "translation_error_m_per_frame": [round(e, 4) for e in errs],
        "final_translation_error_m": errs[-1], "threads": torch.get_num_threads(), "torch": torch.__version__,
        "kind": "reference", "what": "unmodified reference Tracker.tracking (3 iterations) + Mapper.mapping(5) "
                                     "(oracle/_ref), wall clock with device synchronisation on both sides"})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["query", "frames"])
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sample", type=int, default=0, help="queries per step (0 = all 200 000)")
    ap.add_argument("--frames", type=int, default=12)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"mode": a.mode, "unavailable": "oracle/_ref is missing (run oracle/make_ref.py in the build container)"}))
        sys.exit(0)
    print(json.dumps((run_query if a.mode == "query" else run_frames)(a)))
