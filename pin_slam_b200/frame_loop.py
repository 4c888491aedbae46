"""The per-frame tracker + mapper loop of PIN-SLAM (reference: pin_slam.py:238-508, PGO / GUI / mesh
off) driven over seeded synthetic KITTI-shaped scans -- BASELINE configs[2]:
"full per-frame loop: tracker GN (3 iters) + mapper (5 train iters), KITTI 64-beam synthetic".

Only `tracker` (T3-T2) and `mapping` (T6-T5) are timed, like the reference's `time_table` columns the
frames/s metric is defined on; preprocessing and `Mapper.process_frame` (map growth,
sampling: section 8 f1/f2 "next" rows) run untimed between them.
"""
import time
import types

import numpy as np
import torch

from .config import HotPathConfig
from .model import Decoder, NeuralPoints
from .model.neural_points import voxel_down_sample
from .synthetic import lidar_scan, rgbd_frame, rgbd_pose, trajectory_pose
from .utils.mapper import Mapper
from .utils.tracker import Tracker


class FrameLoop:
    def __init__(self, device="cuda", n_track_iter=3, n_map_iter=5, seed=42, cfg=None, rgbd=False):
        """`rgbd`: BASELINE configs[3] -- Replica-shaped 640 x 480 RGB-D frames, run_replica.yaml (colour head,
        photometric registration); otherwise configs[2] -- 64 x 1024 KITTI-shaped scans, run_kitti.yaml."""
        self.rgbd = bool(rgbd)
        self.cfg = cfg or (HotPathConfig.replica(device=str(device)) if rgbd else HotPathConfig.kitti(device=str(device)))
        self.dev = torch.device(device)
        torch.manual_seed(seed)
        self.neural_points = NeuralPoints(self.cfg)
        self.sdf_mlp = Decoder(self.cfg, self.cfg.geo_mlp_hidden_dim, self.cfg.geo_mlp_level, 1)
        self.color_mlp = Decoder(self.cfg, self.cfg.color_mlp_hidden_dim, self.cfg.color_mlp_level,
                                 self.cfg.color_channel) if self.cfg.color_on else None
        decoders = {"sdf": self.sdf_mlp, "semantic": None, "color": self.color_mlp}
        self.dataset = types.SimpleNamespace(processed_frame=0, odom_poses=np.zeros((0, 4, 4)), pgo_poses=None,
                                             gt_poses=None, gt_pose_provided=False, lose_track=False,
                                             stop_status=False, static_mask=None)
        self.tracker = Tracker(self.cfg, self.neural_points, decoders)
        self.mapper = Mapper(self.cfg, self.dataset, self.neural_points, decoders)
        self.n_track_iter, self.n_map_iter = n_track_iter, n_map_iter
        self.travel = [0.0]
        self.poses = []
        self.times = []  # (tracker_ms, mapping_ms) per frame
        self.prep_times = []  # Mapper.process_frame (sampling, map growth, pool) per frame
        self.host_issue_times = []

    def preprocess(self, frame_id):
        """Synthetic scan -> voxel(0.08) + range crop -> map points; voxel(0.6) -> registration source.
        RGB-D: voxel(0.02) -> map points [x, y, z, r, g, b]; voxel(0.06) -> source (run_replica.yaml)."""
        if self.rgbd:
            gt = rgbd_pose(frame_id)
            pts, col = rgbd_frame(gt, seed=frame_id, device=self.dev)
            self.last_frame_points = pts.shape[0]
            keep = voxel_down_sample(pts, 0.02)
            scan = torch.cat((pts[keep], col[keep]), 1)
            source = scan[voxel_down_sample(scan[:, :3].contiguous(), 0.06)]
            return gt, scan.contiguous(), source.contiguous()
        gt = trajectory_pose(frame_id)
        scan = lidar_scan(gt, seed=frame_id, device=self.dev)
        scan = scan[voxel_down_sample(scan, 0.08)]
        rng = scan.norm(dim=1)
        scan = scan[(rng > 3.0) & (rng < self.cfg.max_range) & (scan[:, 2] > -3.5)]
        source = scan[voxel_down_sample(scan, 0.6)]
        return gt, scan.contiguous(), source.contiguous()

    def step(self, frame_id, timed=True, map_iters=None):
        cfg, npm = self.cfg, self.neural_points
        gt, scan, source = self.preprocess(frame_id)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        if frame_id == 0:
            pose = gt.to(self.dev)
            trk_ms = trk_cpu_ms = 0.0
        else:
            # constant-velocity initial guess (pin_slam.py: uniform motion model); the very first motion
            # estimate comes from the synthetic trajectory (a real run starts from rest)
            last = self.poses[-1]
            if len(self.poses) < 2:
                prev_gt = rgbd_pose(frame_id - 1) if self.rgbd else trajectory_pose(frame_id - 1)
                guess = (gt @ torch.linalg.inv(prev_gt)).to(self.dev) @ last
            else:
                guess = last @ torch.linalg.inv(self.poses[-2]) @ last
            ev[0].record()
            c0 = time.perf_counter()
            pose, _ = self.tracker.track_fixed(source[:, :3].contiguous(), guess, self.n_track_iter,
                                               source_colors=source[:, 3:].contiguous() if self.rgbd else None)
            trk_cpu_ms = (time.perf_counter() - c0) * 1e3
            ev[1].record()
            pose = pose.clone()
        self.poses.append(pose)
        # pose bookkeeping the dataset class does (slam_dataset.py:507): travel distance, odometry poses
        if frame_id > 0:
            step_len = float((self.poses[-1][:3, 3] - self.poses[-2][:3, 3]).norm())
            self.travel.append(self.travel[-1] + step_len)
        self.dataset.processed_frame = frame_id
        self.dataset.odom_poses = torch.stack(self.poses).cpu().numpy()
        npm.travel_dist = torch.tensor(self.travel, device=self.dev, dtype=cfg.dtype)
        ev[4].record()
        self.mapper.process_frame(scan, None, pose, frame_id)
        ev[5].record()
        n_iter = self.n_map_iter if map_iters is None else map_iters
        ev[2].record()
        c0 = time.perf_counter()
        self.mapper.mapping(n_iter)
        map_cpu_ms = (time.perf_counter() - c0) * 1e3
        ev[3].record()
        torch.cuda.synchronize()
        if frame_id > 0:
            trk_ms = ev[0].elapsed_time(ev[1])
        map_ms = ev[2].elapsed_time(ev[3])
        prep_ms = ev[4].elapsed_time(ev[5])  # "mapping preparation" (T5-T4 of the reference's time_table)
        if timed:
            self.times.append((trk_ms, map_ms))
            self.prep_times.append(prep_ms)
            self.host_issue_times.append((trk_cpu_ms, map_cpu_ms))  # host time to ISSUE the launches (no sync)
        err = float((pose[:3, 3].cpu() - gt[:3, 3]).norm())
        return {"frame": frame_id, "tracker_ms": trk_ms, "mapping_ms": map_ms, "prep_ms": prep_ms, "n_source": int(source.shape[0]),
                "n_scan": int(scan.shape[0]), "local_points": npm.local_count(), "pool": self.mapper.pool_sample_count,
                "trans_err_m": err}
