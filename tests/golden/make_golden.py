#!/usr/bin/env python3
"""Generate golden fixtures by running the UNMODIFIED reference (PRBonn/PIN_SLAM,
mounted read-only at /root/reference) on seeded synthetic inputs, CPU, fp32.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference has no tests/golden vectors of its own, so these
fixtures are what pins the oracle (``oracle/pin_oracle.py``) and, through it,
the CUDA path.  Each fixture stores the full map state, the inputs, and the
outputs of the reference functions on the hot path:

  NeuralPoints.radius_neighborhood_search   model/neural_points.py:950
  NeuralPoints.query_feature                model/neural_points.py:530
  NeuralPoints.query_certainty              model/neural_points.py:1011
  Decoder.sdf / regress_color               model/decoder.py:83,112
  Tracker.query_source_points               utils/tracker.py:227
  Tracker.registration_step / implicit_reg  utils/tracker.py:367,615
  Mapper.mapping                            utils/mapper.py:600
  Mesher.query_points (dense grid query)    utils/mesher.py:40
  DataSampler.sample (per-ray samples)      utils/data_sampler.py:18
  NeuralPoints.adjust_map / recreate_hash   model/neural_points.py:791,820 (loop-closure map adjustment)
  NeuralPoints.update (map growth)          model/neural_points.py:311
  Mapper.process_frame (sampling + pool)    utils/mapper.py:162

Optional reference imports (open3d, gtsam, ...) that are absent here and unused
by the hot path are stubbed in sys.modules before import.
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

for _m in ["open3d", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "roma", "wandb", "natsort",
           "skimage", "skimage.measure", "pypose", "gtsam", "dtyper", "pyquaternion", "laspy", "evo"]:
    sys.modules[_m] = MagicMock()
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

from model.decoder import Decoder  # noqa: E402
from model.neural_points import NeuralPoints  # noqa: E402
from utils.config import Config  # noqa: E402
from utils.mapper import Mapper  # noqa: E402
from utils.tracker import Tracker  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def scene_points(n, seed, lo=-10.0, hi=10.0):
    """Random points on a floor, two walls and a box (metres)."""
    g = torch.Generator().manual_seed(seed)
    u = lambda k: torch.rand(k, generator=g) * (hi - lo) + lo  # noqa: E731
    k = n // 4
    floor = torch.stack([u(k), u(k), torch.zeros(k)], 1)
    wall1 = torch.stack([u(k), torch.full((k,), hi), torch.rand(k, generator=g) * 4], 1)
    wall2 = torch.stack([torch.full((k,), lo), u(k), torch.rand(k, generator=g) * 4], 1)
    box = torch.stack([torch.rand(k, generator=g) * 3 + 2, torch.rand(k, generator=g) * 3 - 5,
                       torch.full((k,), 1.5)], 1)
    p = torch.cat([floor, wall1, wall2, box], 0)
    return p + 0.01 * torch.randn(p.shape, generator=g)


def make_config(kind):
    cfg = Config()
    if kind == "kitti":
        cfg.load("/root/reference/config/lidar_slam/run_kitti.yaml")
    elif kind == "replica":
        cfg.load("/root/reference/config/rgbd_slam/run_replica.yaml")
        cfg.voxel_size_m = 0.25  # keep the fixture small (synthetic scene is 20 m wide)
    else:  # cfg2-like: defaults + non-default sizes reachable via yaml
        cfg.feature_dim = 32
        cfg.query_nn_k = 8
        cfg.geo_mlp_level = 2
        cfg.voxel_size_m = 0.4
        cfg.track_on = True
    cfg.device = "cpu"
    cfg.pgo_on = False
    cfg.silence = True
    cfg.setup_dtype()
    torch.set_default_dtype(cfg.dtype)
    return cfg


def map_state(npm):
    tab = npm.buffer_pt_index
    occ = torch.nonzero(tab >= 0).flatten()
    d = dict(
        resolution=np.float64(npm.resolution),
        buffer_size=np.int64(npm.buffer_size),
        neural_points=npm.neural_points.numpy(),
        point_orientations=npm.point_orientations.numpy(),
        geo_features=npm.geo_features.numpy(),
        point_ts_create=npm.point_ts_create.numpy(),
        point_ts_update=npm.point_ts_update.numpy(),
        point_certainties=npm.point_certainties.numpy(),
        table_slots=occ.numpy(),
        table_vals=tab[occ].numpy(),
        travel_dist=npm.travel_dist.numpy(),
        cur_ts=np.int64(npm.cur_ts),
        diff_travel_dist_local=np.float64(npm.diff_travel_dist_local),
        max_valid_dist2=np.float64(npm.max_valid_dist2),
        neighbor_dx=npm.neighbor_dx.numpy(),
        after_pgo=np.bool_(npm.after_pgo),
        temporal_local_map_on=np.bool_(npm.temporal_local_map_on),
        local_mask=npm.local_mask.numpy(),
        global2local=npm.global2local.numpy(),
        local_neural_points=npm.local_neural_points.numpy(),
        local_point_orientations=npm.local_point_orientations.numpy(),
        local_geo_features=npm.local_geo_features.detach().numpy(),
        local_point_certainties=npm.local_point_certainties.numpy(),
        local_point_ts_update=npm.local_point_ts_update.numpy(),
    )
    if npm.color_features is not None:
        d["color_features"] = npm.color_features.numpy()
        d["local_color_features"] = npm.local_color_features.detach().numpy()
    return {"map." + k: np.array(v, copy=True) for k, v in d.items()}  # copies: later in-place scatters must not leak in


def dec_state(dec, name):
    d = {}
    for k, v in dec.state_dict().items():
        d[f"{name}.{k}"] = v.numpy().copy()
    d[f"{name}.sdf_scale"] = np.float64(dec.sdf_scale)
    return d


def build_reference_map(cfg, seed, n_frames=3, radius=9.0, diff_td=4.5, feat_std=0.1):
    """Grow a map through the reference's own NeuralPoints.update()."""
    torch.manual_seed(seed)
    cfg.local_map_radius = radius
    npm = NeuralPoints(cfg)
    npm.diff_travel_dist_local = diff_td
    npm.travel_dist = torch.tensor([0.0, 2.0, 4.0, 6.0, 8.0][: n_frames + 1])
    for f in range(n_frames):
        pts = scene_points(6000, seed * 10 + f)
        pts[:, 0] += 1.5 * f
        pos = torch.tensor([1.5 * f, 0.0, 1.0])
        npm.update(pts, pos, torch.eye(3), f)
    # non-zero features / certainties so that errors cannot hide behind zeros
    g = torch.Generator().manual_seed(seed + 100)
    npm.geo_features = feat_std * torch.randn(npm.geo_features.shape, generator=g)
    if npm.color_features is not None:
        npm.color_features = feat_std * torch.randn(npm.color_features.shape, generator=g)
    npm.point_certainties = torch.rand(npm.count(), generator=g) * 3.0
    npm.reset_local_map(pos, torch.eye(3), n_frames - 1)
    return npm, pos


def query_points_near(npm, n, seed, sigma=0.15):
    g = torch.Generator().manual_seed(seed)
    sel = torch.randint(0, npm.count(), (n,), generator=g)
    q = npm.neural_points[sel] + sigma * torch.randn(n, 3, generator=g)
    # a few far-away queries (nn_count == 0 rows) and negative coordinates
    q[:8] = torch.tensor([500.0, -300.0, 40.0]) + torch.randn(8, 3, generator=g)
    return q.contiguous()


def gen_query_fixture(kind, seed, weighted_first, after_pgo=False, color=False, nq=640, name=None):
    cfg = make_config(kind)
    cfg.weighted_first = weighted_first
    cfg.buffer_size = 40009  # small table -> real hash collisions are exercised
    if color:
        cfg.color_on = True
        cfg.color_channel = 3
    npm, pos = build_reference_map(cfg, seed)
    if after_pgo:
        g = torch.Generator().manual_seed(seed + 7)
        qn = torch.randn(npm.count(), 4, generator=g)
        npm.point_orientations = qn / qn.norm(dim=1, keepdim=True)
        npm.after_pgo = True
        npm.reset_local_map(pos, torch.eye(3), npm.cur_ts)
    torch.manual_seed(seed + 1)
    sdf_mlp = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    color_mlp = Decoder(cfg, cfg.color_mlp_hidden_dim, cfg.color_mlp_level, cfg.color_channel) if color else None
    out = {}
    out.update(map_state(npm))
    out.update(dec_state(sdf_mlp, "sdf_mlp"))
    if color:
        out.update(dec_state(color_mlp, "color_mlp"))
    out["cfg.query_nn_k"] = np.int64(cfg.query_nn_k)
    out["cfg.weighted_first"] = np.bool_(cfg.weighted_first)
    out["cfg.feature_dim"] = np.int64(cfg.feature_dim)
    out["cfg.local_map_radius"] = np.float64(cfg.local_map_radius)
    out["sensor_pos"] = pos.numpy()

    q = query_points_near(npm, nq, seed + 2)
    out["q"] = q.numpy()

    # a3
    d2, idx = npm.radius_neighborhood_search(q.clone(), time_filtering=True)
    out["rs.dist2"] = d2.numpy()
    out["rs.idx"] = idx.numpy().astype(np.int32)
    d2g, idxg = npm.radius_neighborhood_search(q.clone(), time_filtering=False)
    out["rs_nofilter.idx"] = idxg.numpy().astype(np.int32)
    # a5
    out["query_certainty"] = npm.query_certainty(q.clone()).numpy()

    # a4 (inference mode, local + global)
    for loc in (True, False):
        geo, col, w, nnc, cert = npm.query_feature(q.clone(), training_mode=False, query_locally=loc,
                                                   query_color_feature=color)
        tag = "qf_local" if loc else "qf_global"
        out[f"{tag}.geo"] = geo.detach().numpy()
        if col is not None:
            out[f"{tag}.color"] = col.detach().numpy()
        out[f"{tag}.weight"] = w.numpy()
        out[f"{tag}.nn_counts"] = nnc.numpy()
        out[f"{tag}.certainty"] = cert.numpy()

    # a8: Tracker.query_source_points
    trk = Tracker(cfg, npm, {"sdf": sdf_mlp, "semantic": None, "color": color_mlp})
    photo = color
    res = trk.query_source_points(q.clone(), cfg.infer_bs, True, True, color, photo,
                                  query_locally=True, mask_min_nn_count=cfg.track_mask_query_nn_k)
    sdf_pred, sdf_grad, color_pred, color_grad, _, mask, certainty, sdf_std = res
    out["trk.sdf"] = sdf_pred.numpy()
    out["trk.grad"] = sdf_grad.numpy()
    out["trk.mask"] = mask.numpy()
    out["trk.certainty"] = certainty.numpy()
    out["trk.sdf_std"] = sdf_std.numpy()
    out["cfg.track_mask_query_nn_k"] = np.int64(cfg.track_mask_query_nn_k)
    if color:
        out["trk.color"] = color_pred.numpy()
        out["trk.color_grad"] = color_grad.numpy()

    # a9/a10: one registration step on the same points (geometry branch)
    if not color:
        cfg_photo = cfg.photometric_loss_on
        cfg.photometric_loss_on = False
        source_sdf = torch.zeros(q.shape[0])
        gnorm = None
        # random-init decoder => tiny |grad|; widen the validity band so the GN system is non-trivial
        min_gn, max_gn = 1e-3, 10.0
        T, cov, eig, _, valid_points, res_cm, _ = trk.registration_step(
            q.clone(), None, source_sdf, None, min_gn, max_gn,
            cfg.reg_GM_dist_m, cfg.reg_GM_grad, cfg.reg_lm_lambda, False)
        cfg.photometric_loss_on = cfg_photo
        out["reg.T"] = T.numpy()
        out["reg.valid_count"] = np.int64(valid_points.shape[0])
        out["reg.residual_cm"] = np.float64(res_cm)
        out["reg.params"] = np.array([min_gn, max_gn,
                                      cfg.surface_sample_range_m * cfg.max_sdf_std_ratio,
                                      cfg.reg_GM_dist_m, cfg.reg_GM_grad, cfg.reg_lm_lambda], dtype=np.float64)

    # a9/a11: registration with colour -- photometric term (implicit_color_reg) and consistency weight
    if color:
        g2 = torch.Generator().manual_seed(seed + 31)
        src_col = torch.rand(q.shape[0], 3, generator=g2)
        out["reg_color.source_colors"] = src_col.numpy()
        min_gn, max_gn = 1e-6, 10.0  # sdf_scale is 0.0055 in the Replica config: random-decoder gradients are tiny
        for tag, photo in (("photo", True), ("consist", False)):
            cfg.photometric_loss_on = photo
            T, _, _, _, valid_points, res_cm, col_res = trk.registration_step(
                q.clone(), None, torch.zeros(q.shape[0]), src_col.clone(), min_gn, max_gn,
                cfg.reg_GM_dist_m, cfg.reg_GM_grad, cfg.reg_lm_lambda, False)
            out[f"reg_color.{tag}.T"] = T.numpy()
            out[f"reg_color.{tag}.valid_count"] = np.int64(valid_points.shape[0])
            out[f"reg_color.{tag}.residual_cm"] = np.float64(res_cm)
            out[f"reg_color.{tag}.color_residual"] = np.float64(-1.0 if col_res is None else col_res)
        out["reg_color.params"] = np.array([min_gn, max_gn, cfg.surface_sample_range_m * cfg.max_sdf_std_ratio,
                                            cfg.reg_GM_dist_m, cfg.reg_GM_grad, cfg.reg_lm_lambda,
                                            cfg.photometric_loss_weight], dtype=np.float64)

    # a4 training-mode side effects (certainty scatter_add, ts amax) on a copy
    ts = torch.full((q.shape[0],), int(npm.cur_ts), dtype=torch.int32)
    ts[::3] = int(npm.cur_ts) + 1
    cert_before = npm.local_point_certainties.clone()
    tsu_before = npm.local_point_ts_update.clone()
    npm.query_feature(q.clone(), ts, training_mode=True, query_color_feature=color)
    out["train_fx.ts"] = ts.numpy()
    out["train_fx.certainties_after"] = npm.local_point_certainties.numpy().copy()
    out["train_fx.ts_update_after"] = npm.local_point_ts_update.numpy().copy()
    npm.local_point_certainties = cert_before
    npm.local_point_ts_update = tsu_before

    name = name or f"query_{kind}_{'wf' if weighted_first else 'nwf'}{'_pgo' if after_pgo else ''}{'_color' if color else ''}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "Mg=", npm.count(), "M=", npm.local_count(), "valid=", int(out["reg.valid_count"]) if "reg.valid_count" in out else "-")


def gen_train_fixture(kind, seed, weighted_first, iters=3, bs=2048, color=False, name=None):
    cfg = make_config(kind)
    cfg.weighted_first = weighted_first
    cfg.buffer_size = 40009
    cfg.bs = bs
    cfg.bs_new_sample = 0
    if color:
        cfg.color_on = True
        cfg.color_channel = 3
    npm, pos = build_reference_map(cfg, seed, feat_std=0.05)
    torch.manual_seed(seed + 1)
    sdf_mlp = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    color_mlp = Decoder(cfg, cfg.color_mlp_hidden_dim, cfg.color_mlp_level, cfg.color_channel) if color else None
    dataset = types.SimpleNamespace(processed_frame=npm.cur_ts, lose_track=False, stop_status=False,
                                    gt_pose_provided=False, odom_poses=None, pgo_poses=None, gt_poses=None)
    mapper = Mapper(cfg, dataset, npm, {"sdf": sdf_mlp, "semantic": None, "color": color_mlp})
    # synthetic replay pool: samples around the surface with sdf labels along a fake normal
    g = torch.Generator().manual_seed(seed + 5)
    npool = 20000
    sel = torch.randint(0, npm.count(), (npool,), generator=g)
    label = 0.15 * torch.randn(npool, generator=g)
    label[::4] = 0.0
    coord = npm.neural_points[sel] + 0.1 * torch.randn(npool, 3, generator=g)
    coord[:, 2] += label
    mapper.global_coord_pool = coord
    mapper.coord_pool = coord.clone()
    mapper.sdf_label_pool = label
    mapper.weight_pool = (torch.rand(npool, generator=g) * 0.8 + 0.6) * torch.where(label.abs() < 0.05, 1.0, -1.0)
    mapper.time_pool = torch.randint(0, npm.cur_ts + 1, (npool,), generator=g).to(torch.int32)
    mapper.color_pool = torch.rand(npool, 3, generator=g) if color else None
    mapper.sem_label_pool = None
    mapper.normal_label_pool = None
    mapper.pool_sample_count = npool
    mapper.used_poses = torch.eye(4, dtype=torch.float64).repeat(npm.cur_ts + 1, 1, 1)

    out = {}
    out.update(map_state(npm))
    out.update(dec_state(sdf_mlp, "sdf_mlp"))
    if color:
        out.update(dec_state(color_mlp, "color_mlp"))
    for k in ["query_nn_k", "feature_dim", "bs", "gradient_decimation", "iters"]:
        out["cfg." + k] = np.int64(getattr(cfg, k))
    out["cfg.weighted_first"] = np.bool_(cfg.weighted_first)
    out["cfg.loss_weight_on"] = np.bool_(cfg.loss_weight_on)
    out["cfg.floats"] = np.array([cfg.sigma_sigmoid_m, mapper.sdf_scale, cfg.weight_e,
                                  cfg.voxel_size_m * cfg.num_grad_step_ratio, cfg.lr, cfg.adam_eps,
                                  cfg.weight_decay, cfg.surface_sample_range_m, cfg.weight_i], dtype=np.float64)

    batches = []
    orig = mapper.get_batch

    def rec(global_coord=False):
        b = orig(global_coord)
        batches.append(b)
        return b

    mapper.get_batch = rec
    torch.manual_seed(seed + 9)
    mapper.mapping(iters)
    out["n_iters"] = np.int64(len(batches))
    for i, b in enumerate(batches):
        coord_b, label_b, ts_b, _, _, color_b, weight_b = b
        out[f"batch{i}.coord"] = coord_b.detach().numpy()
        out[f"batch{i}.sdf_label"] = label_b.numpy()
        out[f"batch{i}.ts"] = ts_b.numpy()
        out[f"batch{i}.weight"] = weight_b.numpy()
        if color_b is not None:
            out[f"batch{i}.color"] = color_b.numpy()
    out["after.local_geo_features"] = npm.local_geo_features.detach().numpy()
    if color:
        out["after.local_color_features"] = npm.local_color_features.detach().numpy()
    out["after.local_point_certainties"] = npm.local_point_certainties.numpy()
    out["after.local_point_ts_update"] = npm.local_point_ts_update.numpy()
    out.update(dec_state(sdf_mlp, "after.sdf_mlp"))
    if color:
        out.update(dec_state(color_mlp, "after.color_mlp"))
    name = name or f"train_{kind}_{'wf' if weighted_first else 'nwf'}{'_color' if color else ''}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "iters", len(batches))


def gen_mesh_fixture(kind, seed, weighted_first, color=False, name=None):
    """Mesher.query_points on a regular grid through the global map (utils/mesher.py:40-164)."""
    from utils.mesher import Mesher

    cfg = make_config(kind)
    cfg.weighted_first = weighted_first
    cfg.buffer_size = 40009
    if color:
        cfg.color_on = True
        cfg.color_channel = 3
    npm, pos = build_reference_map(cfg, seed)
    torch.manual_seed(seed + 1)
    sdf_mlp = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    color_mlp = Decoder(cfg, cfg.color_mlp_hidden_dim, cfg.color_mlp_level, cfg.color_channel) if color else None
    mesher = Mesher(cfg, npm, {"sdf": sdf_mlp, "semantic": None, "color": color_mlp})
    out = {}
    out.update(map_state(npm))
    out.update(dec_state(sdf_mlp, "sdf_mlp"))
    if color:
        out.update(dec_state(color_mlp, "color_mlp"))
    out["cfg.query_nn_k"] = np.int64(cfg.query_nn_k)
    out["cfg.weighted_first"] = np.bool_(cfg.weighted_first)
    out["cfg.feature_dim"] = np.int64(cfg.feature_dim)
    # regular grid around the sensor (spacing 0.35 m: several grid nodes per voxel, many nodes in free space
    # with fewer than 4 or zero neighbours)
    ax = torch.arange(-14, 15, dtype=torch.float32) * 0.35
    az = torch.arange(-4, 5, dtype=torch.float32) * 0.35
    grid = torch.stack(torch.meshgrid(ax, ax, az, indexing="ij"), -1).reshape(-1, 3) + pos.reshape(1, 3).float()
    out["grid"] = grid.numpy()
    sdf, _, col, mask = mesher.query_points(grid.clone(), 1000, query_sdf=True, query_sem=False, query_color=color,
                                            query_mask=True, query_locally=False, mask_min_nn_count=4)
    out["mesh.sdf"] = np.asarray(sdf, np.float32)
    out["mesh.mask"] = np.asarray(mask).astype(np.bool_)
    if color:
        out["mesh.color"] = np.asarray(col, np.float32)
    name = name or f"mesh_{kind}_{'wf' if weighted_first else 'nwf'}{'_color' if color else ''}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "grid", grid.shape[0], "masked-in", int(np.asarray(mask).sum()),
          "no-neighbour rows", int((np.asarray(sdf) == 0).sum()))


def gen_sampler_fixture(kind, seed, color=False, n=257, name=None):
    """DataSampler.sample on seeded sensor-frame points (utils/data_sampler.py:18-260): torch.manual_seed(seed)
    immediately before the call pins the randn/rand stream."""
    from utils.data_sampler import DataSampler

    cfg = make_config(kind)
    g = torch.Generator().manual_seed(seed)
    direction = torch.randn(n, 3, generator=g)
    direction = direction / direction.norm(dim=1, keepdim=True)
    points = direction * (torch.rand(n, 1, generator=g) * 40.0 + 2.0)
    normals = torch.randn(n, 3, generator=g)
    colors = torch.rand(n, 3, generator=g) if color else None
    sampler = DataSampler(cfg)
    torch.manual_seed(seed)
    coord, label, normal, _, col, weight = sampler.sample(points.clone(), normals.clone(), None,
                                                          None if colors is None else colors.clone())
    out = {"points": points.numpy(), "normals": normals.numpy(), "seed": np.int64(seed),
           "cfg.ints": np.array([cfg.surface_sample_n, cfg.free_front_n, cfg.free_behind_n], np.int64),
           "cfg.floats": np.array([cfg.surface_sample_range_m, cfg.free_sample_begin_ratio, cfg.free_sample_end_dist_m,
                                   cfg.dist_weight_scale, cfg.max_range], np.float64),
           "cfg.flags": np.array([cfg.dist_weight_on, cfg.behind_dropoff_on], np.bool_),
           "out.coord": coord.numpy(), "out.label": label.numpy(), "out.normal": normal.numpy(),
           "out.weight": weight.numpy()}
    if color:
        out["colors"] = colors.numpy()
        out["out.color"] = col.numpy()
    name = name or f"sampler_{kind}{'_color' if color else ''}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, coord.shape)


def gen_loop_fixture(kind, seed, name=None):
    """Loop-closure map adjustment: adjust_map with per-frame pose corrections, then
    recreate_hash (kept_points=True and the duplicate-filtering kept_points=False)."""
    import copy

    cfg = make_config(kind)
    cfg.buffer_size = 40009
    npm, pos = build_reference_map(cfg, seed)
    out = {}
    out.update(map_state(npm))
    out["sensor_pos"] = pos.numpy()
    out["cfg.local_map_radius"] = np.float64(cfg.local_map_radius)
    out["cfg.use_mid_ts"] = np.bool_(cfg.use_mid_ts)
    g = torch.Generator().manual_seed(seed + 3)
    n_frames = int(npm.travel_dist.shape[0])
    ang = torch.randn(n_frames, 3, generator=g) * 0.05
    th = ang.norm(dim=1, keepdim=True).clamp(min=1e-9)
    ax = ang / th
    K = torch.zeros(n_frames, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -ax[:, 2], ax[:, 1], ax[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 0], -ax[:, 1], ax[:, 0]
    R = torch.eye(3).unsqueeze(0) + torch.sin(th).unsqueeze(-1) * K + (1 - torch.cos(th)).unsqueeze(-1) * (K @ K)
    pose_diff = torch.eye(4).repeat(n_frames, 1, 1)
    pose_diff[:, :3, :3] = R
    pose_diff[:, :3, 3] = torch.randn(n_frames, 3, generator=g) * 0.3
    out["pose_diff"] = pose_diff.numpy()
    npm.adjust_map(pose_diff)
    out["adjusted.neural_points"] = npm.neural_points.numpy().copy()
    out["adjusted.point_orientations"] = npm.point_orientations.numpy().copy()
    kept = copy.deepcopy(npm)
    kept.recreate_hash(pos, torch.eye(3), kept_points=True, with_ts=True, cur_ts=int(npm.cur_ts))
    occ = torch.nonzero(kept.buffer_pt_index >= 0).flatten()
    out["rehash_kept.table_slots"] = occ.numpy()
    out["rehash_kept.table_vals"] = kept.buffer_pt_index[occ].numpy()
    out["rehash_kept.local_mask"] = kept.local_mask.numpy().copy()
    filt = copy.deepcopy(npm)
    filt.recreate_hash(pos, torch.eye(3), kept_points=False, with_ts=True, cur_ts=int(npm.cur_ts))
    occ = torch.nonzero(filt.buffer_pt_index >= 0).flatten()
    out["rehash_filter.table_slots"] = occ.numpy()
    out["rehash_filter.table_vals"] = filt.buffer_pt_index[occ].numpy()
    out["rehash_filter.neural_points"] = filt.neural_points.numpy().copy()
    out["rehash_filter.geo_features"] = filt.geo_features.numpy().copy()
    out["rehash_filter.point_ts_create"] = filt.point_ts_create.numpy().copy()
    name = name or f"loop_{kind}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "points", npm.count(), "after duplicate filter", filt.count())


def gen_growth_fixture(kind, seed, n_frames=3, name=None):
    """Map growth through NeuralPoints.update (model/neural_points.py:311-422) with a table large enough that no two
    voxels collide (collision winners are unspecified in torch), non-zero feature init so the RNG stream is pinned."""
    cfg = make_config(kind)
    cfg.buffer_size = 2000003
    cfg.feature_std = 0.05
    cfg.local_map_radius = 9.0
    torch.manual_seed(seed)
    npm = NeuralPoints(cfg)
    npm.diff_travel_dist_local = 4.5
    npm.travel_dist = torch.tensor([0.0, 2.0, 4.0, 6.0, 8.0][: n_frames + 1])
    out = {"seed": np.int64(seed), "n_frames": np.int64(n_frames), "cfg.feature_std": np.float64(cfg.feature_std),
           "cfg.local_map_radius": np.float64(cfg.local_map_radius)}
    for f in range(n_frames):
        pts = scene_points(6000, seed * 10 + f)
        pts[:, 0] += 1.5 * f
        pos = torch.tensor([1.5 * f, 0.0, 1.0])
        out[f"frame{f}.points"] = pts.numpy().copy()
        out[f"frame{f}.pos"] = pos.numpy().copy()
        npm.update(pts, pos, torch.eye(3), f)
        out[f"frame{f}.count"] = np.int64(npm.count())
    out.update(map_state(npm))
    name = name or f"growth_{kind}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "points", npm.count())


def gen_frame_fixture(kind, seed, n_frames=3, name=None):
    """Mapper.process_frame over a few frames (utils/mapper.py:162-449): per-ray sampling, map growth, replay-pool
    append + window filter, and the new-sample selection through query_certainty.  Collision-free table; the RNG is
    re-seeded before every frame so that the sampler / pool-filter draws are pinned."""
    cfg = make_config(kind)
    cfg.buffer_size = 2000003
    cfg.local_map_radius = 9.0
    cfg.pool_filter_freq = 2       # exercise the window filter on frame 1
    cfg.window_radius = 12.0
    cfg.adaptive_iters = True
    torch.manual_seed(seed)
    npm = NeuralPoints(cfg)
    npm.diff_travel_dist_local = 4.5
    npm.travel_dist = torch.tensor([0.0, 2.0, 4.0, 6.0, 8.0][: n_frames + 1])
    sdf_mlp = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    poses = np.tile(np.eye(4), (n_frames, 1, 1))
    poses[:, 0, 3] = 1.5 * np.arange(n_frames)
    poses[:, 2, 3] = 1.0
    dataset = types.SimpleNamespace(processed_frame=0, lose_track=False, stop_status=False, gt_pose_provided=False,
                                    odom_poses=poses.copy(), pgo_poses=None, gt_poses=None, static_mask=None)
    mapper = Mapper(cfg, dataset, npm, {"sdf": sdf_mlp, "semantic": None, "color": None})
    out = {"seed": np.int64(seed), "n_frames": np.int64(n_frames), "poses": poses,
           "cfg.local_map_radius": np.float64(cfg.local_map_radius), "cfg.window_radius": np.float64(cfg.window_radius),
           "cfg.pool_filter_freq": np.int64(cfg.pool_filter_freq), "cfg.max_range": np.float64(cfg.max_range),
           "travel_dist": npm.travel_dist.numpy().copy()}
    for f in range(n_frames):
        world = scene_points(3000, seed * 10 + f)
        world[:, 0] += 1.5 * f
        pose = torch.tensor(poses[f], dtype=torch.float64)
        sensor = (world - pose[:3, 3].float())  # identity rotation
        dataset.processed_frame = f
        out[f"frame{f}.points"] = sensor.numpy().copy()
        torch.manual_seed(seed * 100 + f)
        mapper.process_frame(sensor.clone(), None, pose, f)
        out[f"frame{f}.pool_sample_count"] = np.int64(mapper.pool_sample_count)
        out[f"frame{f}.cur_sample_count"] = np.int64(mapper.cur_sample_count)
        out[f"frame{f}.map_count"] = np.int64(npm.count())
        out[f"frame{f}.new_idx"] = mapper.new_idx.numpy().copy()
        out[f"frame{f}.adaptive_iter_offset"] = np.int64(mapper.adaptive_iter_offset)
    out["pool.coord"] = mapper.coord_pool.numpy().copy()
    out["pool.global_coord"] = mapper.global_coord_pool.numpy().copy()
    out["pool.sdf_label"] = mapper.sdf_label_pool.numpy().copy()
    out["pool.weight"] = mapper.weight_pool.numpy().copy()
    out["pool.time"] = mapper.time_pool.numpy().copy()
    name = name or f"frames_{kind}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "pool", mapper.pool_sample_count, "map", npm.count(),
          "new", [int(out[f"frame{f}.new_idx"].shape[0]) for f in range(n_frames)])


def gen_track_fixture(kind, seed, name=None):
    """Tracker.tracking (utils/tracker.py:43-225), the full convergence loop, on a briefly trained map: two frames of
    process_frame + mapping, then the second scan is registered from a perturbed initial pose."""
    cfg = make_config(kind)
    cfg.buffer_size = 2000003
    cfg.local_map_radius = 12.0
    cfg.bs = 4096
    torch.manual_seed(seed)
    npm = NeuralPoints(cfg)
    npm.diff_travel_dist_local = 6.0
    npm.travel_dist = torch.tensor([0.0, 1.0, 2.0])
    sdf_mlp = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    decoders = {"sdf": sdf_mlp, "semantic": None, "color": None}
    poses = np.tile(np.eye(4), (2, 1, 1))
    poses[1, 0, 3] = 0.8
    poses[:, 2, 3] = 1.0
    dataset = types.SimpleNamespace(processed_frame=0, lose_track=False, stop_status=False, gt_pose_provided=False,
                                    odom_poses=poses.copy(), pgo_poses=None, gt_poses=None, static_mask=None)
    mapper = Mapper(cfg, dataset, npm, decoders)
    tracker = Tracker(cfg, npm, decoders)
    scans = []
    for f in range(2):
        world = scene_points(8000, seed * 10 + f)
        pose = torch.tensor(poses[f], dtype=torch.float64)
        sensor = world - pose[:3, 3].float()
        scans.append(sensor)
        dataset.processed_frame = f
        torch.manual_seed(seed * 100 + f)
        mapper.process_frame(sensor.clone(), None, pose, f)
        mapper.mapping(40)
    out = {}
    out.update(map_state(npm))
    out.update(dec_state(sdf_mlp, "sdf_mlp"))
    out["cfg.query_nn_k"] = np.int64(cfg.query_nn_k)
    out["cfg.weighted_first"] = np.bool_(cfg.weighted_first)
    out["cfg.local_map_radius"] = np.float64(cfg.local_map_radius)
    src = scans[1][::7].contiguous()  # ~1100 source points
    ang = 0.02
    init = torch.tensor(poses[1], dtype=torch.float64)
    init[:3, :3] = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    init[:3, 3] += torch.tensor([0.12, -0.08, 0.03], dtype=torch.float64)
    out["source"] = src.numpy().copy()
    out["init_pose"] = init.numpy().copy()
    calls = []
    orig = tracker.registration_step

    def counted(*a, **k):
        r = orig(*a, **k)
        calls.append((float(r[5]), int(r[4].shape[0])))
        return r

    tracker.registration_step = counted
    T, cov, _, valid = tracker.tracking(src.clone(), init.clone(), cur_ts=1)
    out["result.T"] = T.numpy().copy()
    out["result.valid"] = np.bool_(valid)
    out["result.n_iter"] = np.int64(len(calls))
    out["result.residual_cm"] = np.array([c[0] for c in calls])
    out["result.valid_count"] = np.array([c[1] for c in calls])
    cfg_floats = [cfg.reg_min_grad_norm, cfg.reg_max_grad_norm, cfg.reg_GM_dist_m, cfg.reg_GM_grad, cfg.reg_lm_lambda,
                  cfg.reg_term_thre_deg, cfg.reg_term_thre_m, cfg.surface_sample_range_m,
                  cfg.final_residual_ratio_thre, cfg.max_sdf_std_ratio, cfg.eigenvalue_ratio_thre]
    out["cfg.reg_floats"] = np.array(cfg_floats, np.float64)
    out["cfg.reg_ints"] = np.array([cfg.reg_iter_n, cfg.track_mask_query_nn_k], np.int64)
    # one explicit registration step with the covariance / eigenvalue outputs switched on (utils/tracker.py:680-693;
    # consumed by the degeneracy check at :198-223) on the points at the converged pose
    tracker.registration_step = orig
    from utils.tools import transform_torch
    pts_final = transform_torch(src.clone(), T)
    gm_d = cfg.reg_GM_dist_m if cfg.reg_GM_dist_m > 0 else None
    gm_g = cfg.reg_GM_grad if cfg.reg_GM_grad > 0 else None
    rs = tracker.registration_step(pts_final, None, torch.zeros(src.shape[0]), None, cfg.reg_min_grad_norm,
                                   cfg.reg_max_grad_norm, gm_d, gm_g, cfg.reg_lm_lambda, True)
    out["regstep.points"] = pts_final.detach().numpy().copy()
    out["regstep.T"] = rs[0].numpy().copy()
    out["regstep.cov"] = rs[1].detach().numpy().copy()
    out["regstep.eig"] = rs[2].detach().numpy().copy()
    out["regstep.valid_count"] = np.int64(rs[4].shape[0])
    out["regstep.residual_cm"] = np.float64(rs[5])
    # Mapper.dynamic_filter (utils/mapper.py:99-137) on the second scan in the world frame, both strategies
    world_pts = (scans[1] + torch.tensor(poses[1][:3, 3], dtype=torch.float32)).contiguous()
    dyn = world_pts[::3].clone()
    mask2 = mapper.dynamic_filter(dyn.clone(), type_2_on=True)
    mask1 = mapper.dynamic_filter(dyn.clone(), type_2_on=False)
    out["dyn.points"] = dyn.numpy().copy()
    out["dyn.mask_type2"] = mask2.numpy().copy()
    out["dyn.mask_type1"] = mask1.numpy().copy()
    out["dyn.cfg"] = np.array([cfg.dynamic_certainty_thre, cfg.dynamic_sdf_ratio_thre, cfg.dynamic_min_grad_norm_thre,
                               cfg.voxel_size_m], np.float64)
    name = name or f"track_{kind}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    err = np.linalg.norm(T.numpy()[:3, 3] - poses[1][:3, 3])
    print("wrote", name, "iters", len(calls), "valid", valid, "residual_cm", [round(c[0], 2) for c in calls][:12],
          "final translation error %.3f m (init 0.147)" % err)


if __name__ == "__main__":
    torch.set_num_threads(4)
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only == "track":
        gen_track_fixture("kitti", 71)
        sys.exit(0)
    if only == "frames":
        gen_frame_fixture("kitti", 61)
        sys.exit(0)
    if only == "growth":
        gen_growth_fixture("kitti", 51)
        sys.exit(0)
    if only == "loop":
        gen_loop_fixture("kitti", 41)
        sys.exit(0)
    if only == "sampler":
        gen_sampler_fixture("kitti", 31)
        gen_sampler_fixture("replica", 32, color=True)
        sys.exit(0)
    if only == "mesh":  # the dense-grid query fixtures only
        gen_mesh_fixture("kitti", 21, weighted_first=False)
        gen_mesh_fixture("replica", 22, weighted_first=True, color=True)
        sys.exit(0)
    if only == "replica_query":  # regenerate a single fixture (the reference's map growth is not bit-reproducible
        gen_query_fixture("replica", 5, weighted_first=True, color=True)  # run to run: duplicate-slot index_put)
        sys.exit(0)
    gen_query_fixture("kitti", 1, weighted_first=False)
    gen_query_fixture("kitti", 2, weighted_first=True)
    gen_query_fixture("cfg2", 3, weighted_first=True)
    gen_query_fixture("cfg2", 4, weighted_first=False, after_pgo=True)
    gen_query_fixture("replica", 5, weighted_first=True, color=True)
    gen_train_fixture("kitti", 11, weighted_first=False)
    gen_train_fixture("cfg2", 12, weighted_first=True)
    gen_train_fixture("replica", 13, weighted_first=True, color=True)
    gen_mesh_fixture("kitti", 21, weighted_first=False)
    gen_mesh_fixture("replica", 22, weighted_first=True, color=True)
    gen_sampler_fixture("kitti", 31)
    gen_sampler_fixture("replica", 32, color=True)
    gen_loop_fixture("kitti", 41)
    gen_growth_fixture("kitti", 51)
    gen_frame_fixture("kitti", 61)
    gen_track_fixture("kitti", 71)
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by tests/golden/make_golden.py from /root/reference (PRBonn/PIN_SLAM)\n"
                f"torch {torch.__version__} cpu fp32, numpy {np.__version__}\n")
