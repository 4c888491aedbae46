"""GPU parity of the frame-level entry points against fixtures recorded from the UNMODIFIED reference
(tests/golden/make_golden.py): the CUDA path itself -- not the oracle standing in for it, as in
tests/test_oracle_golden.py -- runs `Tracker.tracking` (full convergence loop), `Tracker.registration_step` with the
covariance / eigenvalue outputs, `Mapper.dynamic_filter`, `Mapper.process_frame`, `NeuralPoints.update` and
`reset_local_map` on the device.

The reference draws its random numbers on the CPU generator in these fixtures; `cpu_rng()` makes the drop-in's
device-side `torch.randn / rand / randint` calls draw from the CPU stream too (then move the result), so that the
sampler / feature-init streams are the fixture's and everything downstream can be compared exactly.
"""
import contextlib
import types

import numpy as np
import pytest
import torch

from oracle import pin_oracle as po
from tests.helpers import load_npz, t

pytestmark = pytest.mark.gpu
DEV = "cuda"
QUERY_FIXTURES = ["query_kitti_nwf", "query_kitti_wf", "query_cfg2_wf", "query_cfg2_nwf_pgo", "query_replica_wf_color"]


@contextlib.contextmanager
def cpu_rng():
    real = {n: getattr(torch, n) for n in ("randn", "rand", "randint")}

    def wrap(fn):
        def inner(*a, **k):
            dev = k.pop("device", None)
            out = fn(*a, **k)
            return out if dev is None else out.to(dev)

        return inner

    for n, fn in real.items():
        setattr(torch, n, wrap(fn))
    try:
        yield
    finally:
        for n, fn in real.items():
            setattr(torch, n, fn)


def _npm_from_fixture(fx, color=False, **cfg_kw):
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import NeuralPoints

    g = lambda k: fx["map." + k]  # noqa: E731
    cfg = HotPathConfig.kitti(device=DEV, feature_dim=int(g("geo_features").shape[1]), buffer_size=int(g("buffer_size")),
                              voxel_size_m=float(g("resolution")), local_map_radius=float(fx["cfg.local_map_radius"]),
                              color_on=color, **cfg_kw)
    npm = NeuralPoints(cfg)
    d = lambda k, dt=None: t(g(k), dt).to(DEV).clone()  # noqa: E731
    npm.neural_points = d("neural_points")
    npm.point_orientations = d("point_orientations")
    npm.geo_features = d("geo_features")
    if color:
        npm.color_features = d("color_features")
    npm.point_ts_create = d("point_ts_create")
    npm.point_ts_update = d("point_ts_update")
    npm.point_certainties = d("point_certainties")
    npm.travel_dist = d("travel_dist")
    npm.diff_travel_dist_local = float(g("diff_travel_dist_local"))
    npm.temporal_local_map_on = bool(g("temporal_local_map_on"))
    npm.after_pgo = bool(g("after_pgo"))
    npm.cur_ts = int(g("cur_ts"))
    table = torch.full((int(g("buffer_size")),), -1, dtype=torch.int32)
    table[t(g("table_slots"))] = t(g("table_vals")).to(torch.int32)
    npm.buffer_pt_index = table.to(DEV)
    return npm, cfg


def _decoder_from_fixture(cfg, fx, name, out_dim=1):
    from pin_slam_b200.model import Decoder

    levels = 0
    while f"{name}.layers.{levels}.weight" in fx:
        levels += 1
    dec = Decoder(cfg, 64, levels, out_dim)
    sd = {k[len(name) + 1:]: t(v).to(DEV) for k, v in fx.items() if k.startswith(name + ".") and not k.endswith("sdf_scale")}
    dec.load_state_dict(sd)
    return dec


def _trained_map():
    """The briefly trained KITTI-config map of track_kitti.npz with its local map re-derived on the device."""
    fx = load_npz("track_kitti")
    npm, cfg = _npm_from_fixture(fx)
    cfg.query_nn_k, cfg.weighted_first = int(fx["cfg.query_nn_k"]), bool(fx["cfg.weighted_first"])
    (min_g, max_g, gm_d, gm_g, lm, term_deg, term_m, surf_range, final_ratio, std_ratio,
     eig_thre) = (float(v) for v in fx["cfg.reg_floats"])
    iter_n, min_nn = (int(v) for v in fx["cfg.reg_ints"])
    cfg.reg_min_grad_norm, cfg.reg_max_grad_norm, cfg.reg_GM_dist_m, cfg.reg_GM_grad = min_g, max_g, gm_d, gm_g
    cfg.reg_lm_lambda, cfg.reg_term_thre_deg, cfg.reg_term_thre_m = lm, term_deg, term_m
    cfg.surface_sample_range_m, cfg.final_residual_ratio_thre, cfg.max_sdf_std_ratio = surf_range, final_ratio, std_ratio
    cfg.eigenvalue_ratio_thre, cfg.reg_iter_n, cfg.track_mask_query_nn_k = eig_thre, iter_n, min_nn
    # the reference's last reset_local_map ran in process_frame(frame 1) at the frame-1 sensor position
    npm.reset_local_map(torch.tensor([0.8, 0.0, 1.0], device=DEV), torch.eye(3, device=DEV), 1)
    assert np.array_equal(npm.local_mask.cpu().numpy(), fx["map.local_mask"])
    assert np.array_equal(npm.local_geo_features.data.cpu().numpy(), fx["map.local_geo_features"])
    dec = _decoder_from_fixture(cfg, fx, "sdf_mlp")
    return fx, npm, cfg, dec


def test_cuda_tracking_loop_matches_reference():
    """Tracker.tracking on the CUDA kernels (pinb200_track_iterations: K1 + K4 per iteration, pose on the device)
    against the reference's Tracker.tracking (utils/tracker.py:43-225): same iteration count and validity verdict,
    per-iteration residuals and valid counts, final pose."""
    from pin_slam_b200.utils.tracker import Tracker

    fx, npm, cfg, dec = _trained_map()
    tracker = Tracker(cfg, npm, {"sdf": dec, "semantic": None, "color": None})
    log = []
    inner = tracker._iterate

    def logged(*a, **k):
        o, res, sums = inner(*a, **k)
        r = res.cpu().numpy()
        log.append((float(r[17]), int(r[16])))
        return o, res, sums

    tracker._iterate = logged
    T, cov, _, valid = tracker.tracking(t(fx["source"]).to(DEV), t(fx["init_pose"]).double().to(DEV), cur_ts=1)
    torch.cuda.synchronize()
    assert valid == bool(fx["result.valid"])
    assert len(log) == int(fx["result.n_iter"])
    res = np.array([x[0] for x in log])
    np.testing.assert_allclose(res, fx["result.residual_cm"], rtol=2e-3, atol=2e-3)
    assert np.abs(np.array([x[1] for x in log]) - fx["result.valid_count"]).max() <= 2
    dT = np.abs(T.cpu().numpy() - fx["result.T"])
    print(f"[tracking] {len(log)} iterations, max |dT| rot {dT[:3, :3].max():.2e} trans {dT[:3, 3].max():.2e} m, "
          f"max residual dev {np.abs(res - fx['result.residual_cm']).max():.2e} cm")
    # <= 1e-6 m / 1e-7 rad per GN step; measured on the B200 after 18 chained steps: 2.3e-7 m, 3.4e-8
    assert dT[:3, 3].max() <= 2e-6 and dT[:3, :3].max() <= 5e-7


def test_cuda_covariance_and_eigenvalues_match_reference():
    """registration_step(..., vis_weight_pc=True): cov = inv(N_raw) * mean(w r^2) and the eigenvalues of the
    translation block of N (utils/tracker.py:680-693), consumed by the degeneracy check (:198-223)."""
    from pin_slam_b200.utils.tracker import Tracker

    fx, npm, cfg, dec = _trained_map()
    tracker = Tracker(cfg, npm, {"sdf": dec, "semantic": None, "color": None})
    pts = t(fx["regstep.points"]).to(DEV)
    gm_d = cfg.reg_GM_dist_m if cfg.reg_GM_dist_m > 0 else None
    gm_g = cfg.reg_GM_grad if cfg.reg_GM_grad > 0 else None
    T, cov, eig, _, valid_pts, res_cm, _ = tracker.registration_step(
        pts, None, torch.zeros(pts.shape[0], device=DEV), None, cfg.reg_min_grad_norm, cfg.reg_max_grad_norm, gm_d, gm_g,
        cfg.reg_lm_lambda, True)
    assert abs(valid_pts.shape[0] - int(fx["regstep.valid_count"])) <= 1
    assert abs(res_cm - float(fx["regstep.residual_cm"])) <= 2e-3 * float(fx["regstep.residual_cm"])
    np.testing.assert_allclose(np.sort(eig.numpy()), np.sort(fx["regstep.eig"].astype(np.float64)), rtol=2e-3)
    ref_cov = fx["regstep.cov"].astype(np.float64)
    scale = np.sqrt(np.outer(np.diag(ref_cov), np.diag(ref_cov)))
    assert (np.abs(cov.numpy() - ref_cov) / scale).max() <= 5e-3
    np.testing.assert_allclose(T.cpu().numpy(), fx["regstep.T"], rtol=0, atol=5e-6)


@pytest.mark.parametrize("type2", [True, False])
def test_cuda_dynamic_filter_matches_reference(type2):
    """Mapper.dynamic_filter (utils/mapper.py:99-137) on K1: static mask from certainty, SDF and gradient norm.  The
    decision thresholds are hard: points whose SDF / gradient norm sits within fp32 noise of a threshold may flip."""
    from pin_slam_b200.utils.mapper import Mapper

    fx, npm, cfg, dec = _trained_map()
    cert, sdf_ratio, min_grad, vox = (float(v) for v in fx["dyn.cfg"])
    cfg.dynamic_certainty_thre, cfg.dynamic_sdf_ratio_thre, cfg.dynamic_min_grad_norm_thre = cert, sdf_ratio, min_grad
    assert abs(cfg.voxel_size_m - vox) < 1e-9
    ds = types.SimpleNamespace(processed_frame=1, lose_track=False, stop_status=False, gt_pose_provided=False,
                               odom_poses=None, pgo_poses=None, gt_poses=None, static_mask=None)
    mapper = Mapper(cfg, ds, npm, {"sdf": dec, "semantic": None, "color": None})
    mask = mapper.dynamic_filter(t(fx["dyn.points"]).to(DEV), type_2_on=type2).cpu().numpy()
    ref = fx["dyn.mask_type2" if type2 else "dyn.mask_type1"]
    flips = int((mask != ref).sum())
    print(f"[dynamic_filter type2={type2}] {flips} of {ref.shape[0]} decisions differ; static fraction {ref.mean():.3f}")
    assert flips <= max(2, int(0.002 * ref.shape[0]))
    assert 0.0 < ref.mean() <= 1.0


def test_cuda_map_growth_matches_reference():
    """NeuralPoints.update over three frames on the device (model/neural_points.py:311-422): same points in the same
    order, timestamps, hash table, feature initialisation (CPU RNG stream), local map."""
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import NeuralPoints

    fx = load_npz("growth_kitti")
    g = lambda k: fx["map." + k]  # noqa: E731
    cfg = HotPathConfig.kitti(device=DEV, buffer_size=int(g("buffer_size")), feature_std=float(fx["cfg.feature_std"]),
                              local_map_radius=float(fx["cfg.local_map_radius"]))
    torch.manual_seed(int(fx["seed"]))
    with cpu_rng():
        npm = NeuralPoints(cfg)
        npm.diff_travel_dist_local = float(g("diff_travel_dist_local"))
        npm.travel_dist = t(g("travel_dist")).to(DEV)
        for f in range(int(fx["n_frames"])):
            npm.update(t(fx[f"frame{f}.points"]).to(DEV), t(fx[f"frame{f}.pos"]).to(DEV), torch.eye(3, device=DEV), f)
            assert npm.count() == int(fx[f"frame{f}.count"])
    c = lambda x: x.detach().cpu().numpy()  # noqa: E731
    assert np.array_equal(c(npm.neural_points), g("neural_points"))
    assert np.array_equal(c(npm.point_ts_create), g("point_ts_create"))
    assert np.array_equal(c(npm.point_ts_update), g("point_ts_update"))
    assert np.array_equal(c(npm.geo_features), g("geo_features"))
    table = torch.full((npm.buffer_size,), -1, dtype=torch.int64)
    table[t(g("table_slots"))] = t(g("table_vals")).long()
    assert torch.equal(npm.buffer_pt_index.long().cpu(), table)
    assert np.array_equal(c(npm.local_mask), g("local_mask"))
    assert np.array_equal(c(npm.global2local).astype(np.int64), g("global2local").astype(np.int64))
    assert np.array_equal(c(npm.local_geo_features.data), g("local_geo_features"))
    # and the kernels see the grown map: the probe index answers every local point's own cell with that point
    o = npm.query_sdf(npm.local_neural_points.contiguous(), _plain_decoder(cfg), need_grad=False, save_knn=True)
    near = o["knn_idx"][:, 0].cpu()
    assert bool((near == torch.arange(near.shape[0])).float().mean() > 0.99)


def _plain_decoder(cfg):
    from pin_slam_b200.model import Decoder

    torch.manual_seed(0)
    return Decoder(cfg, 64, 1, 1)


@pytest.mark.parametrize("name", QUERY_FIXTURES)
def test_cuda_reset_local_map_and_write_back(name):
    """reset_local_map / assign_local_to_global on the device reproduce the reference's local map bit for bit
    (model/neural_points.py:424-527), incl. the global2local fill quirk."""
    fx = load_npz(name)
    g = lambda k: fx["map." + k]  # noqa: E731
    color = "map.color_features" in fx
    npm, cfg = _npm_from_fixture(fx, color=color)
    npm.reset_local_map(t(fx["sensor_pos"]).to(DEV), torch.eye(3, device=DEV), int(g("cur_ts")))
    c = lambda x: x.detach().cpu().numpy()  # noqa: E731
    assert np.array_equal(c(npm.local_mask), g("local_mask"))
    assert np.array_equal(c(npm.global2local).astype(np.int64), g("global2local").astype(np.int64))
    for mine, ref in ((npm.local_neural_points, "local_neural_points"),
                      (npm.local_point_orientations, "local_point_orientations"),
                      (npm.local_geo_features.data, "local_geo_features"),
                      (npm.local_point_certainties, "local_point_certainties"),
                      (npm.local_point_ts_update, "local_point_ts_update")):
        assert np.array_equal(c(mine), g(ref)), ref
    ref_geo, ref_cert = t(g("geo_features")).clone(), t(g("point_certainties")).clone()
    npm.local_geo_features.data += 1.0
    npm.local_point_certainties += 0.5
    mask = t(g("local_mask"))
    ref_geo[mask] = npm.local_geo_features.data.cpu()
    ref_cert[mask[:-1]] = npm.local_point_certainties.cpu()
    npm.assign_local_to_global()
    assert torch.equal(npm.geo_features.cpu(), ref_geo) and torch.equal(npm.point_certainties.cpu(), ref_cert)


def test_cuda_process_frame_matches_reference():
    """Mapper.process_frame over three frames on the device (utils/mapper.py:162-449): per-ray sampling (CPU RNG
    stream), map growth, replay-pool append + window filter, new-sample selection through the CUDA
    pinb200_query_certainty."""
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import Decoder, NeuralPoints
    from pin_slam_b200.utils.mapper import Mapper

    fx = load_npz("frames_kitti")
    n_frames = int(fx["n_frames"])
    cfg = HotPathConfig.kitti(device=DEV, buffer_size=2000003, local_map_radius=float(fx["cfg.local_map_radius"]),
                              pool_filter_freq=int(fx["cfg.pool_filter_freq"]), adaptive_iters=True)
    cfg.window_radius = float(fx["cfg.window_radius"])
    torch.manual_seed(int(fx["seed"]))
    with cpu_rng():
        npm = NeuralPoints(cfg)
        npm.diff_travel_dist_local = 4.5
        npm.travel_dist = t(fx["travel_dist"]).to(DEV)
        dec = Decoder(cfg, 64, 1, 1)
        poses = fx["poses"]
        dataset = types.SimpleNamespace(processed_frame=0, lose_track=False, stop_status=False, gt_pose_provided=False,
                                        odom_poses=poses.copy(), pgo_poses=None, gt_poses=None, static_mask=None)
        mapper = Mapper(cfg, dataset, npm, {"sdf": dec, "semantic": None, "color": None})
        for f in range(n_frames):
            dataset.processed_frame = f
            torch.manual_seed(int(fx["seed"]) * 100 + f)
            mapper.process_frame(t(fx[f"frame{f}.points"]).to(DEV), None,
                                 torch.tensor(poses[f], dtype=torch.float64, device=DEV), f)
            assert npm.count() == int(fx[f"frame{f}.map_count"])
            assert mapper.pool_sample_count == int(fx[f"frame{f}.pool_sample_count"])
            assert mapper.cur_sample_count == int(fx[f"frame{f}.cur_sample_count"])
            assert np.array_equal(mapper.new_idx.cpu().numpy(), fx[f"frame{f}.new_idx"])
            assert mapper.adaptive_iter_offset == int(fx[f"frame{f}.adaptive_iter_offset"])
    c = lambda x: x.cpu().numpy()  # noqa: E731
    np.testing.assert_allclose(c(mapper.coord_pool), fx["pool.coord"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(c(mapper.global_coord_pool), fx["pool.global_coord"], rtol=1e-6, atol=1e-5)
    # the sampler's products round differently on the device (fma contraction): a few 1e-6 relative outliers
    np.testing.assert_allclose(c(mapper.sdf_label_pool), fx["pool.sdf_label"], rtol=5e-6, atol=2e-6)
    np.testing.assert_allclose(c(mapper.weight_pool), fx["pool.weight"], rtol=5e-6, atol=2e-6)
    assert np.array_equal(c(mapper.time_pool), fx["pool.time"])


def test_cuda_radius_search_with_time_filter_matches_reference():
    """The drop-in radius_neighborhood_search(time_filtering=True) (round-1 advice: it raised TypeError) returns the
    reference's ids and squared distances with the travel-distance window applied, and the unfiltered ones without
    (model/neural_points.py:950-1009; fixture keys rs.* / rs_nofilter.*)."""
    fx = load_npz("query_kitti_nwf")
    npm, cfg = _npm_from_fixture(fx)
    npm.reset_local_map(t(fx["sensor_pos"]).to(DEV), torch.eye(3, device=DEV), int(fx["map.cur_ts"]))
    q = t(fx["q"]).to(DEV)
    d2, idx = npm.radius_neighborhood_search(q, time_filtering=True)
    assert idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy().astype(np.int32), fx["rs.idx"])
    assert np.array_equal(d2.cpu().numpy(), fx["rs.dist2"])
    _, idxg = npm.radius_neighborhood_search(q, time_filtering=False)
    assert np.array_equal(idxg.cpu().numpy().astype(np.int32), fx["rs_nofilter.idx"])


@pytest.mark.parametrize("n,voxel", [(65536, 0.08), (45000, 0.6), (3000, 0.4), (17, 0.4), (200000, 0.05)])
def test_cuda_voxel_downsample_equals_reference_formulation(n, voxel):
    """pinb200_voxel_downsample (hash set + compaction + sort of the winners) returns exactly what the reference's
    unique / scatter-amin formulation returns (utils/tools.py:583-668): same winners, same ascending-key order, for the
    centre-distance and the min-value variants, with duplicated points (ties -> smaller index)."""
    from pin_slam_b200 import ops
    from pin_slam_b200.model import neural_points as npmod

    g = torch.Generator().manual_seed(n)
    pts = (torch.rand(n, 3, generator=g) * torch.tensor([40.0, 30.0, 6.0]) - torch.tensor([20.0, 15.0, 1.0]))
    pts[n // 2:n // 2 + n // 10] = pts[:n // 10]  # exact duplicates: ties on the quantised distance
    val = torch.randint(0, 50, (n,), generator=g).float()
    cpu_a = npmod.voxel_down_sample(pts, voxel)              # torch formulation (CPU tensors take that path)
    cpu_b = npmod.voxel_down_sample_min_value(pts, voxel, val)
    got_a = ops.voxel_downsample(pts.to(DEV), voxel)
    got_b = ops.voxel_downsample(pts.to(DEV), voxel, val.to(DEV))
    assert torch.equal(got_a.cpu(), cpu_a)
    assert torch.equal(got_b.cpu(), cpu_b)


def test_unchanged_reference_tracker_runs_fused():
    """The reference's unmodified utils/tracker.py (oracle/_ref) over the install()ed drop-ins: its
    query_feature -> Decoder.sdf -> get_gradient sequence is served by the fused K1 kernel (lazy feature handles,
    pin_slam_b200/model/fused_features.py) and returns the fused path's values; the eager path stays available."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isfile(os.path.join(root, "oracle", "_ref", "utils", "tracker.py")):
        pytest.skip("oracle/_ref not vendored (built by __graft_entry__.build() where /root/reference exists)")
    pr = subprocess.run([sys.executable, os.path.join(root, "tests", "unchanged_caller_check.py")], capture_output=True,
                        text=True, timeout=600)
    assert pr.returncode == 0, pr.stdout[-2000:] + pr.stderr[-4000:]
    res = json.loads([l for l in pr.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    for name, r in res.items():
        # one kNN launch for the weights / counts / certainty + one fused launch per decoder call
        assert r["launches"] <= 12, (name, r)  # incl. the one-off probe-index build of a fresh map handle
        assert r["sdf_err"] == 0.0 and r["grad_err"] == 0.0 and r["mask_equal"], (name, r)
        assert r["cert_err"] <= 1e-5, (name, r)
        if "color_err" in r:
            assert r["color_err"] == 0.0 and r["cgrad_err"] == 0.0, (name, r)
        assert r["eager_vs_fused_sdf"] <= 2e-6 and r["eager_vs_fused_grad"] <= 2e-4 * max(r["grad_scale"], 1e-3) * 50, (name, r)


def test_cuda_loop_closure_transform_matches_reference():
    """Row f4 on the device: pinb200_frame_transform (adjust_map / transform_data_pool of the reference,
    model/neural_points.py:791-822, utils/mapper.py:527-531) against the reference's loop-closure fixture and, for
    both time-stamp modes, against the torch formulation on random data."""
    from pin_slam_b200 import ops
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import NeuralPoints
    from pin_slam_b200.model.neural_points import _quat_multiply, _rotmat_to_quat

    fx = load_npz("loop_kitti")
    g = lambda k: fx["map." + k]  # noqa: E731
    cfg = HotPathConfig.kitti(device=DEV, feature_dim=int(g("geo_features").shape[1]), buffer_size=int(g("buffer_size")),
                              voxel_size_m=float(g("resolution")), local_map_radius=float(fx["cfg.local_map_radius"]))
    npm = NeuralPoints(cfg)
    npm.config.use_mid_ts = bool(fx["cfg.use_mid_ts"])
    npm.neural_points = t(g("neural_points")).to(DEV)
    npm.point_orientations = t(g("point_orientations")).to(DEV)
    npm.point_ts_create = t(g("point_ts_create")).to(DEV)
    npm.point_ts_update = t(g("point_ts_update")).to(DEV)
    npm.adjust_map(t(fx["pose_diff"]).to(DEV))
    assert npm.after_pgo
    np.testing.assert_allclose(npm.neural_points.cpu().numpy(), fx["adjusted.neural_points"], rtol=1e-6, atol=2e-6)
    q, qr = npm.point_orientations.cpu().numpy(), fx["adjusted.point_orientations"]
    sign = np.sign((q * qr).sum(1, keepdims=True))  # q and -q are the same rotation
    np.testing.assert_allclose(q * sign, qr, rtol=1e-5, atol=1e-6)

    gen = torch.Generator().manual_seed(0)
    n, nts = 200003, 37
    xyz = torch.randn(n, 3, generator=gen) * 30
    quat = torch.randn(n, 4, generator=gen)
    quat = quat / quat.norm(dim=1, keepdim=True)
    ts_a = torch.randint(0, nts, (n,), generator=gen).int()
    ts_b = torch.randint(0, nts, (n,), generator=gen).int()
    pose = torch.eye(4, dtype=torch.float64).repeat(nts, 1, 1)
    for i in range(nts):
        pose[i, :3, :3] = po.expmap(0.05 * torch.randn(3, generator=gen, dtype=torch.float64))
        pose[i, :3, 3] = torch.randn(3, generator=gen, dtype=torch.float64)
    dq_frames = _rotmat_to_quat(pose[:, :3, :3])
    for mid in (False, True):
        ts = ((ts_a + ts_b) / 2).int().long() if mid else ts_a.long()
        tf = pose[ts].float()
        ref_xyz = (tf[:, :3, :3] @ xyz.unsqueeze(-1)).squeeze(-1) + tf[:, :3, 3]
        ref_q = _quat_multiply(dq_frames[ts].float(), quat)
        x, qd = xyz.to(DEV).clone(), quat.to(DEV).clone()
        ops.frame_transform(x, ts_a.to(DEV), pose.to(DEV), quat=qd, dquat=dq_frames.to(DEV), ts_b=ts_b.to(DEV) if mid else None)
        np.testing.assert_allclose(x.cpu().numpy(), ref_xyz.numpy(), rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(qd.cpu().numpy(), ref_q.numpy(), rtol=1e-6, atol=1e-6)
    # pool form: no orientations
    x = xyz.to(DEV).clone()
    ops.frame_transform(x, ts_a.to(DEV), pose.to(DEV))
    tf = pose[ts_a.long()].float()
    np.testing.assert_allclose(x.cpu().numpy(), ((tf[:, :3, :3] @ xyz.unsqueeze(-1)).squeeze(-1) + tf[:, :3, 3]).numpy(),
                               rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("kind", ["kitti", "replica"])
def test_cuda_ray_sampler_kernel_equals_torch_formulation(kind):
    """Row f2: pinb200_ray_samples against the torch port of utils/data_sampler.py:18-260 (which the CPU tests pin to
    the reference's fixtures) on the same RNG draws: coordinates, labels, weights, colours, ray-major order."""
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.utils.mapper import DataSampler

    cfg = HotPathConfig.kitti(device=DEV) if kind == "kitti" else HotPathConfig.replica(device=DEV)
    g = torch.Generator().manual_seed(5)
    n = 7013
    pts = (torch.randn(n, 3, generator=g) * (20.0 if kind == "kitti" else 2.0)).to(DEV)
    col = torch.rand(n, 3, generator=g).to(DEV) if kind == "replica" else None
    smp = DataSampler(cfg)
    out = {}
    for fused in (True, False):
        DataSampler.FUSED = fused
        try:
            torch.manual_seed(11)
            out[fused] = smp.sample(pts, None, None, col)
        finally:
            DataSampler.FUSED = True
    a, b = out[True], out[False]
    for x, y, tol in ((a[0], b[0], 2e-6), (a[1], b[1], 2e-6), (a[5], b[5], 2e-6)):
        np.testing.assert_allclose(x.cpu().numpy(), y.cpu().numpy(), rtol=tol, atol=tol * 10)
    assert (a[4] is None) == (b[4] is None)
    if a[4] is not None:
        assert torch.equal(a[4], b[4])
    assert torch.equal(torch.sign(a[5]), torch.sign(b[5]))
    # both paths leave the generator in the same state
    torch.manual_seed(11)
    smp.sample(pts, None, None, col)
    r1 = torch.rand(4, device=DEV)
    DataSampler.FUSED = False
    try:
        torch.manual_seed(11)
        smp.sample(pts, None, None, col)
        r2 = torch.rand(4, device=DEV)
    finally:
        DataSampler.FUSED = True
    assert torch.equal(r1, r2)
