"""Pin the CPU oracle (oracle/pin_oracle.py) against outputs of the unmodified
reference recorded by tests/golden/make_golden.py.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import pin_oracle as po
from tests.helpers import decoder_from_fixture, load_npz, map_from_fixture, t

QUERY_FIXTURES = ["query_kitti_nwf", "query_kitti_wf", "query_cfg2_wf", "query_cfg2_nwf_pgo", "query_replica_wf_color"]
TRAIN_FIXTURES = ["train_kitti_nwf", "train_cfg2_wf", "train_replica_wf_color"]


def close(a, b, rtol=1e-6, atol=1e-7):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def close_frac(a, b, atol=2e-6, rtol=1e-5, max_bad_frac=2e-3, max_abs=2e-3):
    """Adam with eps=1e-15 turns rounding-level gradient differences (BLAS thread count, summation
    order) into O(1e-5) parameter differences on a handful of near-zero-gradient elements; allow a
    small fraction of such outliers, bounded in magnitude."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    assert bad.mean() <= max_bad_frac, f"{bad.sum()} / {bad.size} elements differ"
    assert np.abs(a - b).max() <= max_abs


@pytest.mark.parametrize("name", QUERY_FIXTURES)
def test_radius_search_bit_exact(name):
    fx = load_npz(name)
    m = map_from_fixture(fx)
    q = t(fx["q"])
    d2, idx = po.radius_search(m, q.clone(), time_filtering=True)
    assert np.array_equal(idx.numpy(), fx["rs.idx"].astype(np.int64))
    assert np.array_equal(d2.numpy(), fx["rs.dist2"])
    _, idxg = po.radius_search(m, q.clone(), time_filtering=False)
    assert np.array_equal(idxg.numpy(), fx["rs_nofilter.idx"].astype(np.int64))
    close(po.query_certainty(m, q.clone()), fx["query_certainty"], 0, 0)


@pytest.mark.parametrize("name", QUERY_FIXTURES)
def test_probe_offsets_and_local_map(name):
    fx = load_npz(name)
    m = map_from_fixture(fx)
    assert np.array_equal(po.probe_offsets(2, 0.2).numpy(), fx["map.neighbor_dx"])
    m2 = m.clone()
    po.reset_local_map(m2, t(fx["sensor_pos"]), float(fx["cfg.local_map_radius"]), int(fx["map.cur_ts"]))
    assert np.array_equal(m2.local_mask.numpy(), fx["map.local_mask"])
    assert np.array_equal(m2.global2local.numpy(), fx["map.global2local"])
    assert np.array_equal(m2.local_geo_features.numpy(), fx["map.local_geo_features"])


@pytest.mark.parametrize("name", QUERY_FIXTURES)
@pytest.mark.parametrize("local", [True, False])
def test_query_feature(name, local):
    fx = load_npz(name)
    m = map_from_fixture(fx)
    q = t(fx["q"])
    k = int(fx["cfg.query_nn_k"])
    wf = bool(fx["cfg.weighted_first"])
    color = "map.color_features" in fx
    geo, col, w, nnc, cert = po.query_feature(m, q.clone(), None, k, wf, training_mode=False,
                                              query_locally=local, query_color_feature=color)
    tag = "qf_local" if local else "qf_global"
    assert np.array_equal(nnc.numpy(), fx[tag + ".nn_counts"])
    close(w, fx[tag + ".weight"], 0, 0)
    close(geo, fx[tag + ".geo"], 0, 0)
    close(cert, fx[tag + ".certainty"], 0, 0)
    if color:
        close(col, fx[tag + ".color"], 0, 0)


@pytest.mark.parametrize("name", QUERY_FIXTURES)
def test_training_side_effects(name):
    fx = load_npz(name)
    m = map_from_fixture(fx)
    k = int(fx["cfg.query_nn_k"])
    wf = bool(fx["cfg.weighted_first"])
    po.query_feature(m, t(fx["q"]).clone(), t(fx["train_fx.ts"]), k, wf, training_mode=True,
                     query_color_feature="map.color_features" in fx)
    close(m.local_point_certainties, fx["train_fx.certainties_after"], 1e-6, 1e-6)
    assert np.array_equal(m.local_point_ts_update.numpy(), fx["train_fx.ts_update_after"])


@pytest.mark.parametrize("name", QUERY_FIXTURES)
def test_tracker_query(name):
    fx = load_npz(name)
    m = map_from_fixture(fx)
    dec = decoder_from_fixture(fx, "sdf_mlp")
    color = "map.color_features" in fx
    cdec = decoder_from_fixture(fx, "color_mlp") if color else None
    k = int(fx["cfg.query_nn_k"])
    wf = bool(fx["cfg.weighted_first"])
    out = po.query_sdf(m, dec, t(fx["q"]), k, wf, color_dec=cdec, color_grad=color)
    close(out["sdf"], fx["trk.sdf"], 1e-6, 1e-8)
    close(out["grad"], fx["trk.grad"], 1e-5, 1e-8)
    close(out["sdf_std"], fx["trk.sdf_std"], 1e-5, 1e-8)
    close(out["certainty"], fx["trk.certainty"], 1e-6, 1e-7)
    assert np.array_equal((out["nn_count"] >= int(fx["cfg.track_mask_query_nn_k"])).numpy(), fx["trk.mask"])
    if color:
        close(out["color"], fx["trk.color"], 1e-6, 1e-7)
        close(out["color_grad"], fx["trk.color_grad"], 1e-5, 1e-8)


@pytest.mark.parametrize("name", [n for n in QUERY_FIXTURES if "color" not in n])
def test_registration_step(name):
    fx = load_npz(name)
    q = t(fx["q"])
    mn, mx, max_std, gmd, gmg, lam = [float(v) for v in fx["reg.params"]]
    out = po.registration_step(
        q, t(fx["trk.sdf"]), t(fx["trk.grad"]), t(fx["trk.sdf_std"]),
        torch.where(t(fx["trk.mask"]), 100, 0), torch.zeros(q.shape[0]),
        1, mn, mx, max_std, gmd, gmg, lam)
    assert out["valid_count"] == int(fx["reg.valid_count"])
    close(out["residual_cm"], float(fx["reg.residual_cm"]), 1e-6, 0)
    # fp32 J^T W J summation order depends on the BLAS thread count; the random-decoder system is
    # poorly conditioned, so compare the fp64 solve at 2e-5
    close(out["T"], fx["reg.T"], 2e-5, 2e-6)


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_mapping_iterations(name):
    fx = load_npz(name)
    m = map_from_fixture(fx)
    dec = decoder_from_fixture(fx, "sdf_mlp")
    color = "map.color_features" in fx
    cdec = decoder_from_fixture(fx, "color_mlp") if color else None
    k = int(fx["cfg.query_nn_k"])
    wf = bool(fx["cfg.weighted_first"])
    sigma_m, sdf_scale, weight_e, eps_num, lr, adam_eps, wd, surf_range, weight_i = [float(v) for v in fx["cfg.floats"]]
    m.local_geo_features.requires_grad_(True)
    feats = [m.local_geo_features]
    if color:
        m.local_color_features.requires_grad_(True)
        feats.append(m.local_color_features)
    dec.requires_grad_(True)
    groups = [dec.tensors()]
    if color:
        cdec.requires_grad_(True)
        groups.append(cdec.tensors())
    groups.append(feats)
    opt = po.make_adam(groups, lr, adam_eps, wd)
    for i in range(int(fx["n_iters"])):
        loss, _ = po.mapping_loss(
            m, dec, t(fx[f"batch{i}.coord"]), t(fx[f"batch{i}.sdf_label"]), t(fx[f"batch{i}.ts"]),
            t(fx[f"batch{i}.weight"]), k, wf, sdf_scale, bool(fx["cfg.loss_weight_on"]), weight_e,
            int(fx["cfg.gradient_decimation"]), eps_num, color_dec=cdec,
            color_label=t(fx[f"batch{i}.color"]) if color else None, surface_range=surf_range, weight_i=weight_i)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    close_frac(m.local_geo_features.detach(), fx["after.local_geo_features"])
    close_frac(m.local_point_certainties, fx["after.local_point_certainties"], 1e-5, 1e-5)
    assert np.array_equal(m.local_point_ts_update.numpy(), fx["after.local_point_ts_update"])
    for i, (w, b) in enumerate(dec.hidden):
        close_frac(w.detach(), fx[f"after.sdf_mlp.layers.{i}.weight"])
        close_frac(b.detach(), fx[f"after.sdf_mlp.layers.{i}.bias"])
    close_frac(dec.out[0].detach(), fx["after.sdf_mlp.lout.weight"])
    if color:
        close_frac(m.local_color_features.detach(), fx["after.local_color_features"])
        close_frac(cdec.out[0].detach(), fx["after.color_mlp.lout.weight"])


@pytest.mark.parametrize("mode", ["photo", "consist"])
def test_registration_step_with_colour(mode):
    """utils/tracker.py:493-542 + implicit_color_reg (:699-744) on the Replica-config fixture."""
    fx = load_npz("query_replica_wf_color")
    q = t(fx["q"])
    mn, mx, max_std, gmd, gmg, lam, w_photo = [float(v) for v in fx["reg_color.params"]]
    out = po.registration_step(
        q, t(fx["trk.sdf"]), t(fx["trk.grad"]), t(fx["trk.sdf_std"]), torch.where(t(fx["trk.mask"]), 100, 0),
        torch.zeros(q.shape[0]), 1, mn, mx, max_std, gmd, gmg, lam, colors=t(fx["reg_color.source_colors"]),
        color_pred=t(fx["trk.color"]), color_grad=t(fx["trk.color_grad"]), photo_loss_on=(mode == "photo"),
        w_photo=w_photo)
    assert out["valid_count"] == int(fx[f"reg_color.{mode}.valid_count"])
    close(out["residual_cm"], float(fx[f"reg_color.{mode}.residual_cm"]), 1e-6, 0)
    if mode == "photo":
        close(out["color_residual_mean"], float(fx["reg_color.photo.color_residual"]), 1e-6, 0)
    close(out["T"], fx[f"reg_color.{mode}.T"], 2e-4, 2e-5)


@pytest.mark.parametrize("name", ["mesh_kitti_nwf", "mesh_replica_wf_color"])
def test_mesher_grid_query(name):
    """Oracle restatement of Mesher.query_points (utils/mesher.py:40-164) against the reference's output."""
    fx = load_npz(name)
    m = map_from_fixture(fx)
    dec = decoder_from_fixture(fx, "sdf_mlp")
    cdec = decoder_from_fixture(fx, "color_mlp") if "mesh.color" in fx else None
    sdf, col, mask = po.mesher_query_points(m, dec, t(fx["grid"]), int(fx["cfg.query_nn_k"]),
                                            bool(fx["cfg.weighted_first"]), color_dec=cdec)
    assert np.array_equal(mask.numpy(), fx["mesh.mask"])
    assert np.array_equal(sdf.numpy() == 0, fx["mesh.sdf"] == 0)
    np.testing.assert_allclose(sdf.numpy(), fx["mesh.sdf"], rtol=1e-5, atol=1e-6 * dec.sdf_scale)
    if cdec is not None:
        np.testing.assert_allclose(col.numpy(), fx["mesh.color"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", QUERY_FIXTURES)
def test_dropin_reset_local_map_and_write_back(name):
    """The drop-in NeuralPoints.reset_local_map (index-list gathers, one host sync) reproduces the reference's local
    map bit for bit -- local_mask, global2local incl. the +1 fill quirk, every local array -- and
    assign_local_to_global writes the trained rows back where the reference does (model/neural_points.py:424-527).
    Pure torch host logic: runs on the CPU."""
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import NeuralPoints

    fx = load_npz(name)
    g = lambda k: fx["map." + k]  # noqa: E731
    color = "map.color_features" in fx
    cfg = HotPathConfig.kitti(device="cpu", feature_dim=int(g("geo_features").shape[1]), buffer_size=int(g("buffer_size")),
                              voxel_size_m=float(g("resolution")), local_map_radius=float(fx["cfg.local_map_radius"]),
                              color_on=color)
    npm = NeuralPoints(cfg)
    npm.neural_points = t(g("neural_points"))
    npm.point_orientations = t(g("point_orientations"))
    npm.geo_features = t(g("geo_features")).clone()
    if color:
        npm.color_features = t(g("color_features")).clone()
    npm.point_ts_create = t(g("point_ts_create"))
    npm.point_ts_update = t(g("point_ts_update")).clone()
    npm.point_certainties = t(g("point_certainties")).clone()
    npm.travel_dist = t(g("travel_dist"))
    npm.diff_travel_dist_local = float(g("diff_travel_dist_local"))
    npm.temporal_local_map_on = bool(g("temporal_local_map_on"))
    npm.after_pgo = bool(g("after_pgo"))
    npm.reset_local_map(t(fx["sensor_pos"]), torch.eye(3), int(g("cur_ts")))
    assert np.array_equal(npm.local_mask.numpy(), g("local_mask"))
    assert np.array_equal(npm.global2local.numpy().astype(np.int64), g("global2local").astype(np.int64))
    for mine, ref in ((npm.local_neural_points, "local_neural_points"),
                      (npm.local_point_orientations, "local_point_orientations"),
                      (npm.local_geo_features.data, "local_geo_features"),
                      (npm.local_point_certainties, "local_point_certainties"),
                      (npm.local_point_ts_update, "local_point_ts_update")):
        assert np.array_equal(mine.numpy(), g(ref)), ref
    if color:
        assert np.array_equal(npm.local_color_features.data.numpy(), g("local_color_features"))
    # write-back: what the reference does with boolean-mask assignment (neural_points.py:515-527)
    ref_geo, ref_cert = t(g("geo_features")).clone(), t(g("point_certainties")).clone()
    npm.local_geo_features.data += 1.0
    npm.local_point_certainties += 0.5
    mask = t(g("local_mask"))
    ref_geo[mask] = npm.local_geo_features.data
    ref_cert[mask[:-1]] = npm.local_point_certainties
    npm.assign_local_to_global()
    assert torch.equal(npm.geo_features, ref_geo) and torch.equal(npm.point_certainties, ref_cert)


@pytest.mark.parametrize("name", ["sampler_kitti", "sampler_replica_color"])
def test_dropin_data_sampler_matches_reference(name):
    """SURVEY.md section 8 row f2 (host side): the drop-in DataSampler.sample consumes the RNG stream exactly like the
    reference (utils/data_sampler.py:18-260) and returns the same samples, labels and weights."""
    import types

    from pin_slam_b200.utils.mapper import DataSampler

    fx = load_npz(name)
    ns, nf, nb = (int(v) for v in fx["cfg.ints"])
    sr, begin, end, dscale, max_range = (float(v) for v in fx["cfg.floats"])
    cfg = types.SimpleNamespace(device="cpu", surface_sample_n=ns, free_front_n=nf, free_behind_n=nb,
                                surface_sample_range_m=sr, free_sample_begin_ratio=begin, free_sample_end_dist_m=end,
                                dist_weight_scale=dscale, max_range=max_range, dist_weight_on=bool(fx["cfg.flags"][0]),
                                behind_dropoff_on=bool(fx["cfg.flags"][1]))
    colors = t(fx["colors"]) if "colors" in fx else None
    torch.manual_seed(int(fx["seed"]))
    coord, label, normal, _, col, weight = DataSampler(cfg).sample(t(fx["points"]), t(fx["normals"]), None, colors)
    np.testing.assert_allclose(coord.numpy(), fx["out.coord"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(label.numpy(), fx["out.label"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(weight.numpy(), fx["out.weight"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(normal.numpy(), fx["out.normal"])
    if colors is not None:
        assert np.array_equal(col.numpy(), fx["out.color"])


def _dropin_from_fixture(fx, color=False):
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import NeuralPoints

    g = lambda k: fx["map." + k]  # noqa: E731
    cfg = HotPathConfig.kitti(device="cpu", feature_dim=int(g("geo_features").shape[1]), buffer_size=int(g("buffer_size")),
                              voxel_size_m=float(g("resolution")), local_map_radius=float(fx["cfg.local_map_radius"]),
                              color_on=color)
    npm = NeuralPoints(cfg)
    npm.neural_points = t(g("neural_points")).clone()
    npm.point_orientations = t(g("point_orientations")).clone()
    npm.geo_features = t(g("geo_features")).clone()
    if color:
        npm.color_features = t(g("color_features")).clone()
    npm.point_ts_create = t(g("point_ts_create")).clone()
    npm.point_ts_update = t(g("point_ts_update")).clone()
    npm.point_certainties = t(g("point_certainties")).clone()
    npm.travel_dist = t(g("travel_dist"))
    npm.diff_travel_dist_local = float(g("diff_travel_dist_local"))
    npm.temporal_local_map_on = bool(g("temporal_local_map_on"))
    npm.cur_ts = int(g("cur_ts"))
    table = torch.full((int(g("buffer_size")),), -1, dtype=torch.int32)
    table[t(g("table_slots"))] = t(g("table_vals")).to(torch.int32)
    npm.buffer_pt_index = table
    return npm


def _assert_same_hash_table(npm, ref_slots, ref_vals):
    """Same occupied slots; same owner wherever a slot has a single candidate.  Where two voxels collide in the
    (deliberately small) table the winner of the duplicate-index write is unspecified in torch -- for the reference
    as well -- so there the owner only has to be one of the colliding points."""
    table = torch.full((npm.buffer_size,), -1, dtype=torch.int64)
    table[t(ref_slots)] = t(ref_vals).long()
    mine = npm.buffer_pt_index.long()
    assert torch.equal(mine >= 0, table >= 0)
    differ = torch.nonzero(mine != table).flatten()
    for s in differ.tolist():
        a, b = int(mine[s]), int(table[s])
        assert int(npm._slots(npm.neural_points[a:a + 1])) == s == int(npm._slots(npm.neural_points[b:b + 1]))
    assert differ.numel() <= 0.02 * int((table >= 0).sum())


def test_dropin_loop_closure_map_adjustment_matches_reference():
    """SURVEY.md section 8 row f4 (host side): adjust_map (per-frame pose corrections incl. the quaternion update) and
    recreate_hash in both modes reproduce the reference (model/neural_points.py:791-908) on its fixture."""
    fx = load_npz("loop_kitti")
    npm = _dropin_from_fixture(fx)
    npm.config.use_mid_ts = bool(fx["cfg.use_mid_ts"])
    pos, cur_ts = t(fx["sensor_pos"]), int(fx["map.cur_ts"])
    npm.adjust_map(t(fx["pose_diff"]))
    assert npm.after_pgo
    np.testing.assert_allclose(npm.neural_points.numpy(), fx["adjusted.neural_points"], rtol=1e-6, atol=1e-6)
    q, qr = npm.point_orientations.numpy(), fx["adjusted.point_orientations"]
    sign = np.sign((q * qr).sum(1, keepdims=True))  # q and -q are the same rotation
    np.testing.assert_allclose(q * sign, qr, rtol=1e-5, atol=1e-6)
    import copy

    kept = copy.deepcopy(npm)
    kept.recreate_hash(pos, torch.eye(3), kept_points=True, with_ts=True, cur_ts=cur_ts)
    _assert_same_hash_table(kept, fx["rehash_kept.table_slots"], fx["rehash_kept.table_vals"])
    assert np.array_equal(kept.local_mask.numpy(), fx["rehash_kept.local_mask"])
    filt = copy.deepcopy(npm)
    filt.recreate_hash(pos, torch.eye(3), kept_points=False, with_ts=True, cur_ts=cur_ts)
    np.testing.assert_allclose(filt.neural_points.numpy(), fx["rehash_filter.neural_points"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(filt.geo_features.numpy(), fx["rehash_filter.geo_features"])
    assert np.array_equal(filt.point_ts_create.numpy(), fx["rehash_filter.point_ts_create"])
    _assert_same_hash_table(filt, fx["rehash_filter.table_slots"], fx["rehash_filter.table_vals"])


def test_dropin_map_growth_matches_reference():
    """SURVEY.md section 8 row f1 (host side): three frames of NeuralPoints.update on the drop-in grow the same map as
    the reference (model/neural_points.py:311-422): same points in the same order, same timestamps, same hash table,
    the same randn feature initialisation (RNG stream), and the same local map after the built-in reset."""
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import NeuralPoints

    fx = load_npz("growth_kitti")
    g = lambda k: fx["map." + k]  # noqa: E731
    cfg = HotPathConfig.kitti(device="cpu", buffer_size=int(g("buffer_size")), feature_std=float(fx["cfg.feature_std"]),
                              local_map_radius=float(fx["cfg.local_map_radius"]))
    torch.manual_seed(int(fx["seed"]))
    npm = NeuralPoints(cfg)
    npm.diff_travel_dist_local = float(g("diff_travel_dist_local"))
    npm.travel_dist = t(g("travel_dist"))
    for f in range(int(fx["n_frames"])):
        npm.update(t(fx[f"frame{f}.points"]), t(fx[f"frame{f}.pos"]), torch.eye(3), f)
        assert npm.count() == int(fx[f"frame{f}.count"])
    assert np.array_equal(npm.neural_points.numpy(), g("neural_points"))
    assert np.array_equal(npm.point_ts_create.numpy(), g("point_ts_create"))
    assert np.array_equal(npm.point_ts_update.numpy(), g("point_ts_update"))
    assert np.array_equal(npm.geo_features.numpy(), g("geo_features"))
    table = torch.full((npm.buffer_size,), -1, dtype=torch.int64)
    table[t(g("table_slots"))] = t(g("table_vals")).long()
    assert torch.equal(npm.buffer_pt_index.long(), table)
    assert np.array_equal(npm.local_mask.numpy(), g("local_mask"))
    assert np.array_equal(npm.global2local.numpy().astype(np.int64), g("global2local").astype(np.int64))
    assert np.array_equal(npm.local_geo_features.data.numpy(), g("local_geo_features"))


def test_dropin_process_frame_matches_reference():
    """SURVEY.md section 8 row f2 (host side): three frames of the drop-in Mapper.process_frame -- per-ray sampling,
    map growth, replay-pool append + window filter, new-sample selection -- reproduce the reference
    (utils/mapper.py:162-449).  The one CUDA call on this path (query_certainty) is served by the oracle here, so the
    host logic runs on the CPU."""
    import types

    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import Decoder, NeuralPoints
    from pin_slam_b200.utils.mapper import Mapper

    fx = load_npz("frames_kitti")
    n_frames = int(fx["n_frames"])
    cfg = HotPathConfig.kitti(device="cpu", buffer_size=2000003, local_map_radius=float(fx["cfg.local_map_radius"]),
                              pool_filter_freq=int(fx["cfg.pool_filter_freq"]), adaptive_iters=True)
    cfg.window_radius = float(fx["cfg.window_radius"])
    torch.manual_seed(int(fx["seed"]))
    npm = NeuralPoints(cfg)
    npm.diff_travel_dist_local = 4.5
    npm.travel_dist = t(fx["travel_dist"])
    dec = Decoder(cfg, 64, 1, 1)
    poses = fx["poses"]
    dataset = types.SimpleNamespace(processed_frame=0, lose_track=False, stop_status=False, gt_pose_provided=False,
                                    odom_poses=poses.copy(), pgo_poses=None, gt_poses=None, static_mask=None)
    mapper = Mapper(cfg, dataset, npm, {"sdf": dec, "semantic": None, "color": None})

    def oracle_certainty(q):  # stands in for pinb200_query_certainty (CUDA) on this CPU-only run
        n = npm.count()
        m = po.OracleMap(resolution=npm.resolution, buffer_size=npm.buffer_size, feature_dim=cfg.feature_dim,
                         neural_points=npm.neural_points, point_orientations=npm.point_orientations,
                         geo_features=npm.geo_features, color_features=None, point_ts_create=npm.point_ts_create,
                         point_ts_update=npm.point_ts_update, point_certainties=npm.point_certainties,
                         buffer_pt_index=npm.buffer_pt_index.long(), local_neural_points=npm.local_neural_points,
                         local_point_orientations=npm.local_point_orientations,
                         local_geo_features=npm.local_geo_features.data, local_color_features=None,
                         local_point_certainties=npm.local_point_certainties,
                         local_point_ts_update=npm.local_point_ts_update, local_mask=npm.local_mask,
                         global2local=npm.global2local.long(), neighbor_dx=npm.neighbor_dx,
                         max_valid_dist2=npm.max_valid_dist2, travel_dist=npm.travel_dist, cur_ts=npm.cur_ts,
                         diff_travel_dist_local=npm.diff_travel_dist_local, temporal_local_map_on=True,
                         after_pgo=False)
        assert m.neural_points.shape[0] == n
        return po.query_certainty(m, q)

    npm.query_certainty = oracle_certainty
    for f in range(n_frames):
        dataset.processed_frame = f
        torch.manual_seed(int(fx["seed"]) * 100 + f)
        mapper.process_frame(t(fx[f"frame{f}.points"]), None, torch.tensor(poses[f], dtype=torch.float64), f)
        assert npm.count() == int(fx[f"frame{f}.map_count"])
        assert mapper.pool_sample_count == int(fx[f"frame{f}.pool_sample_count"])
        assert mapper.cur_sample_count == int(fx[f"frame{f}.cur_sample_count"])
        assert np.array_equal(mapper.new_idx.numpy(), fx[f"frame{f}.new_idx"])
        assert mapper.adaptive_iter_offset == int(fx[f"frame{f}.adaptive_iter_offset"])
    np.testing.assert_allclose(mapper.coord_pool.numpy(), fx["pool.coord"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(mapper.global_coord_pool.numpy(), fx["pool.global_coord"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(mapper.sdf_label_pool.numpy(), fx["pool.sdf_label"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mapper.weight_pool.numpy(), fx["pool.weight"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(mapper.time_pool.numpy(), fx["pool.time"])


def test_dropin_tracking_loop_matches_reference():
    """The host-side convergence loop of the drop-in Tracker.tracking (iteration count, residual / valid-point checks,
    termination thresholds, final pose) against the reference's Tracker.tracking (utils/tracker.py:43-225) on a
    briefly trained map.  The per-iteration CUDA call (pinb200_track_iterations: K1 + K4) is served by the oracle's
    query + registration step here, in the result layout of include/pinb200.h, so the loop runs on the CPU."""
    import types

    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.utils.tracker import Tracker

    fx = load_npz("track_kitti")
    m = map_from_fixture(fx)
    dec = decoder_from_fixture(fx, "sdf_mlp")
    k, wf = int(fx["cfg.query_nn_k"]), bool(fx["cfg.weighted_first"])
    (min_g, max_g, gm_d, gm_g, lm, term_deg, term_m, surf_range, final_ratio, std_ratio,
     eig_thre) = (float(v) for v in fx["cfg.reg_floats"])
    iter_n, min_nn = (int(v) for v in fx["cfg.reg_ints"])
    cfg = HotPathConfig.kitti(device="cpu")
    cfg.reg_min_grad_norm, cfg.reg_max_grad_norm, cfg.reg_GM_dist_m, cfg.reg_GM_grad = min_g, max_g, gm_d, gm_g
    cfg.reg_lm_lambda, cfg.reg_term_thre_deg, cfg.reg_term_thre_m = lm, term_deg, term_m
    cfg.surface_sample_range_m, cfg.final_residual_ratio_thre, cfg.max_sdf_std_ratio = surf_range, final_ratio, std_ratio
    cfg.eigenvalue_ratio_thre, cfg.reg_iter_n, cfg.track_mask_query_nn_k = eig_thre, iter_n, min_nn
    tracker = Tracker(cfg, types.SimpleNamespace(), {"sdf": None, "semantic": None, "color": None})
    log = []

    def oracle_iterate(src, t_dev, n_iter, normals, sdf_label, colors, cdec, cgrad, cmode):
        assert n_iter == 1 and cmode == 0
        xyz = po.transform_points(src, t_dev)
        q = po.query_sdf(m, dec, xyz, k, wf, query_locally=True, need_grad=True)
        r = po.registration_step(xyz, q["sdf"], q["grad"], q["sdf_std"], q["nn_count"],
                                 torch.zeros(src.shape[0]) if sdf_label is None else sdf_label, min_nn, min_g, max_g,
                                 surf_range * std_ratio, gm_d if gm_d > 0 else None, gm_g if gm_g > 0 else None, lm,
                                 normals=normals)
        res = torch.zeros(32, dtype=torch.float64)  # layout of pinb200_gn_step's `result`
        res[:16] = r["T"].reshape(-1)
        res[16] = r["valid_count"]
        res[17] = r["residual_cm"]
        t_dev.copy_(r["T"] @ t_dev)  # T <- dT @ T on the "device"
        log.append((float(res[17]), int(res[16])))
        return None, res, torch.zeros(64, dtype=torch.float64)

    tracker._iterate = oracle_iterate
    T, cov, _, valid = tracker.tracking(t(fx["source"]), t(fx["init_pose"]).double(), cur_ts=1)
    assert valid == bool(fx["result.valid"])
    assert len(log) == int(fx["result.n_iter"])
    np.testing.assert_allclose([x[0] for x in log], fx["result.residual_cm"], rtol=2e-3, atol=2e-3)
    assert np.abs(np.array([x[1] for x in log]) - fx["result.valid_count"]).max() <= 2
    np.testing.assert_allclose(T.numpy(), fx["result.T"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("name", QUERY_FIXTURES)
@pytest.mark.parametrize("local", [True, False])
def test_dropin_query_feature_assembly_matches_reference(name, local, monkeypatch):
    """The reference-signature hot call of the drop-in (NeuralPoints.query_feature, differentiable torch assembly on top
    of the kNN ids) against the reference's outputs.  The kNN search itself (pinb200_knn_search on the GPU) is served
    by the oracle here, so the assembly -- global->local quirk, neighbour vectors, IDW weights, certainty,
    weighted-first reduction -- is checked on the CPU."""
    from pin_slam_b200 import ops

    fx = load_npz(name)
    m = map_from_fixture(fx)
    color = "map.color_features" in fx
    npm = _dropin_from_fixture(fx, color=color)
    npm.config.weighted_first = bool(fx["cfg.weighted_first"])
    npm.config.query_nn_k = int(fx["cfg.query_nn_k"])
    npm.after_pgo = bool(fx["map.after_pgo"])
    npm.reset_local_map(t(fx["sensor_pos"]), torch.eye(3), int(fx["map.cur_ts"]))
    npm.map_handle = lambda query_locally=True: ("oracle", bool(query_locally))

    def oracle_knn(handle, q, k, want_gidx=False):
        loc = handle[1]
        d2, gidx = po.radius_search(m, q, time_filtering=m.temporal_local_map_on and loc)
        idx = m.global2local[gidx] if loc else gidx.clone()
        cnt = (idx >= 0).sum(-1)
        d2 = d2.clone()
        d2[idx == -1] = 9e3
        sd, order = torch.sort(d2, dim=1)
        idx, gidx = idx.gather(1, order)[:, :k], gidx.gather(1, order)[:, :k]
        gidx = torch.where(idx >= 0, gidx, torch.full_like(gidx, -1))
        return idx.int(), sd[:, :k], None, cnt.int(), gidx.int()

    monkeypatch.setattr(ops, "knn_search", oracle_knn)
    q = t(fx["q"])
    geo, _, w, cnt, cert = npm.query_feature(q, training_mode=False, query_locally=local)
    tag = "qf_local" if local else "qf_global"
    assert np.array_equal(cnt.numpy(), fx[tag + ".nn_counts"])
    np.testing.assert_allclose(w.numpy(), fx[tag + ".weight"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(geo.detach().numpy(), fx[tag + ".geo"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(cert.numpy(), fx[tag + ".certainty"], rtol=1e-5, atol=1e-6)


def test_map_arena_adoption_and_pickling():
    """Host logic of the growth arenas (pin_slam_b200/model/neural_points.py: _MapArena): the public map tensors become
    views of fixed-capacity buffers, tensors assigned from outside are adopted on the next reserve, and pickling a map
    whose tensors are arena views stores the rows, not the capacity."""
    import pickle

    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import NeuralPoints
    from pin_slam_b200.model.neural_points import _MapArena

    cfg = HotPathConfig.kitti(device="cpu", buffer_size=1009)
    npm = NeuralPoints(cfg)
    g = torch.Generator().manual_seed(0)
    n = 37
    npm.neural_points = torch.randn(n, 3, generator=g)
    npm.point_orientations = torch.randn(n, 4, generator=g)
    npm.point_ts_create = torch.arange(n, dtype=torch.int32)
    npm.point_ts_update = torch.arange(n, dtype=torch.int32) + 1
    npm.point_certainties = torch.rand(n, generator=g)
    npm.geo_features = torch.randn(n + 1, cfg.feature_dim, generator=g)
    before = {k: getattr(npm, k).clone() for k in _MapArena.NAMES + ("geo_features",)}
    arena = _MapArena()
    assert not arena.owns(npm)
    arena.reserve(npm, n + 100)
    assert arena.owns(npm) and arena.cap >= n + 100
    for k, v in before.items():
        assert torch.equal(getattr(npm, k), v), k
        assert getattr(npm, k).untyped_storage().nbytes() > v.numel() * v.element_size()  # a view of the arena
    # in-place growth: rows behind the count are written through the buffers, bind() extends the views
    arena.buf["neural_points"][n:n + 2] = 7.0
    arena.bind(npm, n + 2)
    assert npm.count() == n + 2 and float(npm.neural_points[-1, 0]) == 7.0 and npm.geo_features.shape[0] == n + 3
    cap = arena.cap
    arena.reserve(npm, n + 50)  # fits: no reallocation
    assert arena.cap == cap and arena.owns(npm)
    # a tensor assigned from outside (pruning, rehash, loop closure) is adopted
    npm.neural_points = npm.neural_points.clone()
    assert not arena.owns(npm)
    arena.reserve(npm, n + 50)
    assert arena.owns(npm) and npm.count() == n + 2
    # pickling keeps the rows only
    npm.__dict__["_arena"] = arena
    blob = pickle.dumps(npm)
    assert len(blob) < 200_000
    back = pickle.loads(blob)
    assert torch.equal(back.neural_points, npm.neural_points) and "_arena" not in back.__dict__


def test_forward_mode_seeds_equal_autograd():
    """The closed-form derivatives of the IDW weights / of the interpolated neighbour vector that search_kernel hands
    to the tangent rows of the decode (oracle.idw_tangent_seeds) against torch.autograd in float64, incl. invalid
    neighbours and nearly coincident ones; and the identity the forward-mode path rests on:
    d (sum_k w_k f_k) / d q_j = sum_k omega_kj (f_k - f_0)."""
    g = torch.Generator().manual_seed(3)
    n, k, f = 64, 8, 5
    q = torch.randn(n, 3, generator=g, dtype=torch.float64)
    nb = q.unsqueeze(1) + 0.3 * torch.randn(n, k, 3, generator=g, dtype=torch.float64)
    nb[:4, 1] = nb[:4, 0] + 1e-9  # coincident neighbours
    valid = torch.rand(n, k, generator=g) > 0.2
    valid[:, 0] = True
    feats = torch.randn(n, k, f, generator=g, dtype=torch.float64)
    w, omega, P = po.idw_tangent_seeds(q, nb, valid)

    def weights(qq):
        d2 = ((qq.unsqueeze(1) - nb) ** 2).sum(-1)
        u = (1.0 / (d2 + 1e-15)) * valid
        return u / u.sum(1, keepdim=True)

    qq = q.clone().requires_grad_(True)
    ww = weights(qq)
    assert torch.allclose(ww.detach(), w, rtol=1e-12, atol=1e-14)
    for kk in range(k):
        gk = torch.autograd.grad(ww[:, kk].sum(), qq, retain_graph=True)[0]
        assert torch.allclose(gk, omega[:, kk], rtol=1e-7, atol=1e-9 * float(omega.abs().max()))
    xn = (ww.unsqueeze(-1) * (qq.unsqueeze(1) - nb)).sum(1)  # [n,3]
    for i in range(3):
        gi = torch.autograd.grad(xn[:, i].sum(), qq, retain_graph=True)[0]  # d xn_i / d q_j
        assert torch.allclose(gi, P[:, :, i], rtol=1e-7, atol=1e-8 * float(P.abs().max()))
    xf = (ww.unsqueeze(-1) * feats).sum(1)  # [n,f]
    T = torch.einsum("nkj,nkf->njf", omega, feats - feats[:, :1])
    for c in range(f):
        gc = torch.autograd.grad(xf[:, c].sum(), qq, retain_graph=True)[0]
        assert torch.allclose(gc, T[:, :, c], rtol=1e-6, atol=1e-8 * float(T.abs().max()))
