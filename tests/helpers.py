"""Shared test helpers: golden-fixture loading and seeded synthetic maps."""
import os

import numpy as np
import torch

from oracle import pin_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARITY_SLACK = []  # filled by tests/test_cuda_parity.py::assert_rel_close, reported by tests/conftest.py


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def t(x, dtype=None):
    y = torch.from_numpy(np.ascontiguousarray(x))
    return y if dtype is None else y.to(dtype)


def map_from_fixture(fx, device="cpu") -> po.OracleMap:
    g = lambda k: fx["map." + k]  # noqa: E731
    bsz = int(g("buffer_size"))
    table = torch.full((bsz,), -1, dtype=torch.int64)
    table[t(g("table_slots"))] = t(g("table_vals"))
    m = po.OracleMap(
        resolution=float(g("resolution")),
        buffer_size=bsz,
        feature_dim=int(g("geo_features").shape[1]),
        neural_points=t(g("neural_points")),
        point_orientations=t(g("point_orientations")),
        geo_features=t(g("geo_features")),
        color_features=t(g("color_features")) if "map.color_features" in fx else None,
        point_ts_create=t(g("point_ts_create")),
        point_ts_update=t(g("point_ts_update")),
        point_certainties=t(g("point_certainties")),
        buffer_pt_index=table,
        local_neural_points=t(g("local_neural_points")),
        local_point_orientations=t(g("local_point_orientations")),
        local_geo_features=t(g("local_geo_features")),
        local_color_features=t(g("local_color_features")) if "map.local_color_features" in fx else None,
        local_point_certainties=t(g("local_point_certainties")),
        local_point_ts_update=t(g("local_point_ts_update")),
        local_mask=t(g("local_mask")),
        global2local=t(g("global2local")),
        neighbor_dx=t(g("neighbor_dx")),
        max_valid_dist2=float(g("max_valid_dist2")),
        travel_dist=t(g("travel_dist")),
        cur_ts=int(g("cur_ts")),
        diff_travel_dist_local=float(g("diff_travel_dist_local")),
        temporal_local_map_on=bool(g("temporal_local_map_on")),
        after_pgo=bool(g("after_pgo")),
    )
    return m.to(device)


def decoder_from_fixture(fx, name, device="cpu") -> po.DecoderParams:
    hidden = []
    i = 0
    while f"{name}.layers.{i}.weight" in fx:
        hidden.append((t(fx[f"{name}.layers.{i}.weight"]), t(fx[f"{name}.layers.{i}.bias"])))
        i += 1
    out = (t(fx[f"{name}.lout.weight"]), t(fx[f"{name}.lout.bias"]))
    return po.DecoderParams(hidden, out, float(fx[f"{name}.sdf_scale"])).to(device)


def synthetic_surface_points(n, seed, extent=20.0):
    """Random points on a floor, walls and a few boxes of a closed room (metres)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda k: torch.rand(k, generator=g)  # noqa: E731
    k = n // 5
    lo, hi = -extent / 2, extent / 2
    u = lambda k: r(k) * (hi - lo) + lo  # noqa: E731
    parts = [
        torch.stack([u(k), u(k), torch.zeros(k)], 1),
        torch.stack([u(k), torch.full((k,), hi), r(k) * 4], 1),
        torch.stack([torch.full((k,), lo), u(k), r(k) * 4], 1),
        torch.stack([u(k), torch.full((k,), lo), r(k) * 4], 1),
        torch.stack([r(n - 4 * k) * 3 + 2, r(n - 4 * k) * 3 - 5, torch.full((n - 4 * k,), 1.5)], 1),
    ]
    p = torch.cat(parts, 0)
    return p + 0.01 * torch.randn(p.shape, generator=g)


def synthetic_map(n_surface=20000, seed=0, resolution=0.4, buffer_size=int(5e7), feature_dim=8,
                  color=False, extent=20.0, n_ts=3, local_radius=None, diff_td=1e9, after_pgo=False,
                  num_nei_cells=2, search_alpha=0.2):
    """A seeded OracleMap built with oracle.build_map (one neural point per voxel)."""
    g = torch.Generator().manual_seed(seed + 17)
    pts = synthetic_surface_points(n_surface, seed, extent)
    cells = po.cell_of(pts, resolution)
    _, first = np.unique(cells.numpy(), axis=0, return_index=True)
    first = torch.from_numpy(np.sort(first))
    sp = pts[first].contiguous()
    mg = sp.shape[0]
    geo = 0.1 * torch.randn(mg + 1, feature_dim, generator=g)
    col = 0.1 * torch.randn(mg + 1, feature_dim, generator=g) if color else None
    ts = torch.randint(0, n_ts, (mg,), generator=g).to(torch.int32)
    cert = torch.rand(mg, generator=g) * 3
    ori = None
    if after_pgo:
        qn = torch.randn(mg, 4, generator=g)
        ori = qn / qn.norm(dim=1, keepdim=True)
    m = po.build_map(sp, resolution, buffer_size, geo, col, ts, cert, ori, num_nei_cells, search_alpha)
    m.after_pgo = after_pgo
    m.travel_dist = torch.arange(n_ts + 1, dtype=torch.float32) * 2.0
    m.diff_travel_dist_local = diff_td
    sensor = torch.tensor([0.0, 0.0, 1.0])
    po.reset_local_map(m, sensor, local_radius if local_radius is not None else 1e6, n_ts - 1)
    return m


def queries_near(m, n, seed, sigma=0.15):
    g = torch.Generator().manual_seed(seed)
    sel = torch.randint(0, m.neural_points.shape[0], (n,), generator=g)
    return (m.neural_points[sel] + sigma * torch.randn(n, 3, generator=g)).contiguous()


# ---------------------------------------------------------------------------
# CUDA-side helpers (only used by -m gpu tests, smoke and bench)
# ---------------------------------------------------------------------------
def map_handle_from_oracle(m, query_locally=True, device="cuda", color=True):
    """Upload an OracleMap's tensors and wrap them in a pin_slam_b200.ops.MapHandle."""
    from pin_slam_b200 import ops

    dev = torch.device(device)
    up = lambda x, dt=None: None if x is None else (x if dt is None else x.to(dt)).contiguous().to(dev)  # noqa: E731
    local = query_locally
    return ops.MapHandle(
        slot_table=up(m.buffer_pt_index, torch.int32),
        buffer_size=m.buffer_size,
        points=up(m.neural_points),
        ts_create=up(m.point_ts_create, torch.int32),
        travel_dist=up(m.travel_dist),
        global2local=up(m.global2local, torch.int32) if local else None,
        nb_points=up(m.local_neural_points if local else m.neural_points),
        nb_orient=up(m.local_point_orientations if local else m.point_orientations),
        geo_feat=up((m.local_geo_features if local else m.geo_features).detach()),
        color_feat=up((m.local_color_features if local else m.color_features).detach())
        if (color and m.color_features is not None) else None,
        certainty=up(m.local_point_certainties if local else m.point_certainties),
        ts_update=up(m.local_point_ts_update if local else m.point_ts_update, torch.int32),
        probe_dx=up(m.neighbor_dx, torch.int32),
        resolution=m.resolution,
        max_valid_dist2=m.max_valid_dist2,
        time_filter=m.temporal_local_map_on and local,
        cur_ts=m.cur_ts,
        diff_travel_dist_local=m.diff_travel_dist_local,
        after_pgo=m.after_pgo,
    )


def decoder_handle_from_oracle(dec, device="cuda", sigmoid_out=False):
    from pin_slam_b200 import ops

    dev = torch.device(device)
    ws = [w.detach().contiguous().to(dev) for w, _ in dec.hidden]
    bs = [b.detach().contiguous().to(dev) for _, b in dec.hidden]
    return ops.DecoderHandle(ws, bs, dec.out[0].detach().contiguous().to(dev), dec.out[1].detach().contiguous().to(dev),
                             out_scale=1.0 if sigmoid_out else dec.sdf_scale, leaky=dec.leaky, sigmoid_out=sigmoid_out)


def flat_decoder_params(dec):
    """[w0|b0|w1|b1|...|w_out|b_out] -- the layout pinb200_train_backward / adam use."""
    parts = []
    for w, b in dec.hidden:
        parts += [w.detach().reshape(-1), b.detach().reshape(-1)]
    parts += [dec.out[0].detach().reshape(-1), dec.out[1].detach().reshape(-1)]
    return torch.cat(parts)
