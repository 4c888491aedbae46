"""Seeded synthetic inputs of the benchmark workloads (there is no network for datasets).

  * `room_surface_points`  -- points on the surfaces of a closed analytic scene (ground, four
    walls, a few boxes that break the corridor symmetry), used to grow a neural point map;
  * `lidar_scan`           -- a 64 x 1024 KITTI-shaped scan of that scene by analytic ray casting
    (64 elevations in [+2, -24.8] deg, 1024 azimuths, 1 cm range noise), sensor height 1.73 m;
  * `build_map`            -- grows a `NeuralPoints` map from surface points through its own
    `update()` (the reference's map-growth entry point).
"""
import math

import torch

def _make_boxes(n=28, seed=7):
    """Seeded clutter (cx, cy, half_x, half_y, height): boxes / pillars that break the corridor symmetry so
    that registration is constrained along the driving direction; the lane |y| < 4 m stays free."""
    import random

    rnd = random.Random(seed)
    boxes = [(8.0, 6.0, 2.0, 1.5, 2.5), (-12.0, -9.0, 1.5, 3.0, 3.0), (20.0, -14.0, 3.0, 2.0, 2.0),
             (-25.0, 15.0, 2.5, 2.5, 4.0), (3.0, -20.0, 1.0, 4.0, 1.5), (30.0, 22.0, 2.0, 2.0, 3.5)]
    while len(boxes) < n:
        cx, cy = rnd.uniform(-36, 36), rnd.uniform(-36, 36)
        hx, hy, h = rnd.uniform(0.3, 2.5), rnd.uniform(0.3, 2.5), rnd.uniform(1.0, 4.5)
        if abs(cy) - hy < 4.0:
            continue
        boxes.append((cx, cy, hx, hy, h))
    return boxes


BOXES = _make_boxes()


def room_surface_points(n, seed, extent=80.0, wall_h=6.0, device="cpu", noise=0.01):
    g = torch.Generator().manual_seed(seed)
    r = lambda k: torch.rand(k, generator=g)  # noqa: E731
    half = extent / 2
    u = lambda k: r(k) * extent - half  # noqa: E731
    area = [extent * extent] + [extent * wall_h] * 4 + [4 * (2 * b[2] + 2 * b[3]) * b[4] + 4 * b[2] * b[3] for b in BOXES]
    tot = sum(area)
    cnt = [max(1, int(n * a / tot)) for a in area]
    parts = [torch.stack([u(cnt[0]), u(cnt[0]), torch.zeros(cnt[0])], 1)]
    for i, (ax, val) in enumerate([(1, half), (1, -half), (0, half), (0, -half)]):
        k = cnt[1 + i]
        a, z = u(k), r(k) * wall_h
        fixed = torch.full((k,), val)
        parts.append(torch.stack([a, fixed, z], 1) if ax == 1 else torch.stack([fixed, a, z], 1))
    for j, (cx, cy, hx, hy, h) in enumerate(BOXES):
        k = cnt[5 + j]
        side = torch.randint(0, 5, (k,), generator=g)
        x = cx + (r(k) * 2 - 1) * hx
        y = cy + (r(k) * 2 - 1) * hy
        z = r(k) * h
        x = torch.where(side == 0, torch.full_like(x, cx - hx), torch.where(side == 1, torch.full_like(x, cx + hx), x))
        y = torch.where(side == 2, torch.full_like(y, cy - hy), torch.where(side == 3, torch.full_like(y, cy + hy), y))
        z = torch.where(side == 4, torch.full_like(z, h), z)
        parts.append(torch.stack([x, y, z], 1))
    p = torch.cat(parts, 0)
    p = p + noise * torch.randn(p.shape, generator=g)
    return p.to(device)


def _ray_box(o, d, lo, hi):
    """Slab test, rays [N,3] against one axis-aligned box; returns entry t (inf if missed)."""
    inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
    t0 = (lo - o) * inv
    t1 = (hi - o) * inv
    tmin = torch.minimum(t0, t1).max(dim=1)[0]
    tmax = torch.maximum(t0, t1).min(dim=1)[0]
    hit = (tmax >= tmin) & (tmax > 0)
    t = torch.where(tmin > 0, tmin, tmax)
    return torch.where(hit, t, torch.full_like(t, float("inf")))


def lidar_scan(pose: torch.Tensor, seed: int, extent=80.0, wall_h=6.0, beams=64, azimuths=1024, noise=0.01,
               max_range=80.0, device="cpu"):
    """Sensor-frame points [<=beams*azimuths, 3] of one scan taken at `pose` (4x4, sensor->world)."""
    g = torch.Generator().manual_seed(seed)
    el = torch.deg2rad(torch.linspace(2.0, -24.8, beams))
    az = torch.linspace(-math.pi, math.pi, azimuths + 1)[:-1]
    ee, aa = torch.meshgrid(el, az, indexing="ij")
    d_s = torch.stack([torch.cos(ee) * torch.cos(aa), torch.cos(ee) * torch.sin(aa), torch.sin(ee)], -1).reshape(-1, 3)
    rot, org = pose[:3, :3].float(), pose[:3, 3].float()
    d_w = d_s @ rot.T
    o = org.expand_as(d_w)
    half = extent / 2
    # inside of the room: distance to the exit of the big box = the wall/ground hit
    inv = 1.0 / torch.where(d_w.abs() < 1e-9, torch.full_like(d_w, 1e-9), d_w)
    lo = torch.tensor([-half, -half, 0.0])
    hi = torch.tensor([half, half, wall_h + 100.0])  # no ceiling: rays that go up leave the scene
    t_exit = torch.maximum((lo - o) * inv, (hi - o) * inv).min(dim=1)[0]
    hit_z = o[:, 2] + t_exit * d_w[:, 2]
    t = torch.where(hit_z <= wall_h + 1e-3, t_exit, torch.full_like(t_exit, float("inf")))
    for cx, cy, hx, hy, h in BOXES:
        t = torch.minimum(t, _ray_box(o, d_w, torch.tensor([cx - hx, cy - hy, 0.0]), torch.tensor([cx + hx, cy + hy, h])))
    ok = torch.isfinite(t) & (t < max_range) & (t > 1.0)
    rng = t + noise * torch.randn(t.shape, generator=g)
    return (d_s * rng.unsqueeze(1))[ok].to(device)


def trajectory_pose(frame: int, step=0.8, yaw_per_frame=0.004, height=1.73):
    yaw = yaw_per_frame * frame
    c, s = math.cos(yaw), math.sin(yaw)
    T = torch.eye(4, dtype=torch.float64)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = c, -s, s, c
    T[0, 3] = -20.0 + step * frame
    T[1, 3] = 0.5 * math.sin(0.05 * frame)
    T[2, 3] = height
    return T


def build_map(config, n_surface=3_000_000, seed=0, extent=80.0, feature_seed=1234):
    """A NeuralPoints map grown from synthetic surface points through NeuralPoints.update()."""
    from .model import NeuralPoints

    dev = torch.device(config.device)
    npm = NeuralPoints(config)
    npm.travel_dist = torch.zeros(4, device=dev)
    pts = room_surface_points(n_surface, seed, extent=extent, device=dev)
    gen_state = torch.cuda.get_rng_state(dev) if dev.type == "cuda" else torch.get_rng_state()
    torch.manual_seed(feature_seed)
    npm.update(pts, torch.tensor([0.0, 0.0, 1.0], device=dev), torch.eye(3, device=dev), 0)
    if dev.type == "cuda":
        torch.cuda.set_rng_state(gen_state, dev)
    else:
        torch.set_rng_state(gen_state)
    return npm


def surface_queries(npm, n, seed, sigma=0.1):
    """n query points: neural point positions + N(0, sigma^2) offsets."""
    dev = npm.neural_points.device
    g = torch.Generator().manual_seed(seed)
    sel = torch.randint(0, npm.count(), (n,), generator=g).to(dev)
    off = (sigma * torch.randn(n, 3, generator=g)).to(dev)
    return (npm.neural_points[sel] + off).contiguous()


# ---------------------------------------------------------------------------
# BASELINE configs[3]: Replica-shaped RGB-D frames (640 x 480 pinhole depth + colour) of an analytic room
# ---------------------------------------------------------------------------
ROOM = (6.0, 4.0, 3.0)  # x, y, z extent in metres, origin at the floor centre
FURNITURE = [(-1.8, -1.0, 0.6, 0.5, 0.8), (1.5, 1.2, 0.8, 0.4, 0.45), (0.3, -1.4, 0.4, 0.4, 1.2),
             (2.3, -0.9, 0.35, 0.6, 0.9), (-2.2, 1.3, 0.5, 0.45, 0.5)]  # (cx, cy, half_x, half_y, height)


def surface_colour(p: torch.Tensor) -> torch.Tensor:
    """Smooth analytic RGB texture in [0, 1] as a function of the world position."""
    f = torch.tensor([[1.3, 0.7, 2.1], [0.9, 1.9, 0.5], [2.3, 1.1, 1.7]], device=p.device, dtype=p.dtype)
    ph = torch.tensor([0.3, 1.1, 2.0], device=p.device, dtype=p.dtype)
    return 0.5 + 0.45 * torch.sin(p @ f.T + ph)


def rgbd_pose(frame: int):
    """Slow hand-held style trajectory inside the room: camera z forward, looking along +x, slightly down."""
    T = torch.eye(4, dtype=torch.float64)
    yaw = 0.02 * frame
    c, s = math.cos(yaw), math.sin(yaw)
    # camera axes in world: z_cam = forward (cos yaw, sin yaw, -0.15), x_cam = right, y_cam = down
    fwd = torch.tensor([c, s, -0.15], dtype=torch.float64)
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    T[:3, 0], T[:3, 1], T[:3, 2] = right, down, fwd
    T[0, 3] = -2.0 + 0.03 * frame
    T[1, 3] = 0.2 * math.sin(0.1 * frame)
    T[2, 3] = 1.4
    return T


def rgbd_frame(pose: torch.Tensor, seed: int, width=640, height=480, fx=525.0, fy=525.0, noise=0.002, max_range=10.0,
               device="cpu"):
    """Camera-frame points [<= H*W, 3] and colours [.., 3] of one RGB-D frame taken at `pose` (4x4, camera->world)."""
    dev = torch.device(device)
    g = torch.Generator().manual_seed(seed)
    u = torch.arange(width, device=dev, dtype=torch.float32) - (width - 1) / 2
    v = torch.arange(height, device=dev, dtype=torch.float32) - (height - 1) / 2
    vv, uu = torch.meshgrid(v, u, indexing="ij")
    d_c = torch.stack([uu / fx, vv / fy, torch.ones_like(uu)], -1).reshape(-1, 3)  # z = 1 plane: depth = t
    rot, org = pose[:3, :3].float().to(dev), pose[:3, 3].float().to(dev)
    d_w = d_c @ rot.T
    o = org.expand_as(d_w)
    hx, hy, hz = ROOM[0] / 2, ROOM[1] / 2, ROOM[2]
    inv = 1.0 / torch.where(d_w.abs() < 1e-9, torch.full_like(d_w, 1e-9), d_w)
    lo = torch.tensor([-hx, -hy, 0.0], device=dev)
    hi = torch.tensor([hx, hy, hz], device=dev)
    t = torch.maximum((lo - o) * inv, (hi - o) * inv).min(dim=1)[0]  # exit of the room box = wall / floor / ceiling
    for cx, cy, bx, by, bh in FURNITURE:
        t = torch.minimum(t, _ray_box(o, d_w, torch.tensor([cx - bx, cy - by, 0.0], device=dev),
                                      torch.tensor([cx + bx, cy + by, bh], device=dev)))
    ok = torch.isfinite(t) & (t < max_range) & (t > 0.05)
    depth = t + (noise * torch.randn(t.shape, generator=g)).to(dev)
    world = o + d_w * t.unsqueeze(1)
    return (d_c * depth.unsqueeze(1))[ok].contiguous(), surface_colour(world)[ok].contiguous()
