#!/usr/bin/env python3
"""Benchmark of the B200-native PIN-SLAM hot path (contract: see the task statement / DESIGN.md §4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1: BASELINE.json configs[1] -- the fused kNN + SDF-MLP query (K1) over a batch of
200 000 query points, K=8 neighbours, 32-d features, 2x64 decoder, C=33 probe cells, with the
analytic d sdf / d x.  One *step* is one pass of the hot path over that batch.

  value     whole-job throughput, inputs resident in HBM, in ALGORITHMIC GB/s
            (bytes per query from SURVEY.md §8d: 12 + 4C + 16 N_occ + K_v (4F+4) + 28; N_occ and K_v
            are measured on the workload) -- per-step CUDA-event time on the launching stream, L2
            flushed between steps; max over ranks
  e2e       the same metric through the reference-facing call (NeuralPoints.query_sdf on HOST,
            pinned, query points; results copied back to the host) with the copies inside the timed
            region
  roofline  achieved algorithmic GB/s of the K1 kernel / measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
            the CPU oracle (a restatement of the reference's PyTorch path, oracle/pin_oracle.py) on
            the host cores over a bounded sample of the same batch
N>1: the query path has no exchange step -- every rank runs the same per-GPU batch on its own map
replica (weak scaling); no collective on the data path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_QUERY = 200_000
FALLBACK_HBM_GBS = 6650.0


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


MAP_SCALE = 1.0  # --map-scale: the synthetic map grows by scale^2 at constant point density (M = 250 k / 1 M runs)


def build_workload(device):
    import torch

    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import Decoder
    from pin_slam_b200.synthetic import build_map, surface_queries

    cfg = HotPathConfig.cfg2(device=str(device), feature_std=0.1, local_map_radius=1e4)
    npm = build_map(cfg, n_surface=int(3_000_000 * MAP_SCALE**2), seed=0, extent=80.0 * MAP_SCALE)
    torch.manual_seed(42)
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    q = surface_queries(npm, N_QUERY, seed=1, sigma=0.1)
    return cfg, npm, dec, q


def workload_stats(npm, q, k):
    """Measured N_occ (occupied probes / query) and K_v (selected valid neighbours / query)."""
    import torch

    from pin_slam_b200 import ops

    _, idx = ops.radius_search(npm.map_handle(False), q[:50000].contiguous())
    n_occ = float((idx >= 0).sum(1).float().mean())
    _, _, _, cnt = ops.knn_search(npm.map_handle(True), q, k)
    k_v = float(torch.clamp(cnt, max=k).float().mean())
    return n_occ, k_v, float((cnt >= k).float().mean())


def bytes_per_query(c, n_occ, k_v, f):
    return 12 + 4 * c + 16 * n_occ + k_v * (4 * f + 4) + 28


def oracle_map_from(npm):
    """CPU OracleMap holding copies of the NeuralPoints state (baseline legs only)."""
    from oracle import pin_oracle as po

    c = lambda x: None if x is None else x.detach().cpu().clone()  # noqa: E731
    m = po.OracleMap(
        resolution=npm.resolution, buffer_size=npm.buffer_size, feature_dim=npm.geo_feature_dim,
        neural_points=c(npm.neural_points), point_orientations=c(npm.point_orientations),
        geo_features=c(npm.geo_features), color_features=c(npm.color_features),
        point_ts_create=c(npm.point_ts_create), point_ts_update=c(npm.point_ts_update),
        point_certainties=c(npm.point_certainties), buffer_pt_index=c(npm.buffer_pt_index).long(),
        local_neural_points=c(npm.local_neural_points), local_point_orientations=c(npm.local_point_orientations),
        local_geo_features=c(npm.local_geo_features), local_color_features=None,
        local_point_certainties=c(npm.local_point_certainties), local_point_ts_update=c(npm.local_point_ts_update),
        local_mask=c(npm.local_mask), global2local=c(npm.global2local).long(), neighbor_dx=c(npm.neighbor_dx),
        max_valid_dist2=npm.max_valid_dist2, travel_dist=c(npm.travel_dist), cur_ts=npm.cur_ts,
        diff_travel_dist_local=npm.diff_travel_dist_local, temporal_local_map_on=npm.temporal_local_map_on,
        after_pgo=npm.after_pgo)
    return m


def oracle_decoder_from(dec):
    from oracle import pin_oracle as po

    hidden = [(l.weight.detach().cpu().clone(), l.bias.detach().cpu().clone()) for l in dec.layers]
    return po.DecoderParams(hidden, (dec.lout.weight.detach().cpu().clone(), dec.lout.bias.detach().cpu().clone()),
                            dec.sdf_scale)


def time_cpu_oracle(m, d, q_cpu, k, wf, sample, repeats=1):
    """Seconds per `sample` queries of the reference algorithm (oracle port) on the host cores."""
    import torch

    from oracle import pin_oracle as po

    qs = q_cpu[:sample]
    po.query_sdf(m, d, qs[: min(2000, sample)], k, wf)  # warm the thread pool / allocator
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        po.query_sdf(m, d, qs, k, wf)
        best = min(best, time.perf_counter() - t0)
    return best


def calibrate_cpu_threads(m, d, q_cpu, k, wf):
    """The thread count at which the CPU path is fastest on this host.  The path is ~200 small ATen ops per call;
    with one thread per core on a 128-core host the OpenMP fork/join cost dominates them (measured: 25x slower
    than 16 threads), so `all cores` would understate the reference.  Returns (threads, {threads: seconds})."""
    import torch

    from oracle import pin_oracle as po

    ncpu = os.cpu_count() or 1
    cand = sorted({t for t in (4, 8, 16, 32, 64, ncpu) if t <= ncpu})
    probe = q_cpu[:4000]
    timings = {}
    for t in cand:
        torch.set_num_threads(t)
        po.query_sdf(m, d, probe[:1000], k, wf)  # warm the pool at this size
        t0 = time.perf_counter()
        po.query_sdf(m, d, probe, k, wf)
        timings[t] = time.perf_counter() - t0
    best = min(timings, key=timings.get)
    torch.set_num_threads(best)
    return best, timings


def run_reference(args):
    """--impl reference: the reference's own algorithm for this path on the host CPU (the reference is
    pure Python/PyTorch and /root/reference does not exist on the GPU box, so the oracle port runs it)."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    cfg, npm, dec, q = build_workload(dev)
    from pin_slam_b200 import ops  # stats only (C, N_occ, K_v of the workload)

    n_occ, k_v, _ = workload_stats(npm, q, cfg.query_nn_k) if dev.type == "cuda" else (11.0, 8.0, 1.0)
    bq = bytes_per_query(npm.neighbor_K, n_occ, k_v, cfg.feature_dim)
    m, d = oracle_map_from(npm), oracle_decoder_from(dec)
    qc = q.cpu()
    sample = 20000
    threads, tried = calibrate_cpu_threads(m, d, qc, cfg.query_nn_k, cfg.weighted_first)
    for _ in range(args.warmup):
        time_cpu_oracle(m, d, qc, cfg.query_nn_k, cfg.weighted_first, 2000)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        from oracle import pin_oracle as po

        po.query_sdf(m, d, qc[:sample], cfg.query_nn_k, cfg.weighted_first)
    dt = (time.perf_counter() - t0) / args.steps
    val = bq * sample / dt / 1e9
    line = {
        "impl": "reference", "metric": "kNN+MLP fused query throughput (algorithmic bytes)", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, npm, n_occ, k_v, bq, sample=sample),
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port",
                         "host_cores": os.cpu_count(),
                         "threads_tried_s_per_4000_queries": {str(t): round(v, 4) for t, v in tried.items()},
                         "sample": f"{sample} of the {N_QUERY} queries per step, torch CPU ops, at the fastest "
                                   "thread count of the ones tried"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "queries_per_s": sample / dt,
    }
    print(json.dumps(line))


def workload_config(cfg, npm, n_occ, k_v, bq, sample=None):
    return {
        "workload": "BASELINE configs[1]: fused kNN+SDF-MLP query, 200k query pts, K=8, F=32, 2x64 decoder, "
                    "C=33 probes, with d sdf/dx",
        "n_query": N_QUERY if sample is None else sample, "nn_k": cfg.query_nn_k, "feature_dim": cfg.feature_dim,
        "decoder": f"{cfg.geo_mlp_level}x{cfg.geo_mlp_hidden_dim}", "weighted_first": cfg.weighted_first,
        "n_probe": int(npm.neighbor_K), "map_points": int(npm.count()), "local_points": int(npm.local_count()),
        "buffer_size": int(npm.buffer_size), "occupied_probes_mean": round(n_occ, 3), "valid_knn_mean": round(k_v, 3),
        "bytes_per_query": round(bq, 1), "l2": "flushed between timed steps (512 MiB memset)",
        "parallelism": "replicas (no collective on the query path)",
    }


def frame_benchmark(dev, n_frames=12, warm_iters=100):
    """BASELINE configs[2]: per-frame tracker (3 GN iterations) + mapper (5 training iterations) on synthetic
    64x1024 KITTI-shaped scans; frames/s over tracker+mapper only.  Also times the reference's op sequence
    (oracle port, PyTorch eager) on the same GPU and map state for the same two stages."""
    import torch

    from pin_slam_b200 import ops
    from pin_slam_b200.frame_loop import FrameLoop

    loop = FrameLoop(device=dev)
    loop.step(0, timed=False, map_iters=warm_iters)   # frame 0 bootstraps the map (untimed, like the reference's init)
    loop.step(1, timed=False)
    l0 = ops.launch_count()
    info = [loop.step(f) for f in range(2, 2 + n_frames)]
    launches = (ops.launch_count() - l0) / n_frames
    trk = sorted(t for t, _ in loop.times)[len(loop.times) // 2]
    mp = sorted(m for _, m in loop.times)[len(loop.times) // 2]
    out = {"workload": "BASELINE configs[2]: tracker GN x3 + mapper x5 per frame, 64x1024 synthetic KITTI scan, "
                       "run_kitti.yaml parameters (F=8, K=6, 1x64, weighted_first=False, bs 16384)",
           "frames": n_frames, "tracker_ms_median": trk, "mapping_ms_median": mp,
           "frames_per_s": 1000.0 / (trk + mp), "kernel_launches_per_frame": launches,
           "source_points": info[-1]["n_source"], "scan_points": info[-1]["n_scan"],
           "local_map_points": info[-1]["local_points"], "pool_samples": info[-1]["pool"],
           "final_translation_error_m": info[-1]["trans_err_m"]}
    try:
        out["torch_eager_gpu_baseline"] = frame_baseline_torch(loop, dev)
        out["speedup_vs_torch_eager_gpu"] = out["frames_per_s"] / out["torch_eager_gpu_baseline"]["frames_per_s"]
    except Exception as e:  # noqa: BLE001
        out["torch_eager_gpu_baseline"] = {"error": str(e)[:200]}
    return out


def frame_baseline_torch(loop, dev, device=None, reps=3):
    """The reference's per-frame op sequence (oracle port) on `device` for the current map state:
    3 x (query_source_points + registration_step) and 5 x one Mapper.mapping iteration."""
    import torch

    from oracle import pin_oracle as po

    device = device or dev
    cfg, npm, mapper = loop.cfg, loop.neural_points, loop.mapper
    m = oracle_map_from(npm).to(device)
    d = oracle_decoder_from(loop.sdf_mlp).to(device)
    _, _, source = loop.preprocess(len(loop.poses))
    source = source.to(device)
    pose = loop.poses[-1].to(device)
    sync = (lambda: torch.cuda.synchronize()) if torch.device(device).type == "cuda" else (lambda: None)

    def track():
        T = pose.clone()
        for _ in range(3):
            pts = po.transform_points(source, T)
            o = po.query_sdf(m, d, pts, cfg.query_nn_k, cfg.weighted_first)
            r = po.registration_step(pts, o["sdf"], o["grad"], o["sdf_std"], o["nn_count"], torch.zeros_like(o["sdf"]),
                                     cfg.track_mask_query_nn_k, cfg.reg_min_grad_norm, cfg.reg_max_grad_norm,
                                     cfg.surface_sample_range_m * cfg.max_sdf_std_ratio, cfg.reg_GM_dist_m,
                                     cfg.reg_GM_grad, cfg.reg_lm_lambda)
            T = r["T"] @ T

    def train():
        mm, dd = m.clone(), d.clone()
        mm.local_geo_features.requires_grad_(True)
        dd.requires_grad_(True)
        opt = po.make_adam([dd.tensors(), [mm.local_geo_features]], cfg.lr, cfg.adam_eps, cfg.weight_decay)
        for _ in range(5):
            coord, label, ts, _, _, _, w = mapper.get_batch(global_coord=True)
            coord, label, ts, w = coord.to(device), label.to(device), ts.to(device), w.to(device)
            loss, _ = po.mapping_loss(mm, dd, coord, label, ts, w, cfg.query_nn_k, cfg.weighted_first, mapper.sdf_scale,
                                      cfg.loss_weight_on, cfg.weight_e, cfg.gradient_decimation,
                                      cfg.voxel_size_m * cfg.num_grad_step_ratio)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()

    track(); train(); sync()
    tt, tm = [], []
    for _ in range(reps):
        t0 = time.perf_counter(); track(); sync(); t1 = time.perf_counter(); train(); sync(); t2 = time.perf_counter()
        tt.append((t1 - t0) * 1e3); tm.append((t2 - t1) * 1e3)
    trk, mp = sorted(tt)[len(tt) // 2], sorted(tm)[len(tm) // 2]
    return {"tracker_ms": trk, "mapping_ms": mp, "frames_per_s": 1000.0 / (trk + mp), "device": str(device),
            "what": "reference op sequence (oracle port) in PyTorch eager: 3 x (query + GN step), 5 x training iteration"}


def mapper_benchmark(args):
    """BASELINE configs[4]: mapper-only data-parallel training.  A 2M-sample replay pool sharded over the ranks,
    per-GPU batch 16384 (weak scaling), K1 forward + loss heads + K2 backward + ONE NCCL all-reduce of
    [feature grads | decoder grads | certainty increments] + K3 Adam per iteration.  value = samples/s (all ranks)."""
    import types

    import torch
    import torch.distributed as dist

    from pin_slam_b200 import ops
    from pin_slam_b200.config import HotPathConfig
    from pin_slam_b200.model import Decoder
    from pin_slam_b200.synthetic import build_map, surface_queries
    from pin_slam_b200.utils.mapper import Mapper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = HotPathConfig.kitti(device=str(dev), feature_std=0.05, bs_new_sample=0, local_map_radius=1e4)
    npm = build_map(cfg, n_surface=2_000_000, seed=0, extent=80.0)   # identical replica on every rank (same seed)
    torch.manual_seed(42)
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    ds = types.SimpleNamespace(processed_frame=0, lose_track=False, stop_status=False, gt_pose_provided=False,
                               odom_poses=None, pgo_poses=None, gt_poses=None)
    mapper = Mapper(cfg, ds, npm, {"sdf": dec, "semantic": None, "color": None})
    pool_total = 2_000_000
    n = pool_total // world                                        # this rank's shard of the replay pool
    g = torch.Generator().manual_seed(100 + rank)
    coord = surface_queries(npm, n, seed=200 + rank, sigma=0.15)
    mapper.global_coord_pool = coord
    mapper.coord_pool = coord
    mapper.sdf_label_pool = (0.15 * torch.randn(n, generator=g)).to(dev)
    mapper.weight_pool = (torch.rand(n, generator=g) * 0.8 + 0.6).to(dev)
    mapper.time_pool = torch.zeros(n, dtype=torch.int32, device=dev)
    mapper.pool_sample_count = n
    torch.manual_seed(1000 + rank)                                  # per-rank batch draws

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mapper.mapping(max(args.warmup, 3))
    barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mapper.mapping(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": "mapper training throughput (samples/s, all GPUs)", "value": cfg.bs * world / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: mapper-only, 2M-sample pool sharded over ranks, bs/GPU 16384, "
                                   "run_kitti.yaml parameters, NCCL all-reduce of feature+decoder grads per iteration",
                       "local_points": int(npm.local_count()), "pool_per_rank": n,
                       "allreduce_floats": int(npm.local_geo_features.numel() + dec.flat_parameters().numel()
                                               + npm.local_count()),
                       "parallelism": f"dp{world}"},
            "gpu_launches": ops.launch_count() - l0}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frame", action="store_true", help="skip the per-frame tracker+mapper measurement")
    ap.add_argument("--map-scale", type=float, default=1.0,
                    help="query workload only: scale the synthetic map (points ~ scale^2, same density); "
                         "1.55 ~ 250 k points, 3.1 ~ 1 M points (map >> L2).  The default (1.0, 106 k points) is the "
                         "configuration every committed number refers to")
    ap.add_argument("--workload", default="query", choices=["query", "mapper"],
                    help="query = BASELINE configs[1] (default, the headline); mapper = configs[4] data-parallel training")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    global MAP_SCALE
    MAP_SCALE = float(args.map_scale)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "mapper":
        mapper_benchmark(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the hot path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from pin_slam_b200 import ops

    cfg, npm, dec, q = build_workload(dev)
    k = cfg.query_nn_k
    n_occ, k_v, full_k = workload_stats(npm, q, k)
    bq = bytes_per_query(npm.neighbor_K, n_occ, k_v, cfg.feature_dim)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out = {}

    def step():
        return npm.query_sdf(q, dec, need_grad=True, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- device-resident throughput: per-step CUDA events on the launching stream, L2 flushed in between
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.zero_()
        a.record()
        step()
        b.record()
    barrier()
    launches = ops.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public call with host buffers
    q_host = q.cpu().pin_memory()
    res_host = {n: torch.empty(s, dtype=d).pin_memory() for n, s, d in
                [("sdf", (N_QUERY,), torch.float32), ("grad", (N_QUERY, 3), torch.float32),
                 ("sdf_std", (N_QUERY,), torch.float32), ("nn_count", (N_QUERY,), torch.int32),
                 ("certainty", (N_QUERY,), torch.float32)]}
    def e2e_step():
        # the public host-facing call: pinned host queries in, pinned host results out; the batch is cut into 4
        # pieces on two streams so that the PCIe copies overlap the kernel (pin_slam_b200/model/neural_points.py)
        npm.query_sdf_host(q_host, dec, res_host, chunks=4, need_grad=True)

    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = bq * N_QUERY * world / (ms_per_step * 1e-3) / 1e9
    e2e_val = bq * N_QUERY * world / (e2e_ms / args.steps * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    kernel_ms = sorted(step_ms)[len(step_ms) // 2]  # one K1 launch per step: the step time IS the kernel time
    achieved = bq * N_QUERY / (ms_per_step * 1e-3) / 1e9

    line = None
    if rank == 0:
        line = {
            "metric": "kNN+MLP fused query throughput (algorithmic bytes)", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, npm, n_occ, k_v, bq),
            "queries_per_s": N_QUERY * world / (ms_per_step * 1e-3),
            "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": 12 * N_QUERY,
                    "d2h_bytes_per_step": 28 * N_QUERY, "ms_per_step": e2e_ms / args.steps,
                    "call": "NeuralPoints.query_sdf_host(pinned host queries) -> pinned host sdf/grad/std/nn_count/certainty, "
                            "4 pieces pipelined over 2 streams"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "pinb::query_kernel<64,36>",
                         "kernel_ms_median": kernel_ms,
                         "note": "achieved = algorithmic bytes/query x queries / mean CUDA-event time of the K1 launch"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            m, d = oracle_map_from(npm), oracle_decoder_from(dec)
            sample = 20000
            threads, tried = calibrate_cpu_threads(m, d, q.cpu(), k, cfg.weighted_first)
            sec = time_cpu_oracle(m, d, q.cpu(), k, cfg.weighted_first, sample, repeats=3)
            line["cpu_baseline"] = {"value": bq * sample / sec / 1e9, "unit": "GB/s", "cores": threads,
                                    "host_cores": os.cpu_count(), "kind": "port", "queries_per_s": sample / sec,
                                    "threads_tried_s_per_4000_queries": {str(t): round(v, 4) for t, v in tried.items()},
                                    "sample": f"{sample} of the {N_QUERY} queries, oracle (torch CPU restatement of "
                                              "the reference path) at the fastest thread count tried, best of 3"}
            # the reference's own GPU mode = the same PyTorch op sequence on the device (the >=10x denominator)
            try:
                mg, dg = m.clone().to(dev), d.to(dev)
                from oracle import pin_oracle as po

                po.query_sdf(mg, dg, q, k, cfg.weighted_first)  # warm-up at full size (allocator, cuBLAS init)
                torch.cuda.synchronize()
                sec_g = float("inf")
                for _ in range(3):
                    t0 = time.perf_counter()
                    po.query_sdf(mg, dg, q, k, cfg.weighted_first)
                    torch.cuda.synchronize()
                    sec_g = min(sec_g, time.perf_counter() - t0)
                line["torch_eager_gpu_baseline"] = {
                    "value": bq * N_QUERY / sec_g / 1e9, "unit": "GB/s", "ms_per_step": sec_g * 1e3,
                    "what": "reference op sequence (oracle port) in PyTorch eager on the same B200, 200k queries, best of 3"}
            except Exception as e:  # noqa: BLE001
                line["torch_eager_gpu_baseline"] = {"error": str(e)[:200]}
        if world == 1 and not args.no_frame:
            del flush
            torch.cuda.empty_cache()
            try:
                line["per_frame"] = frame_benchmark(dev)
            except Exception as e:  # noqa: BLE001
                line["per_frame"] = {"error": repr(e)[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
