"""Debug helper: per-iteration tracker / mapper diagnostics on synthetic frames (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pin_slam_b200.frame_loop import FrameLoop
from pin_slam_b200.synthetic import trajectory_pose

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 100
loop = FrameLoop(device="cuda", n_track_iter=3, n_map_iter=5)
info = loop.step(0, map_iters=warm)
print("frame0", info, "losses", loop.mapper.last_losses.tolist())
npm, dec = loop.neural_points, loop.sdf_mlp
for f in range(1, 5):
    gt, scan, source = loop.preprocess(f)
    last = loop.poses[-1]
    guess = ((gt @ torch.linalg.inv(trajectory_pose(f - 1))).cuda() @ last) if len(loop.poses) < 2 else last @ torch.linalg.inv(loop.poses[-2]) @ last
    T = guess.clone().contiguous()
    print(f"--- frame {f}: source {source.shape[0]} guess err {float((guess[:3,3].cpu()-gt[:3,3]).norm()):.3f}")
    for it in range(8):
        o = npm.query_sdf(source, dec, need_grad=True, transform=T, want_xyz=True)
        gn = o["grad"].norm(dim=-1)
        res, sums = loop.tracker._gn(o["xyz"], o, None, None, loop.cfg.reg_min_grad_norm, loop.cfg.reg_max_grad_norm,
                                     loop.cfg.reg_GM_dist_m, loop.cfg.reg_GM_grad, loop.cfg.reg_lm_lambda, T)
        r = res.cpu().numpy()
        print(f"  it{it}: valid {int(r[16])} res_cm {r[17]:.2f} |dt| {float((r[3]**2+r[7]**2+r[11]**2)**0.5):.4f} "
              f"gnorm med {float(gn.median()):.3f} nn>=6 {float((o['nn_count']>=6).float().mean()):.2f} "
              f"sdf med {float(o['sdf'].abs().median()):.3f} err {float((T[:3,3].cpu()-gt[:3,3]).norm()):.3f}")
    info = loop.step(f)
    print("   step:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in info.items()}, loop.mapper.last_losses.tolist())
