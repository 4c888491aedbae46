#!/usr/bin/env python3
"""Recipe for oracle/_ref: the UNMODIFIED reference sources this hot path needs, copied from the read-only
reference checkout so that `bench.py --impl reference` (and the per-frame baseline leg) can run the reference's own
`NeuralPoints` / `Decoder` / `Tracker` / `Mapper` on the GPU box, where /root/reference does not exist.

    python oracle/make_ref.py [--src /root/reference]

oracle/_ref is git-ignored (no reference source enters the history) but travels with `gpurun` snapshots, exactly like
a compiled C reference would.  `__graft_entry__.build()` runs this whenever the reference checkout is present.
Nothing under pin_slam_b200/ imports from here: it is test / benchmark infrastructure, like the rest of oracle/.
"""
import argparse
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "model/__init__.py", "model/neural_points.py", "model/decoder.py",
    "utils/__init__.py", "utils/config.py", "utils/tools.py", "utils/tracker.py", "utils/mapper.py",
    "utils/loss.py", "utils/data_sampler.py",
    "config/lidar_slam/run_kitti.yaml", "config/rgbd_slam/run_replica.yaml", "LICENSE",
]


def make(src="/root/reference"):
    if not os.path.isdir(src):
        return False
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if (not os.path.exists(d)) or os.path.getmtime(s) > os.path.getmtime(d) or os.path.getsize(s) != os.path.getsize(d):
            shutil.copyfile(s, d)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("Verbatim copies of PRBonn/PIN_SLAM files (see LICENSE) made by oracle/make_ref.py from %s;\n"
                "benchmark baseline only, never imported by pin_slam_b200/.\n" % src)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    print("oracle/_ref ready" if make(a.src) else f"{a.src} not found: oracle/_ref not (re)built")
