#!/usr/bin/env python
"""Run the engine over a KITTI odometry sequence directory (calib.txt, image_0/, image_1/) exactly as the reference's
run_kitti_stereo tool does (src/visual_odometry.cpp: half-resolution processing, config default.yaml), write
keyframes.txt / landmarks.pcd, and — when the ground-truth poses file is given — report ATE / RPE.

    python scripts/run_kitti.py /data/kitti/sequences/05 --poses /data/kitti/poses/05.txt --out /tmp/out05
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sequence")
    ap.add_argument("--poses", default=None, help="KITTI ground truth poses/XX.txt")
    ap.add_argument("--out", default=None, help="directory for keyframes.txt / landmarks.pcd")
    ap.add_argument("--max-frames", type=int, default=0)
    ap.add_argument("--num-features", type=int, default=150)
    ap.add_argument("--num-active-keyframes", type=int, default=10)
    ap.add_argument("--no-backend", action="store_true")
    args = ap.parse_args()
    import svslam
    from svslam import kitti
    seq = kitti.KittiSequence(args.sequence, 0, 1, half=True)
    first = seq.frame_by_id(0)
    if first is None:
        raise SystemExit("no images in %s" % args.sequence)
    H, W = first[0].shape
    ctx = svslam.Context(0)
    slam = ctx.slam(1, W, H, seq.K_left, seq.baseline, half=True, num_features=args.num_features,
                    num_active_keyframes=args.num_active_keyframes, backend_on=0 if args.no_backend else 1)
    poses, status = [], []
    t0 = time.perf_counter()
    for i, (l, r) in enumerate(seq):
        if args.max_frames and i >= args.max_frames:
            break
        poses.append(slam.add_frames(l[None], r[None])[0].copy())
        status.append(int(slam.status[0]))
        if status[-1] == 3:
            break
    dt = time.perf_counter() - t0
    out = {"frames": len(poses), "seconds": dt, "frames_per_sec": len(poses) / dt, "lost": status[-1] == 3,
           "keyframes": int(len(slam.keyframes(0)[0])), "landmarks": int(len(slam.landmarks(0)[0]))}
    if args.out:
        kitti.save_slam_output(slam, 0, args.out, args.sequence, 0)
    if args.poses:
        gt = kitti.read_poses(args.poses)[:len(poses)]
        c, Twc = kitti.pose7_to_Twc(np.array(poses))
        out["ate_rmse_m"] = kitti.ate_rmse(c, gt[:, :, 3])
        te, re = kitti.rpe(Twc, gt)
        out["rpe_trans_m"], out["rpe_rot_rad"] = te, re
    slam.close(); ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
