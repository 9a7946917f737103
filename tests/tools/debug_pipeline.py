"""Per-frame diff trace: GPU pipeline vs CPU oracle pipeline (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
import numpy as np, cv2
import svslam
from svslam import synth
from oracle import pipeline as op, cv_stages

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 36
backend = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nak = int(sys.argv[3]) if len(sys.argv) > 3 else 10
g = cv_stages.calibrate_granule(cv2)
cor = synth.Corridor("kitti05", seed=0, n_frames=80)
L, R, T = cor.sequence(nf)
ctx = svslam.Context(0)
slam = ctx.slam(1, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, backend_on=backend, num_active_keyframes=nak, oracle_simd_granule=g)
o = op.Pipeline(cor.K_half(), cor.baseline, op.Cfg(backend_on=backend, num_active_keyframes=nak, granule=g), stages="oracle")
for i in range(nf):
    p = slam.add_frames(L[i:i+1], R[i:i+1])[0].copy()
    w = o.add_frame(L[i], R[i])
    xy, ids, _ = slam.features(0)
    wxy, wids = o.current_features()
    same_n = len(xy) == len(wxy)
    dxy = np.abs(xy - wxy).max() if same_n and len(xy) else -1
    nbad = int((np.abs(xy - wxy).max(1) > 1e-6).sum()) if same_n and len(xy) else -1
    print("f%02d st %d/%d kf %d/%d inl %3d/%3d nfeat %3d/%3d ids_eq %s dxy %.2e (#%d) dpose %.2e gt_err %.3f/%.3f" % (
        i, slam.status[0], o.status, slam.is_kf[0], o.is_kf, slam.inliers[0], o.tracking_inliers, len(xy), len(wxy),
        same_n and np.array_equal(ids, wids), dxy, nbad, np.abs(p - w).max(), np.linalg.norm(p[4:]-T[i][4:]), np.linalg.norm(w[4:]-T[i][4:])))
    lid, lxyz, _ = slam.landmarks(0)
    wl = np.array([o.lms[k].pos for k in sorted(o.lms)]).reshape(-1, 3)
    if len(lid) == len(wl) and len(wl):
        d = np.linalg.norm(lxyz - wl, axis=1) / np.linalg.norm(wl, axis=1)
        print("      landmarks %d max rel %.2e median %.2e" % (len(wl), d.max(), np.median(d)))
    else:
        print("      landmarks count %d vs %d" % (len(lid), len(wl)))
