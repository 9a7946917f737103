"""The oracle PIPELINE (oracle/pipeline.py: the Python mirror of Frontend / Backend / Map that the GPU pipeline is checked
against) run once on the oracle's own stage restatements and once on cv2's real routines with the reference's arguments
(`stages="cv2"`, the configuration bench.py times as the CPU baseline): the two must take the same discrete decisions
frame by frame.  Teacher-forced like the GPU lock-step test (LK differs from cv2 in rare points, test_oracle_fuzz.py).
CPU only."""
import cv2
import numpy as np

from oracle import pipeline as op
from svslam import synth


def _state(p):
    f = p.cur
    xy = np.array([[q.x, q.y] for q in f.fl], np.float32).reshape(-1, 2)
    rxy = np.array([[q.x, q.y] if q is not None else [0, 0] for q in f.fr], np.float32).reshape(-1, 2)
    rvalid = np.array([q is not None for q in f.fr], bool)
    return f.pose.copy(), xy, rxy, rvalid, {k: v.pose.copy() for k, v in p.kfs.items()}, {k: v.pos.copy() for k, v in p.lms.items()}


def test_oracle_pipeline_equals_cv2_pipeline(granule):
    cor = synth.Corridor("kitti05", seed=0, n_frames=80)
    n = 26
    L, R, T = cor.sequence(n)
    a = op.Pipeline(cor.K_half(), cor.baseline, op.Cfg(granule=granule), stages="oracle")
    b = op.Pipeline(cor.K_half(), cor.baseline, op.Cfg(granule=granule), stages="cv2", cv2=cv2)
    nk = 0
    worst_xy = worst_pose = 0.0
    for i in range(n):
        pa, pb = a.add_frame(L[i], R[i]), b.add_frame(L[i], R[i])
        assert a.status == b.status and a.is_kf == b.is_kf, i
        assert i == 0 or a.tracking_inliers == b.tracking_inliers, i
        (xa, ia), (xb, ib) = a.current_features(), b.current_features()
        assert len(xa) == len(xb) and np.array_equal(ia, ib), i
        assert sorted(a.lms) == sorted(b.lms) and sorted(a.kfs) == sorted(b.kfs) and sorted(a.active_kfs) == sorted(b.active_kfs), i
        worst_xy = max(worst_xy, float(np.abs(xa - xb).max()))
        worst_pose = max(worst_pose, float(np.abs(pa - pb).max()))
        d = np.abs(xa - xb).max(1)
        assert np.quantile(d, 0.97) < 1e-3 and d.max() < 0.5, (i, float(d.max()))      # rare f32-lane vs exact-sum LK differences
        assert np.abs(pa - pb).max() < 1e-4 * max(1.0, np.abs(pa[4:]).max()), i
        nk += int(a.is_kf)
        b.force_state(*_state(a))
    assert nk >= 3 and a.status == 1
    print("oracle-stage vs cv2-stage pipeline over %d frames: worst keypoint diff %.2e px, worst pose diff %.2e" % (n, worst_xy, worst_pose))
