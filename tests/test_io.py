"""Host-side data formats either side of the hot path (no GPU): KITTI calib reader (reference src/dataset.cpp:24-80),
keyframes.txt / landmarks.pcd writers (src/visual_odometry.cpp:198-310), sequence reader and the ATE / RPE evaluator."""
import os

import numpy as np
import pytest

from svslam import kitti

# KITTI odometry sequence 00 calibration (P0..P3), as distributed in calib.txt
CALIB00 = """P0: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 0.000000000000e+00 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 0.000000000000e+00
P1: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 -3.861448000000e+02 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 0.000000000000e+00
P2: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 4.538225000000e+01 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 -1.130887000000e-01 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 3.779761000000e-03
P3: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 -3.372877000000e+02 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 2.369057000000e+00 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 4.915215000000e-03
"""


def _write_calib(d):
    p = os.path.join(d, "calib.txt")
    with open(p, "w") as f:
        f.write(CALIB00)
    return p


@pytest.mark.parametrize("half", [True, False])
def test_calib_matches_reference_formula(tmp_path, half):
    K, t, b = kitti.read_calib(_write_calib(str(tmp_path)), half)
    P = np.array([[float(x) for x in line.split()[1:]] for line in CALIB00.strip().splitlines()]).reshape(4, 3, 4)
    for i in range(4):
        Kf = P[i][:, :3]
        ti = np.linalg.inv(Kf) @ P[i][:, 3]            # src/dataset.cpp:63-66
        assert np.allclose(t[i], ti, rtol=0, atol=1e-12)
        assert abs(b[i] - np.linalg.norm(ti)) < 1e-12
        s = 0.5 if half else 1.0                        # :73
        assert np.allclose(K[i], [Kf[0, 0] * s, Kf[1, 1] * s, Kf[0, 2] * s, Kf[1, 2] * s], rtol=0, atol=1e-12)
    assert abs(b[1] - 0.5371657188644179) < 1e-9       # the seq-00 stereo baseline
    with pytest.raises(IOError):
        kitti.read_calib(os.path.join(str(tmp_path), "missing.txt"))


def test_sequence_reader_and_end(tmp_path):
    cv2 = pytest.importorskip("cv2")
    d = str(tmp_path)
    _write_calib(d)
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 255, (2, 3, 37, 122), dtype=np.uint8)
    for cam in (0, 1):
        os.makedirs(os.path.join(d, "image_%d" % cam))
        for i in range(3):
            cv2.imwrite(os.path.join(d, "image_%d" % cam, "%06d.png" % i), imgs[cam, i])
    seq = kitti.KittiSequence(d)
    got = list(seq)
    assert len(got) == 3 and seq.next_frame() is None
    for i, (l, r) in enumerate(got):
        assert np.array_equal(l, imgs[0, i]) and np.array_equal(r, imgs[1, i])
    assert abs(seq.baseline - 0.5371657188644179) < 1e-9 and abs(seq.K_left[0] - 359.428) < 1e-9
    assert np.array_equal(seq.frame_by_id(1)[1], imgs[1, 1])


def _rand_poses(n, seed):
    rng = np.random.RandomState(seed)
    q = rng.randn(n, 4); q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([q, rng.randn(n, 3) * 10], 1)


def test_keyframes_txt_format(tmp_path):
    p7 = _rand_poses(5, 1)
    ids = np.array([0, 7, 13, 22, 40], np.int64)
    path = os.path.join(str(tmp_path), "keyframes.txt")
    kitti.write_keyframes_txt(path, "/data/sequences/05", 0, ids, p7)
    lines = open(path).read().splitlines()
    assert lines[0] == "/data/sequences/05" and lines[1] == "0" and len(lines) == 2 + 5
    for k, ln in enumerate(lines[2:]):
        v = ln.split(" ")
        assert len(v) == 13 and int(v[0]) == ids[k]
        T = np.array([float(x) for x in v[1:]]).reshape(3, 4)
        x, y, z, w = p7[k, :4]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(T[:, :3], R, atol=1e-5) and np.allclose(T[:, 3], p7[k, 4:], rtol=1e-5)   # %g: 6 significant digits
    kitti.write_keyframes_txt(path, "x", 2, np.zeros(0, np.int64), np.zeros((0, 7)))
    assert open(path).read() == "x\n2\n"


def test_landmarks_pcd_format(tmp_path):
    xyz = np.random.RandomState(2).randn(11, 3) * 30
    path = os.path.join(str(tmp_path), "landmarks.pcd")
    kitti.write_landmarks_pcd(path, xyz)
    lines = open(path).read().splitlines()
    assert lines[:11] == ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS x y z", "SIZE 4 4 4", "TYPE F F F",
                          "COUNT 1 1 1", "WIDTH 11", "HEIGHT 1", "VIEWPOINT 0 0 0 1 0 0 0", "POINTS 11", "DATA ascii"]
    got = np.array([[float(v) for v in ln.split()] for ln in lines[11:]])
    assert got.shape == (11, 3) and np.array_equal(got.astype(np.float32), xyz.astype(np.float32))


def test_ate_and_rpe():
    rng = np.random.RandomState(3)
    t = np.linspace(0, 6, 200)
    gt = np.stack([10 * np.cos(t), 0.1 * t, 10 * np.sin(t)], 1)
    a = 0.7
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    est = gt @ R.T + np.array([3.0, -2.0, 5.0])
    assert kitti.ate_rmse(est, gt) < 1e-9                          # a rigid transform is removed by the alignment
    assert kitti.ate_rmse(est, gt, align=False) > 1.0
    noisy = est + rng.randn(*est.shape) * 0.05
    assert 0.06 < kitti.ate_rmse(noisy, gt) < 0.11                 # sigma * sqrt(3) ~ 0.087
    # poses: identity rotation, centres on the curve -> T_cw = [I | -c]
    p7 = np.concatenate([np.tile([0, 0, 0, 1.0], (200, 1)), -gt], 1)
    c, Twc = kitti.pose7_to_Twc(p7)
    assert np.allclose(c, gt)
    te, re = kitti.rpe(Twc, Twc)
    assert te < 1e-12 and re < 1e-7
    Twc2 = Twc.copy(); Twc2[:, :, 3] *= 1.01                       # 1 % scale drift -> translation RPE ~ 1 % of the step
    te, _ = kitti.rpe(Twc2, Twc)
    step = np.linalg.norm(np.diff(gt, axis=0), axis=1).mean()
    assert 0.005 * step < te < 0.02 * step
