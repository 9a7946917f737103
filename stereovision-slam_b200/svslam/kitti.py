"""KITTI odometry sequence reader and trajectory evaluation — the data formats either side of the hot path
(SURVEY.md §8f ranks 1-2).

  KittiSequence   <- slam::Dataset (reference src/dataset.cpp:24-173): calib.txt, image_<cam>/%06d.png pairs.  The PNG
                     decode stays on the host (cv2.imread, out of scope of the path); the half-resolution resize the
                     reference applies in Dataset::NextFrame is done by the engine (svs_frameset_push / k_half_nearest).
  read_poses / ate_rmse / rpe   the evaluator the reference does not have: KITTI poses/XX.txt ground truth (T_w_cam0,
                     3x4 row-major per line) against keyframes.txt / per-frame poses (T_cw).
"""
import ctypes as C
import os

import numpy as np

from . import load_library


def read_calib(calib_path, half=True):
    """-> K [4,4] (fx fy cx cy per camera), t [4,3], baseline [4]   (svs_kitti_read_calib)."""
    lib = load_library()
    K = np.zeros((4, 4)); t = np.zeros((4, 3)); b = np.zeros(4)
    rc = lib.svs_kitti_read_calib(str(calib_path).encode(), int(bool(half)), K.ctypes.data_as(C.c_void_p),
                                  t.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise IOError("cannot read KITTI calibration %s (svs_status %d)" % (calib_path, rc))
    return K, t, b


class KittiSequence:
    """One sequence directory: calib.txt, image_0/ image_1/ (gray) or image_2/ image_3/ (colour)."""

    def __init__(self, path, left_cam=0, right_cam=1, half=True):
        self.path, self.left_cam, self.right_cam, self.half = str(path), left_cam, right_cam, half
        self.K, self.t, self.baselines = read_calib(os.path.join(self.path, "calib.txt"), half)
        self.index = 0

    @property
    def K_left(self):
        return self.K[self.left_cam]

    @property
    def baseline(self):
        """Distance between the two cameras of the pair (Camera::baseline_ of the right camera for the pair (0, 1))."""
        return float(np.linalg.norm(self.t[self.right_cam] - self.t[self.left_cam]))

    def frame_by_id(self, i, gray=True):
        import cv2
        flag = cv2.IMREAD_GRAYSCALE if gray else cv2.IMREAD_COLOR
        l = cv2.imread(os.path.join(self.path, "image_%d" % self.left_cam, "%06d.png" % i), flag)
        r = cv2.imread(os.path.join(self.path, "image_%d" % self.right_cam, "%06d.png" % i), flag)
        if l is None or r is None:
            return None
        return l, r

    def next_frame(self, gray=True):
        """Dataset::NextFrame without the resize: full-resolution pair or None at the end of the sequence."""
        f = self.frame_by_id(self.index, gray)
        if f is not None:
            self.index += 1
        return f

    def __iter__(self):
        self.index = 0
        while True:
            f = self.next_frame()
            if f is None:
                return
            yield f


def read_poses(path):
    """KITTI poses/XX.txt -> [N,3,4] (T_w_cam0 per frame)."""
    a = np.loadtxt(path).reshape(-1, 3, 4)
    return a


def pose7_to_Twc(poses7):
    """Engine poses (T_cw as qx qy qz qw tx ty tz) -> camera centres in the world frame [N,3] and T_wc [N,3,4]."""
    p = np.asarray(poses7, float).reshape(-1, 7)
    x, y, z, w = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    Rt = np.transpose(R, (0, 2, 1))
    c = -np.einsum("nij,nj->ni", Rt, p[:, 4:7])
    return c, np.concatenate([Rt, c[:, :, None]], 2)


def align_rigid(src, dst):
    """Least-squares rotation + translation (Horn / Umeyama without scale) mapping src [N,3] onto dst [N,3]."""
    ms, md = src.mean(0), dst.mean(0)
    U, _, Vt = np.linalg.svd((dst - md).T @ (src - ms))
    S = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    R = U @ S @ Vt
    return R, md - R @ ms


def ate_rmse(est_xyz, gt_xyz, align=True):
    """Absolute trajectory error: RMSE of the camera-centre distances, after a rigid alignment when align=True."""
    est, gt = np.asarray(est_xyz, float), np.asarray(gt_xyz, float)
    if align and len(est) >= 3:
        R, t = align_rigid(est, gt)
        est = est @ R.T + t
    return float(np.sqrt(((est - gt) ** 2).sum(1).mean()))


def rpe(est_Twc, gt_Twc, delta=1):
    """Relative pose error over `delta` frames: (translation RMSE [m], rotation RMSE [rad])."""
    def to4(T):
        T = np.asarray(T, float)
        out = np.tile(np.eye(4), (len(T), 1, 1))
        out[:, :3, :] = T
        return out
    E, G = to4(est_Twc), to4(gt_Twc)
    te, re = [], []
    for i in range(len(E) - delta):
        dE = np.linalg.inv(E[i]) @ E[i + delta]
        dG = np.linalg.inv(G[i]) @ G[i + delta]
        D = np.linalg.inv(dG) @ dE
        te.append(np.linalg.norm(D[:3, 3]))
        re.append(np.arccos(np.clip((np.trace(D[:3, :3]) - 1) / 2, -1, 1)))
    return float(np.sqrt(np.mean(np.square(te)))), float(np.sqrt(np.mean(np.square(re))))


def write_keyframes_txt(path, dataset_dir, left_cam_index, frame_ids, poses7):
    lib = load_library()
    ids = np.ascontiguousarray(frame_ids, np.int64)
    p = np.ascontiguousarray(poses7, np.float64).reshape(-1, 7)
    rc = lib.svs_write_keyframes_txt(str(path).encode(), str(dataset_dir).encode(), int(left_cam_index), len(ids),
                                     ids.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise IOError("cannot write %s" % path)


def write_landmarks_pcd(path, xyz):
    lib = load_library()
    p = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
    rc = lib.svs_write_landmarks_pcd(str(path).encode(), len(p), p.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise IOError("cannot write %s" % path)


def save_slam_output(slam, stream, out_dir, dataset_dir="", left_cam_index=0):
    """VisualOdometry::saveSLAMOutputInFile (src/visual_odometry.cpp:198-310) for one stream of a Slam pipeline."""
    kf_ids, frame_ids, poses = slam.keyframes(stream)
    _, xyz, _ = slam.landmarks(stream)
    os.makedirs(out_dir, exist_ok=True)
    write_landmarks_pcd(os.path.join(out_dir, "landmarks.pcd"), xyz)
    write_keyframes_txt(os.path.join(out_dir, "keyframes.txt"), dataset_dir, left_cam_index, frame_ids, poses)
    return len(kf_ids), len(xyz)
