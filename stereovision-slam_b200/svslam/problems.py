"""Synthetic bundle-adjustment problems at BASELINE config-4 scale and the small geometry helpers they
need.  Shared by bench.py and the tests (the tests import them from here through tests/util.py)."""
import numpy as np

# KITTI seq-05 half-resolution calibration
K05 = np.array([707.0912 * 0.5, 707.0912 * 0.5, 601.8873 * 0.5, 183.1104 * 0.5])
BASELINE = 0.5371657
EXT_L = np.array([0, 0, 0, 1, 0, 0, 0.0])
EXT_R = np.array([0, 0, 0, 1, -BASELINE, 0, 0.0])
W05, H05 = 613, 185


def quat_from_rotvec(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.array([0.5 * w[0], 0.5 * w[1], 0.5 * w[2], 1.0])
    return np.concatenate([np.sin(th / 2) * w / th, [np.cos(th / 2)]])


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def pose_Tcw(center, rotvec):
    """T_cw from a camera centre and a world->camera rotation vector."""
    q = quat_from_rotvec(np.asarray(rotvec, float))
    R = quat_to_R(q)
    return np.concatenate([q, -R @ np.asarray(center, float)])


def project(T, p, K=K05, ext=EXT_L):
    pc = quat_to_R(T[:4]) @ p + T[4:]
    pc = quat_to_R(ext[:4]) @ pc + ext[4:]
    return np.array([K[0] * pc[0] / pc[2] + K[2], K[1] * pc[1] / pc[2] + K[3]]), pc[2]


def R_to_quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w)
    y = (R[0, 2] - R[2, 0]) / (4 * w)
    z = (R[1, 0] - R[0, 1]) / (4 * w)
    q = np.array([x, y, z, w])
    return q / np.linalg.norm(q)


def ba_problem_big(seed, n_kf=50, n_lm=100000, noise=0.5, outlier_frac=0.02, max_follow=6):
    """Vectorised generator of the BASELINE config-4 BA problem: N poses on an arc, L landmarks in the
    frusta at 5-80 m, observed by the birth keyframe (left+right) and the next k in U{1..6} keyframes (left)."""
    rng = np.random.RandomState(seed)
    poses_true = np.array([pose_Tcw([4.0 * np.sin(0.06 * k), 0.0, 1.0 * k], [0.0, -0.02 * k, 0.0]) for k in range(n_kf)])
    Rs = np.array([quat_to_R(p[:4]) for p in poses_true]); ts = poses_true[:, 4:]
    birth = rng.randint(0, n_kf, n_lm)
    z = rng.uniform(5, 80, n_lm); u = rng.uniform(20, W05 - 20, n_lm); v = rng.uniform(10, H05 - 10, n_lm)
    pc = np.stack([(u - K05[2]) * z / K05[0], (v - K05[3]) * z / K05[1], z], 1)
    pw = np.einsum("nji,nj->ni", Rs[birth], pc - ts[birth])
    follow = rng.randint(1, max_follow + 1, n_lm)
    ekf, elm, ecam, euv = [], [], [], []
    for d in range(0, max_follow + 1):
        for cam in ((0, 1) if d == 0 else (0,)):
            k = birth + d
            m = (k < n_kf) & (d <= follow)
            kk = k[m]
            p = np.einsum("nij,nj->ni", Rs[kk], pw[m]) + ts[kk]
            if cam:
                p = p + EXT_R[4:]
            px = np.stack([K05[0] * p[:, 0] / p[:, 2] + K05[2], K05[1] * p[:, 1] / p[:, 2] + K05[3]], 1)
            good = (p[:, 2] > 1) & (px[:, 0] >= 0) & (px[:, 0] < W05) & (px[:, 1] >= 0) & (px[:, 1] < H05)
            idx = np.flatnonzero(m)[good]
            n2 = rng.randn(len(idx), 2) * noise
            out = rng.rand(len(idx)) < outlier_frac
            n2[out] += rng.uniform(-20, 20, (int(out.sum()), 2))
            ekf.append(kk[good]); elm.append(idx); ecam.append(np.full(len(idx), cam)); euv.append(px[good] + n2)
    ekf, elm, ecam, euv = np.concatenate(ekf), np.concatenate(elm), np.concatenate(ecam), np.concatenate(euv)
    order = np.lexsort((ecam, ekf, elm))          # landmark-major, as Backend::Optimize walks the graph
    poses0 = poses_true.copy()
    for k in range(n_kf):
        q0 = quat_from_rotvec(rng.randn(3) * 0.0035)
        poses0[k, :4] = R_to_quat(quat_to_R(q0) @ Rs[k])
        poses0[k, 4:] += rng.randn(3) * 0.02
    return dict(poses=poses0, lms=pw + rng.randn(n_lm, 3) * 0.1, edge_kf=ekf[order].astype(np.int32),
                edge_lm=elm[order].astype(np.int32), edge_cam=ecam[order].astype(np.uint8), edge_uv=euv[order])
