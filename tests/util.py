"""Shared synthetic inputs for the tests (SURVEY.md §8d value distributions)."""
import cv2
import numpy as np


def texture(h, w, seed, rects=None):
    """Gaussian-blurred (sigma 2) uniform noise normalised to 0..255 plus random filled rectangles."""
    rng = np.random.RandomState(seed)
    a = rng.rand(h, w).astype(np.float32)
    a = cv2.GaussianBlur(a, (0, 0), 2.0)
    a = (a - a.min()) / (a.max() - a.min()) * 255
    img = a.astype(np.uint8)
    if rects is None:
        rects = max(4, int(40 * h * w / (620 * 188)))
    for _ in range(rects):
        x, y = rng.randint(0, w), rng.randint(0, h)
        ww, hh = rng.randint(5, 60), rng.randint(5, 40)
        cv2.rectangle(img, (x, y), (x + ww, y + hh), int(rng.randint(0, 256)), -1)
    return img


def moved_pair(h, w, seed):
    """An image and a slightly affinely-warped copy of it (for LK)."""
    big = texture(h + 40, w + 40, seed)
    a = big[20:20 + h, 20:20 + w].copy()
    m = np.array([[1.01, 0.005, 3.3], [-0.004, 0.995, -2.1]], np.float32)
    b = cv2.warpAffine(big, m, (w + 40, h + 40))[20:20 + h, 20:20 + w].copy()
    return a, b


def lk_points(a, seed, n_corner=300, n_rand=60):
    h, w = a.shape
    kps = cv2.GFTTDetector_create(n_corner, 0.01, 10).detect(a, None)
    p0 = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    rng = np.random.RandomState(seed)
    extra = np.stack([rng.rand(n_rand) * w, rng.rand(n_rand) * h], 1).astype(np.float32)
    edge = np.array([[0.5, 0.5], [w - 1.2, h - 1.1], [2, h - 2], [w - 3, 3], [w - 0.5, 10], [5, h - 0.6]], np.float32)
    p0 = np.concatenate([p0, extra, edge])
    init = p0 + rng.randn(*p0.shape).astype(np.float32) * 2.0 + np.array([3, -2], np.float32)
    init[-3:] += np.array([30, 30], np.float32)
    return p0, init.astype(np.float32)


def stereo_pair(h, w, seed, dmin=2.0, dmax=120.0):
    """Rectified pair with a horizontal disparity ramp dmin..dmax: R(x) = L(x + d(x))."""
    big = texture(h, w + 130, seed)
    disp = (np.linspace(dmin, dmax, w)[None, :] * np.ones((h, 1))).astype(np.float32)
    xs = np.arange(w)[None, :] + disp
    ys = (np.arange(h)[:, None] * np.ones((1, w))).astype(np.float32)
    r = cv2.remap(big, xs.astype(np.float32), ys, cv2.INTER_LINEAR)
    return big[:, :w].copy(), r


# ---------------------------------------------------------------------------------------------
# Geometry problems (KITTI seq-05 half-resolution calibration); the generators bench.py also uses
# live in the package
from svslam.problems import (K05, BASELINE, EXT_L, EXT_R, W05, H05, quat_from_rotvec, quat_to_R, pose_Tcw, project,  # noqa: E402,F401
                             R_to_quat, ba_problem_big)


def pose_problem(seed, m=150, noise=0.5, outlier_frac=0.1, perturb=(0.05, 0.01)):
    """A pose-only problem like Frontend::EstimateCurrentPose sees: m landmarks, noisy pixels, gross outliers."""
    rng = np.random.RandomState(seed)
    T_true = pose_Tcw(rng.randn(3) * 0.3, rng.randn(3) * 0.05)
    pts, uv = [], []
    while len(pts) < m:
        z = rng.uniform(4, 60)
        u, v = rng.uniform(5, W05 - 5), rng.uniform(5, H05 - 5)
        pc = np.array([(u - K05[2]) * z / K05[0], (v - K05[3]) * z / K05[1], z])
        R = quat_to_R(T_true[:4])
        pw = R.T @ (pc - T_true[4:])
        pts.append(pw)
        uv.append([u, v])
    pts, uv = np.array(pts).reshape(-1, 3), np.array(uv).reshape(-1, 2) + rng.randn(m, 2) * noise
    nout = int(m * outlier_frac)
    uv[:nout] += rng.uniform(-30, 30, (nout, 2))
    T0 = pose_Tcw(-quat_to_R(T_true[:4]).T @ T_true[4:] + rng.randn(3) * perturb[0], rng.randn(3) * perturb[1] * 0 )
    # perturbed initial pose: same construction with a noisy rotation
    q0 = quat_from_rotvec(rng.randn(3) * perturb[1])
    T0 = np.concatenate([q0, T_true[4:] + rng.randn(3) * perturb[0]])
    # compose the small rotation on the left of T_true's rotation
    Rn = quat_to_R(q0) @ quat_to_R(T_true[:4])
    T0[:4] = R_to_quat(Rn)
    return pts, uv, K05.copy(), T0, T_true


def ba_problem(seed, n_kf=10, n_lm=300, noise=0.5, outlier_frac=0.02, pose_sigma=(0.02, 0.0035), lm_sigma=0.1,
               max_follow=6, unused_kf=False, unused_lm=0):
    """A sliding-window BA problem shaped like Backend::Optimize's graph (SURVEY.md §8d config 4 recipe, scaled)."""
    rng = np.random.RandomState(seed)
    poses_true = []
    for k in range(n_kf):
        ang = 0.02 * k
        center = np.array([4.0 * np.sin(ang * 3), 0.02 * rng.randn(), 1.0 * k])
        poses_true.append(pose_Tcw(center, [0.0, -ang, 0.0]))
    poses_true = np.array(poses_true)
    lms_true, ekf, elm, ecam, euv = [], [], [], [], []
    for l in range(n_lm):
        birth = rng.randint(0, n_kf)
        for _ in range(50):
            z = rng.uniform(5, 80)
            u, v = rng.uniform(20, W05 - 20), rng.uniform(10, H05 - 10)
            pc = np.array([(u - K05[2]) * z / K05[0], (v - K05[3]) * z / K05[1], z])
            T = poses_true[birth]
            pw = quat_to_R(T[:4]).T @ (pc - T[4:])
            obs = []
            for (k, cam) in [(birth, 0), (birth, 1)] + [(k, 0) for k in range(birth + 1, min(n_kf, birth + 1 + rng.randint(1, max_follow + 1)))]:
                px, depth = project(poses_true[k], pw, K05, EXT_R if cam else EXT_L)
                if depth > 1 and 0 <= px[0] < W05 and 0 <= px[1] < H05:
                    obs.append((k, cam, px))
            if len(obs) >= 2:
                break
        lms_true.append(pw)
        for (k, cam, px) in obs:
            ekf.append(k); elm.append(l); ecam.append(cam)
            n2 = rng.randn(2) * noise
            if rng.rand() < outlier_frac:
                n2 += rng.uniform(-20, 20, 2)
            euv.append(px + n2)
    lms_true = np.array(lms_true)
    poses0 = poses_true.copy()
    for k in range(n_kf):
        q0 = quat_from_rotvec(rng.randn(3) * pose_sigma[1])
        poses0[k, :4] = R_to_quat(quat_to_R(q0) @ quat_to_R(poses_true[k, :4]))
        poses0[k, 4:] += rng.randn(3) * pose_sigma[0]
    lms0 = lms_true + rng.randn(*lms_true.shape) * lm_sigma
    prob = dict(poses=poses0, lms=lms0, edge_kf=np.array(ekf, np.int32), edge_lm=np.array(elm, np.int32),
                edge_cam=np.array(ecam, np.uint8), edge_uv=np.array(euv))
    if unused_kf:       # a keyframe with no edges (inactive vertex) at index 0
        prob["poses"] = np.concatenate([pose_Tcw([9, 9, 9], [0.1, 0.2, 0.3])[None], prob["poses"]])
        prob["edge_kf"] = prob["edge_kf"] + 1
    if unused_lm:
        prob["lms"] = np.concatenate([prob["lms"], rng.randn(unused_lm, 3) * 10])
    return prob, poses_true, lms_true


def rel_to_norm(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-12)




def pose_graph_problem(seed, n=40, drift=(0.02, 0.004), loops=((-1, 2),), step=0.15):
    """A keyframe chain with odometry drift and loop edges, like LoopClosure::PoseGraphOptimization sees
    (src/loopclosure.cpp:641-746): returns dict(poses (drifted initial estimate), fixed, edge_a, edge_b, meas, gt)."""
    from oracle import geom
    rng = np.random.RandomState(seed)

    def rnd(st, sr):
        return geom.se3_exp(np.concatenate([rng.randn(3) * st, rng.randn(3) * sr]))
    gt = [geom.se3_inv(geom.se3_exp(np.array([3 * np.sin(step * i), 0.1 * np.sin(2 * step * i), 3 * (1 - np.cos(step * i)), 0, step * i, 0])))
          for i in range(n)]
    ea, eb, meas = [], [], []
    for i in range(1, n):
        ea.append(i); eb.append(i - 1)
        meas.append(geom.se3_mul(geom.se3_mul(gt[i], geom.se3_inv(gt[i - 1])), rnd(*drift)))
    for (a, b) in loops:
        a, b = a % n, b % n
        ea.append(a); eb.append(b)
        meas.append(geom.se3_mul(geom.se3_mul(gt[a], geom.se3_inv(gt[b])), rnd(drift[0] * 0.2, drift[1] * 0.2)))
    init = [gt[0]]
    for i in range(1, n):
        init.append(geom.se3_mul(meas[i - 1], init[-1]))
    fixed = np.zeros(n, np.uint8)
    fixed[0] = 1
    return dict(poses=np.array(init), fixed=fixed, edge_a=np.array(ea, np.int32), edge_b=np.array(eb, np.int32), meas=np.array(meas), gt=np.array(gt))
