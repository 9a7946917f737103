"""ctypes loader for oracle/geom.c + NumPy triangulation restatement.

TEST INFRASTRUCTURE ONLY (see oracle/geom.c header).  `build()` compiles the C
restatement with gcc into oracle/_build/liboracle.so (git-ignored).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "geom.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99",
                           "-o", _SO, src, "-lm"])
    return _SO


class OrcImg(C.Structure):
    _fields_ = [("p", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("stride", C.c_int)]


class LmStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("trials", C.c_int), ("linearizations", C.c_int),
                ("solves", C.c_int), ("lambda_", C.c_double), ("chi2", C.c_double)]


class BaStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("trials", C.c_int), ("linearizations", C.c_int),
                ("solves", C.c_int), ("lambda_", C.c_double), ("chi2", C.c_double),
                ("chi2_init", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def lk_track(prev_levels, next_levels, prev_xy, init_xy, win=11, max_iter=30, eps=0.01):
    """cv::calcOpticalFlowPyrLK(..., Size(win,win), len(levels)-1, {COUNT+EPS,max_iter,eps},
    OPTFLOW_USE_INITIAL_FLOW) on prebuilt pyramids.  Returns (next_xy, status, iters)."""
    n = len(prev_xy)
    nl = min(len(prev_levels), len(next_levels))
    keep = [np.ascontiguousarray(a, np.uint8) for a in list(prev_levels[:nl]) + list(next_levels[:nl])]
    P = (OrcImg * nl)(*[OrcImg(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0]) for a in keep[:nl]])
    N = (OrcImg * nl)(*[OrcImg(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0]) for a in keep[nl:]])
    pxy = np.ascontiguousarray(prev_xy, np.float32).reshape(-1, 2)
    nxy = np.array(init_xy, np.float32).reshape(-1, 2).copy()
    st = np.zeros(n, np.uint8)
    it = np.zeros(n, np.int32)
    lib().orc_lk_track(P, N, C.c_int(nl), _p(pxy), _p(nxy), C.c_int(n), C.c_int(win),
                       C.c_int(max_iter), C.c_double(eps), _p(st), _p(it))
    return nxy, st, it


def pose_only_lm(pts_w, uv, K4, T0, chi2_th=5.991, rounds=4, iters=10):
    """Frontend::EstimateCurrentPose's g2o block (src/frontend.cpp:408-527)."""
    pts_w = np.ascontiguousarray(pts_w, np.float64).reshape(-1, 3)
    uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
    m = len(pts_w)
    K4 = np.ascontiguousarray(K4, np.float64)
    T0 = np.ascontiguousarray(T0, np.float64)
    T = np.zeros(7)
    outl = np.zeros(max(m, 1), np.uint8)
    ninl = C.c_int(0)
    st = LmStats()
    lib().orc_pose_only_lm(_p(pts_w), _p(uv), C.c_int(m), _p(K4), _p(T0), C.c_double(chi2_th),
                           C.c_int(rounds), C.c_int(iters), _p(T), _p(outl), C.byref(ninl), C.byref(st))
    return T, outl[:m], ninl.value, st


def ba_optimize(poses, lms, edge_kf, edge_lm, edge_cam, edge_uv, K_left, K_right, ext_left, ext_right,
                huber_delta=5.991, max_iter=10, jac_mode=0):
    """Backend::Optimize's g2o block (src/backend.cpp:22-164). Returns new poses, lms, chi2, stats."""
    poses = np.array(poses, np.float64).reshape(-1, 7).copy()
    lms = np.array(lms, np.float64).reshape(-1, 3).copy()
    ekf = np.ascontiguousarray(edge_kf, np.int32)
    elm = np.ascontiguousarray(edge_lm, np.int32)
    ecam = np.ascontiguousarray(edge_cam, np.uint8)
    euv = np.ascontiguousarray(edge_uv, np.float64).reshape(-1, 2)
    E = len(ekf)
    chi2 = np.zeros(max(E, 1))
    st = BaStats()
    args = [np.ascontiguousarray(a, np.float64) for a in (K_left, K_right, ext_left, ext_right)]
    lib().orc_ba_optimize(C.c_int(len(poses)), _p(poses), C.c_int(len(lms)), _p(lms), C.c_int(E),
                          _p(ekf), _p(elm), _p(ecam), _p(euv), _p(args[0]), _p(args[1]), _p(args[2]),
                          _p(args[3]), C.c_double(huber_delta), C.c_int(max_iter), C.c_int(jac_mode),
                          _p(chi2), C.byref(st))
    return poses, lms, chi2[:E], st


def se3_mul(a, b):
    o = np.zeros(7)
    lib().orc_se3_mul(_p(np.ascontiguousarray(a, np.float64)), _p(np.ascontiguousarray(b, np.float64)), _p(o))
    return o


def se3_inv(a):
    o = np.zeros(7)
    lib().orc_se3_inv(_p(np.ascontiguousarray(a, np.float64)), _p(o))
    return o


def se3_exp(v):
    o = np.zeros(7)
    lib().orc_se3_exp(_p(np.ascontiguousarray(v, np.float64)), _p(o))
    return o


def se3_log(a):
    o = np.zeros(6)
    lib().orc_se3_log(_p(np.ascontiguousarray(a, np.float64)), _p(o))
    return o


def se3_act(T, p):
    o = np.zeros(3)
    lib().orc_se3_act(_p(np.ascontiguousarray(T, np.float64)), _p(np.ascontiguousarray(p, np.float64)), _p(o))
    return o


def world2pixel_batch(T, ext, K4, pts_w):
    """Camera::world2pixel for n points (bit-identical to se3_act(ext, se3_act(T, p)) + projection per point)."""
    pts = np.ascontiguousarray(pts_w, np.float64).reshape(-1, 3)
    out = np.zeros((len(pts), 2))
    if len(pts):
        lib().orc_world2pixel_batch(_p(np.ascontiguousarray(T, np.float64)), _p(np.ascontiguousarray(ext, np.float64)),
                                    _p(np.ascontiguousarray(K4, np.float64)), len(pts), _p(pts), _p(out))
    return out


def triangulate(left_xy, right_xy, K_left, K_right, baseline):
    """slam::triangulation for a rectified pair (include/StereoVisionSLAM/algorithm.h:59-86) with
    the callers' pixel2camera (src/camera.cpp:58-72): left extrinsic identity, right [I | (-b,0,0)].
    NumPy/LAPACK SVD (Eigen's bdcSvd is un-vendored).  Returns (xyz [n,3], ok [n])."""
    l = np.asarray(left_xy, np.float32).reshape(-1, 2).astype(np.float64)
    r = np.asarray(right_xy, np.float32).reshape(-1, 2).astype(np.float64)
    n = len(l)
    xyz = np.zeros((n, 3))
    ok = np.zeros(n, np.uint8)
    if n == 0:
        return xyz, ok
    x1 = (l[:, 0] - K_left[2]) / K_left[0]
    y1 = (l[:, 1] - K_left[3]) / K_left[1]
    x2 = (r[:, 0] - K_right[2]) / K_right[0]
    y2 = (r[:, 1] - K_right[3]) / K_right[1]
    A = np.zeros((n, 4, 4))
    m1 = np.hstack([np.eye(3), np.zeros((3, 1))])
    m2 = np.hstack([np.eye(3), np.array([[-baseline], [0.0], [0.0]])])
    A[:, 0] = x1[:, None] * m1[2] - m1[0]
    A[:, 1] = y1[:, None] * m1[2] - m1[1]
    A[:, 2] = x2[:, None] * m2[2] - m2[0]
    A[:, 3] = y2[:, None] * m2[2] - m2[1]
    _, s, vt = np.linalg.svd(A)
    v = vt[:, 3, :]
    xyz = v[:, :3] / v[:, 3:4]
    ok = (s[:, 3] / s[:, 2] < 1e-2).astype(np.uint8)
    return xyz, ok


def pose_graph_optimize(poses, fixed, edge_a, edge_b, meas, max_iter=22, jac_mode=1):
    """LoopClosure::PoseGraphOptimization's g2o block (src/loopclosure.cpp:641-746): returns (poses, stats)."""
    P = np.array(poses, np.float64).reshape(-1, 7).copy()
    fx = np.ascontiguousarray(fixed, np.uint8)
    ea, eb = np.ascontiguousarray(edge_a, np.int32), np.ascontiguousarray(edge_b, np.int32)
    M = np.ascontiguousarray(meas, np.float64).reshape(-1, 7)
    st = BaStats()
    lib().orc_pose_graph_optimize(len(P), _p(P), _p(fx), len(ea), _p(ea), _p(eb), _p(M), int(max_iter), int(jac_mode), C.byref(st))
    return P, st


def pg_edge_jac(M, A, B, mode):
    Ja, Jb = np.zeros((6, 6)), np.zeros((6, 6))
    lib().orc_pg_edge_jac(_p(np.ascontiguousarray(M, np.float64)), _p(np.ascontiguousarray(A, np.float64)),
                          _p(np.ascontiguousarray(B, np.float64)), int(mode), _p(Ja), _p(Jb))
    return Ja, Jb


def move_landmarks(lms, lm_kf, old_poses, new_poses):
    """src/loopclosure.cpp:749-777: pos_w = new_pose(kf)^-1 * (old_pose(kf) * pos) with kf = keyframe of the landmark's first
    valid observation; lm_kf < 0: untouched."""
    out = np.array(lms, np.float64).reshape(-1, 3).copy()
    for i, k in enumerate(lm_kf):
        if k >= 0:
            out[i] = se3_act(se3_inv(new_poses[k]), se3_act(old_poses[k], out[i]))
    return out
