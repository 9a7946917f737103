"""Python binding (ctypes) of the C ABI in include/svslam.h.

This is plumbing for tests and bench.py: NumPy arrays in, NumPy arrays out, every call goes
straight into libsvslam.so (hand-written sm_100a CUDA).  There is no CPU fallback: if the shared
library is missing or no B200 is visible, construction raises.
"""
import ctypes as C
import os
import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_ROOT, "libsvslam.so")
_lib = None


class SvsError(RuntimeError):
    pass


class LmStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("trials", C.c_int32), ("linearizations", C.c_int32),
                ("solves", C.c_int32), ("lambda_", C.c_double), ("chi2", C.c_double)]


class BaStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("trials", C.c_int32), ("linearizations", C.c_int32),
                ("solves", C.c_int32), ("lambda_", C.c_double), ("chi2", C.c_double),
                ("chi2_init", C.c_double)]


def load_library():
    """dlopen libsvslam.so (no device needed).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvsError("libsvslam.so is not built (run `python stereovision-slam_b200/build.py`); "
                       "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.svs_create.restype = C.c_void_p
    lib.svs_create.argtypes = [C.c_int]
    lib.svs_create_error.restype = C.c_char_p
    lib.svs_last_error.restype = C.c_char_p
    lib.svs_last_error.argtypes = [C.c_void_p]
    lib.svs_destroy.argtypes = [C.c_void_p]
    lib.svs_stream.restype = C.c_void_p
    lib.svs_stream.argtypes = [C.c_void_p]
    lib.svs_launch_count.restype = C.c_longlong
    lib.svs_launch_count.argtypes = [C.c_void_p]
    lib.svs_host_alloc.restype = C.c_void_p
    lib.svs_host_alloc.argtypes = [C.c_size_t]
    lib.svs_host_free.argtypes = [C.c_void_p]
    lib.svs_frameset_create.restype = C.c_void_p
    lib.svs_frameset_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    _lib = lib
    return lib


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


class FrameSet:
    def __init__(self, ctx, n_streams, in_w, in_h, half=True, lk_win=11, lk_max_level=3):
        self.ctx = ctx
        self.h = ctx.lib.svs_frameset_create(ctx.h, n_streams, in_w, in_h, int(bool(half)), lk_win, lk_max_level)
        if not self.h:
            raise SvsError(ctx.last_error())
        w, hh, nl = C.c_int(), C.c_int(), C.c_int()
        ctx.lib.svs_frameset_size(C.c_void_p(self.h), C.byref(w), C.byref(hh), C.byref(nl))
        self.n_streams, self.in_w, self.in_h = n_streams, in_w, in_h
        self.w, self.hgt, self.n_levels = w.value, hh.value, nl.value

    def push(self, left, right):
        """left/right: uint8 [n_streams, in_h, in_w] host arrays."""
        left, right = _u8(left), _u8(right)
        assert left.shape == (self.n_streams, self.in_h, self.in_w) == right.shape
        self.ctx._chk(self.ctx.lib.svs_frameset_push(
            C.c_void_p(self.ctx.h), C.c_void_p(self.h), _p(left), _p(right), C.c_size_t(self.in_w),
            C.c_size_t(self.in_w * self.in_h), 0))

    def push_ptr(self, left_ptr, right_ptr, on_device):
        self.ctx._chk(self.ctx.lib.svs_frameset_push(
            C.c_void_p(self.ctx.h), C.c_void_p(self.h), C.c_void_p(left_ptr), C.c_void_p(right_ptr),
            C.c_size_t(self.in_w), C.c_size_t(self.in_w * self.in_h), int(on_device)))

    def push_ptrs(self, left_ptrs, right_ptrs, on_device, row_stride=None):
        """One pointer per stream.  on_device: 0 host (staged DMA), 1 device, 2 pinned host read zero-copy over PCIe."""
        n = self.n_streams
        lp = (C.c_void_p * n)(*[int(x) for x in left_ptrs])
        rp = (C.c_void_p * n)(*[int(x) for x in right_ptrs]) if right_ptrs is not None else None     # None: left eye only
        self.ctx._chk(self.ctx.lib.svs_frameset_push_ptrs(
            C.c_void_p(self.ctx.h), C.c_void_p(self.h), lp, rp, C.c_size_t(row_stride or self.in_w), int(on_device)))

    def prefetch_ptrs(self, left_ptrs, right_ptrs, on_device, row_stride=None):
        """Start the ingest of the NEXT pair on the ingest stream (svs_frameset_prefetch_ptrs)."""
        n = self.n_streams
        lp = (C.c_void_p * n)(*[int(x) for x in left_ptrs])
        rp = (C.c_void_p * n)(*[int(x) for x in right_ptrs]) if right_ptrs is not None else None
        self.ctx._chk(self.ctx.lib.svs_frameset_prefetch_ptrs(
            C.c_void_p(self.ctx.h), C.c_void_p(self.h), lp, rp, C.c_size_t(row_stride or self.in_w), int(on_device)))

    def fetch_right_ptrs(self, stream_ids, right_ptrs, on_device, row_stride=None):
        """Lazy right-eye ingest for the selected streams (svs_frameset_fetch_right_ptrs)."""
        ids = np.ascontiguousarray(stream_ids, np.int32)
        rp = (C.c_void_p * len(ids))(*[int(x) for x in right_ptrs])
        self.ctx._chk(self.ctx.lib.svs_frameset_fetch_right_ptrs(
            C.c_void_p(self.ctx.h), C.c_void_p(self.h), _p(ids), len(ids), rp, C.c_size_t(row_stride or self.in_w), int(on_device)))

    def download(self, stream, which, level):
        w, h = self.w, self.hgt
        for _ in range(level):
            w, h = (w + 1) // 2, (h + 1) // 2
        out = np.zeros((h, w), np.uint8)
        self.ctx._chk(self.ctx.lib.svs_frameset_download(
            C.c_void_p(self.ctx.h), C.c_void_p(self.h), stream, which, level, _p(out), w))
        return out

    def close(self):
        if self.h:
            self.ctx.lib.svs_frameset_destroy(C.c_void_p(self.ctx.h), C.c_void_p(self.h))
            self.h = None


class Context:
    """One svs_ctx (one CUDA stream).  Raises SvsError when no B200 is available."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = self.lib.svs_create(device)
        if not self.h:
            raise SvsError("svs_create failed: " + self.lib.svs_create_error().decode())

    def close(self):
        if self.h:
            self.lib.svs_destroy(C.c_void_p(self.h))
            self.h = None

    def last_error(self):
        return self.lib.svs_last_error(C.c_void_p(self.h)).decode()

    def _chk(self, rc):
        if rc != 0:
            raise SvsError("svs error %d: %s" % (rc, self.last_error()))

    def sync(self):
        self._chk(self.lib.svs_sync(C.c_void_p(self.h)))

    def set_wait_mode(self, mode):
        """0: blocking calls spin on the stream (default); 1: they sleep on a blocking-sync event (svs_set_wait_mode)."""
        self._chk(self.lib.svs_set_wait_mode(C.c_void_p(self.h), int(mode)))

    def reserve_headroom(self, factor=2.0):
        """Grow every variable-size scratch buffer to factor x its largest request so far (svs_reserve_headroom): call after
        warm-up so that steady-state steps never reallocate (a regrowth synchronises the whole device)."""
        self.lib.svs_reserve_headroom.argtypes = [C.c_void_p, C.c_double]
        self._chk(self.lib.svs_reserve_headroom(C.c_void_p(self.h), float(factor)))

    def set_ba_schedule(self, high_priority, threads_per_window=0):
        """Window-solver scheduling (svs_set_ba_schedule): own high-priority stream, forced CTA size (0 = automatic)."""
        self._chk(self.lib.svs_set_ba_schedule(C.c_void_p(self.h), int(high_priority), int(threads_per_window)))

    def stream_ptr(self):
        return self.lib.svs_stream(C.c_void_p(self.h))

    def launch_count(self):
        return int(self.lib.svs_launch_count(C.c_void_p(self.h)))

    def frameset(self, *a, **k):
        return FrameSet(self, *a, **k)

    # ---- a0
    def half_nearest(self, imgs):
        imgs = _u8(imgs)
        single = imgs.ndim == 2
        if single:
            imgs = imgs[None]
        n, h, w = imgs.shape
        dw, dh = int(np.rint(w * 0.5)), int(np.rint(h * 0.5))
        out = np.zeros((n, dh, dw), np.uint8)
        self._chk(self.lib.svs_half_nearest(C.c_void_p(self.h), _p(imgs), w, h, w, n, C.c_size_t(w * h), _p(out)))
        return out[0] if single else out

    # ---- a1
    def corner_min_eig(self, img, granule=32):
        img = _u8(img)
        h, w = img.shape
        out = np.zeros((h, w), np.float32)
        self._chk(self.lib.svs_corner_min_eig(C.c_void_p(self.h), _p(img), w, h, w, granule, _p(out)))
        return out

    def gftt_detect(self, img, mask=None, occupied_xy=None, max_corners=150, quality=0.01, min_distance=20.0,
                    granule=32):
        img = _u8(img)
        h, w = img.shape
        m = _u8(mask) if mask is not None else None
        occ = _f32(occupied_xy).reshape(-1, 2) if occupied_xy is not None else None
        xy = np.zeros((max_corners, 2), np.float32)
        resp = np.zeros(max_corners, np.float32)
        n = C.c_int(0)
        self._chk(self.lib.svs_gftt_detect(
            C.c_void_p(self.h), _p(img), w, h, w, _p(m), w, _p(occ), 0 if occ is None else len(occ), max_corners,
            C.c_double(quality), C.c_double(min_distance), granule, _p(xy), _p(resp), C.byref(n)))
        return xy[:n.value].copy(), resp[:n.value].copy()

    def gftt_detect_batch(self, fs, stream_ids, occupied, max_corners=150, quality=0.01, min_distance=20.0,
                          granule=32):
        """occupied: list (per selected stream) of [k,2] float arrays.  Returns list of (xy, resp)."""
        ids = _i32(stream_ids)
        n_sel = len(ids)
        off = np.zeros(n_sel + 1, np.int32)
        for i, o in enumerate(occupied):
            off[i + 1] = off[i] + len(o)
        occ = _f32(np.concatenate([np.asarray(o, np.float32).reshape(-1, 2) for o in occupied])
                   if off[-1] > 0 else np.zeros((0, 2), np.float32))
        xy = np.zeros((n_sel, max_corners, 2), np.float32)
        resp = np.zeros((n_sel, max_corners), np.float32)
        n = np.zeros(n_sel, np.int32)
        self._chk(self.lib.svs_gftt_detect_batch(
            C.c_void_p(self.h), C.c_void_p(fs.h), _p(ids), n_sel, _p(off), _p(occ), max_corners, C.c_double(quality),
            C.c_double(min_distance), granule, _p(xy), _p(resp), _p(n)))
        return [(xy[i, :n[i]].copy(), resp[i, :n[i]].copy()) for i in range(n_sel)]

    # ---- a2 / a3
    def lk_track(self, prev, nxt, prev_xy, init_xy, win=11, max_level=3, max_iter=30, eps=0.01):
        prev, nxt = _u8(prev), _u8(nxt)
        h, w = prev.shape
        pxy = _f32(prev_xy).reshape(-1, 2)
        nxy = np.array(init_xy, np.float32).reshape(-1, 2).copy()
        st = np.zeros(len(pxy), np.uint8)
        self._chk(self.lib.svs_lk_track(C.c_void_p(self.h), _p(prev), _p(nxt), w, h, w, _p(pxy), _p(nxy), len(pxy),
                                        win, max_level, max_iter, C.c_double(eps), _p(st)))
        return nxy, st

    def lk_track_batch(self, fs, pair, prev_list, init_list, max_iter=30, eps=0.01):
        off = np.zeros(fs.n_streams + 1, np.int32)
        for i, p in enumerate(prev_list):
            off[i + 1] = off[i] + len(p)
        tot = int(off[-1])
        pxy = _f32(np.concatenate([np.asarray(p, np.float32).reshape(-1, 2) for p in prev_list])) if tot else np.zeros((0, 2), np.float32)
        nxy = np.array(np.concatenate([np.asarray(p, np.float32).reshape(-1, 2) for p in init_list]), np.float32) if tot else np.zeros((0, 2), np.float32)
        st = np.zeros(max(tot, 1), np.uint8)
        self._chk(self.lib.svs_lk_track_batch(C.c_void_p(self.h), C.c_void_p(fs.h), pair, _p(off), _p(pxy), _p(nxy),
                                              max_iter, C.c_double(eps), _p(st)))
        return [(nxy[off[i]:off[i + 1]].copy(), st[off[i]:off[i + 1]].copy()) for i in range(fs.n_streams)]


# ---------------------------------------------------------------- geometry / dense (methods added to Context)
def _triangulate(self, left_xy, right_xy, K_left, K_right, baseline):
    l, r = _f32(left_xy).reshape(-1, 2), _f32(right_xy).reshape(-1, 2)
    n = len(l)
    xyz = np.zeros((max(n, 1), 3))
    ok = np.zeros(max(n, 1), np.uint8)
    self._chk(self.lib.svs_triangulate(C.c_void_p(self.h), _p(l), _p(r), n, _p(_f64(K_left)), _p(_f64(K_right)),
                                       C.c_double(baseline), _p(xyz), _p(ok)))
    return xyz[:n], ok[:n]


def _pose_only_lm(self, problems, chi2_th=5.991, rounds=4, iters=10):
    """problems: list of (pts_w [m,3], uv [m,2], K4, T0[7]).  Returns list of (T, outlier, n_inlier, stats)."""
    n = len(problems)
    off = np.zeros(n + 1, np.int32)
    for i, p in enumerate(problems):
        off[i + 1] = off[i] + len(p[0])
    M = int(off[-1])
    pts = _f64(np.concatenate([np.asarray(p[0], np.float64).reshape(-1, 3) for p in problems])) if M else np.zeros((0, 3))
    uv = _f64(np.concatenate([np.asarray(p[1], np.float64).reshape(-1, 2) for p in problems])) if M else np.zeros((0, 2))
    K = _f64(np.stack([np.asarray(p[2], np.float64) for p in problems]))
    T0 = _f64(np.stack([np.asarray(p[3], np.float64) for p in problems]))
    T = np.zeros((n, 7))
    outl = np.zeros(max(M, 1), np.uint8)
    ninl = np.zeros(n, np.int32)
    st = (LmStats * n)()
    self._chk(self.lib.svs_pose_only_lm(C.c_void_p(self.h), n, _p(off), _p(pts), _p(uv), _p(K), _p(T0), C.c_double(chi2_th),
                                        rounds, iters, _p(T), _p(outl), _p(ninl), st))
    return [(T[i].copy(), outl[off[i]:off[i + 1]].copy(), int(ninl[i]), st[i]) for i in range(n)]


def _ba_optimize(self, problems, K_left, K_right, ext_left, ext_right, huber_delta=5.991, max_iter=10, jac_mode=0):
    """problems: list of dicts(poses [N,7], lms [L,3], edge_kf, edge_lm, edge_cam, edge_uv).
    Returns list of (poses, lms, chi2, stats)."""
    n = len(problems)
    ko, lo, eo = (np.zeros(n + 1, np.int32) for _ in range(3))
    for i, p in enumerate(problems):
        ko[i + 1] = ko[i] + len(p["poses"]); lo[i + 1] = lo[i] + len(p["lms"]); eo[i + 1] = eo[i] + len(p["edge_kf"])
    cat = lambda k, dt, shp: np.ascontiguousarray(np.concatenate([np.asarray(p[k], dt).reshape(shp) for p in problems]), dt)
    poses, lms = cat("poses", np.float64, (-1, 7)), cat("lms", np.float64, (-1, 3))
    ekf, elm = cat("edge_kf", np.int32, (-1,)), cat("edge_lm", np.int32, (-1,))
    ecam, euv = cat("edge_cam", np.uint8, (-1,)), cat("edge_uv", np.float64, (-1, 2))
    chi2 = np.zeros(max(int(eo[-1]), 1))
    st = (BaStats * n)()
    self._chk(self.lib.svs_ba_optimize(
        C.c_void_p(self.h), n, _p(ko), _p(poses), _p(lo), _p(lms), _p(eo), _p(ekf), _p(elm), _p(ecam), _p(euv),
        _p(_f64(K_left)), _p(_f64(K_right)), _p(_f64(ext_left)), _p(_f64(ext_right)), C.c_double(huber_delta), max_iter,
        jac_mode, _p(chi2), st))
    return [(poses[ko[i]:ko[i + 1]].copy(), lms[lo[i]:lo[i + 1]].copy(), chi2[eo[i]:eo[i + 1]].copy(), st[i]) for i in range(n)]


def _stereo_bm(self, left, right, ndisp=128, block=15):
    left, right = _u8(left), _u8(right)
    single = left.ndim == 2
    if single:
        left, right = left[None], right[None]
    n, h, w = left.shape
    out = np.zeros((n, h, w), np.int16)
    self._chk(self.lib.svs_stereo_bm(C.c_void_p(self.h), _p(left), _p(right), w, h, w, n, C.c_size_t(w * h), ndisp, block, _p(out)))
    return out[0] if single else out


def _bgr2gray(self, bgr):
    bgr = _u8(bgr)
    single = bgr.ndim == 3
    if single:
        bgr = bgr[None]
    n, h, w, _ = bgr.shape
    out = np.zeros((n, h, w), np.uint8)
    self._chk(self.lib.svs_bgr2gray(C.c_void_p(self.h), _p(bgr), w, h, n, _p(out)))
    return out[0] if single else out


def _backproject(self, disp, bgr, K, baseline, cam_pose_inv, T_cw):
    disp = np.ascontiguousarray(disp, np.int16)
    bgr = _u8(bgr)
    h, w = disp.shape
    xyz = np.zeros((h * w, 3), np.float32)
    rgb = np.zeros((h * w, 3), np.uint8)
    n = C.c_int32(0)
    self._chk(self.lib.svs_backproject(C.c_void_p(self.h), _p(disp), _p(bgr), w, h, _p(_f64(K)), C.c_double(baseline),
                                       _p(_f64(cam_pose_inv)), _p(_f64(T_cw)), _p(xyz), _p(rgb), C.byref(n)))
    return xyz[:n.value].copy(), rgb[:n.value].copy()


Context.triangulate = _triangulate
def _pose_graph_optimize(self, poses, fixed, edge_a, edge_b, meas, max_iter=22, jac_mode=1):
    P = _f64(poses).reshape(-1, 7).copy()
    fx, ea, eb, M = _u8(fixed), _i32(edge_a), _i32(edge_b), _f64(meas).reshape(-1, 7)
    st = BaStats()
    self._chk(self.lib.svs_pose_graph_optimize(C.c_void_p(self.h), len(P), _p(P), _p(fx), len(ea), _p(ea), _p(eb), _p(M), int(max_iter),
                                               int(jac_mode), C.byref(st)))
    return P, st


def _pose_graph_move_landmarks(self, lms, lm_kf, old_poses, new_poses):
    L = _f64(lms).reshape(-1, 3).copy()
    k, po, pn = _i32(lm_kf), _f64(old_poses).reshape(-1, 7), _f64(new_poses).reshape(-1, 7)
    self._chk(self.lib.svs_pose_graph_move_landmarks(C.c_void_p(self.h), len(L), _p(L), _p(k), len(po), _p(po), _p(pn)))
    return L


def _pointcloud_sor(self, xyz, mean_k=50, stddev_mul=1.0):
    P = _f32(xyz).reshape(-1, 3)
    keep = np.zeros(len(P), np.uint8); md = np.zeros(len(P), np.float32)
    nk = C.c_int(0)
    self._chk(self.lib.svs_pointcloud_sor(C.c_void_p(self.h), _p(P), len(P), int(mean_k), C.c_double(stddev_mul), _p(keep), _p(md), C.byref(nk)))
    return keep.astype(bool), md


def _voxel_grid(self, xyz, rgb=None, leaf=0.02):
    P = _f32(xyz).reshape(-1, 3)
    col = _u8(rgb).reshape(-1, 3) if rgb is not None else None
    out = np.zeros((len(P), 3), np.float32)
    oc = np.zeros((len(P), 3), np.uint8) if col is not None else None
    n = C.c_int(0)
    self._chk(self.lib.svs_voxel_grid(C.c_void_p(self.h), _p(P), _p(col) if col is not None else None, len(P), C.c_double(leaf), _p(out),
                                      _p(oc) if oc is not None else None, C.byref(n)))
    return out[:n.value].copy(), (oc[:n.value].copy() if oc is not None else None)


Context.pointcloud_sor = _pointcloud_sor
Context.voxel_grid = _voxel_grid
Context.pose_graph_optimize = _pose_graph_optimize
Context.pose_graph_move_landmarks = _pose_graph_move_landmarks
Context.pose_only_lm = _pose_only_lm
Context.ba_optimize = _ba_optimize
Context.stereo_bm = _stereo_bm
Context.bgr2gray = _bgr2gray
Context.backproject = _backproject


# ---------------------------------------------------------------- pipeline (svs_slam_*)
class SlamConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "num_features", "num_features_init", "num_features_tracking", "num_features_tracking_bad",
        "num_features_needed_for_keyframe", "num_active_keyframes", "backend_on", "lk_win", "lk_max_level",
        "lk_max_iter", "ba_max_iter", "ba_jacobian_mode", "oracle_simd_granule")] + [(n, C.c_double) for n in (
        "max_triangulation_depth", "chi2_th", "gftt_quality", "gftt_min_distance", "lk_eps")] + [
        ("lazy_right_ingest", C.c_int32), ("device_tracking", C.c_int32)]


class Slam:
    """n_streams independent stereo streams stepped in lock-step: Frontend::AddFrame for a batch of streams."""

    def __init__(self, ctx, n_streams, in_w, in_h, K, baseline, half=True, **cfg):
        self.ctx, self.n = ctx, n_streams
        lib = ctx.lib
        lib.svs_slam_create.restype = C.c_void_p
        lib.svs_slam_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        lib.svs_slam_destroy.argtypes = [C.c_void_p]
        self.cfg = SlamConfig()
        lib.svs_slam_default_config(C.byref(self.cfg))
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise KeyError(k)
            setattr(self.cfg, k, v)
        K = _f64(K)
        self.h = lib.svs_slam_create(ctx.h, n_streams, in_w, in_h, int(bool(half)), C.byref(self.cfg), _p(K), baseline)
        if not self.h:
            raise SvsError("svs_slam_create failed: " + ctx.last_error())
        self.in_w, self.in_h = in_w, in_h
        self._lp = (C.c_void_p * n_streams)()
        self._rp = (C.c_void_p * n_streams)()
        self.poses = np.zeros((n_streams, 7))
        self.status = np.zeros(n_streams, np.int32)
        self.is_kf = np.zeros(n_streams, np.int32)
        self.inliers = np.zeros(n_streams, np.int32)
        self._pp, self._ps, self._pk, self._pi = _p(self.poses), _p(self.status), _p(self.is_kf), _p(self.inliers)

    def add_frames_ptrs(self, left_ptrs, right_ptrs, on_device=False, row_stride=None, next_left_ptrs=None, next_right_ptrs=None):
        """left_ptrs/right_ptrs: one address per stream (host pinned or device memory).  next_*: the frames of the
        FOLLOWING call (svs_slam_hint_next): their ingest overlaps this step's kernels."""
        if next_left_ptrs is not None:
            nl = (C.c_void_p * self.n)(*[int(x) for x in next_left_ptrs])
            nr = (C.c_void_p * self.n)(*[int(x) for x in next_right_ptrs])
            self.ctx._chk(self.ctx.lib.svs_slam_hint_next(C.c_void_p(self.h), nl, nr))
        for i in range(self.n):
            self._lp[i] = int(left_ptrs[i])
            self._rp[i] = int(right_ptrs[i])
        rc = self.ctx.lib.svs_slam_add_frames(C.c_void_p(self.h), self._lp, self._rp, C.c_size_t(row_stride or self.in_w),
                                              int(on_device), _p(self.poses), _p(self.status), _p(self.is_kf), _p(self.inliers))
        self.ctx._chk(rc)
        return self.poses

    @staticmethod
    def ptr_array(ptrs):
        """A reusable ctypes pointer array for add_frames_arrays (build once, pass every step)."""
        return (C.c_void_p * len(ptrs))(*[int(x) for x in ptrs])

    def add_frames_arrays(self, lp, rp, on_device, next_lp=None, next_rp=None, row_stride=None):
        """add_frames_ptrs with prebuilt ctypes arrays (Slam.ptr_array): no per-step Python work, so many pipelines can be
        stepped from Python threads without serialising on the interpreter lock."""
        lib, h = self.ctx.lib, C.c_void_p(self.h)
        if next_lp is not None:
            self.ctx._chk(lib.svs_slam_hint_next(h, next_lp, next_rp))
        self.ctx._chk(lib.svs_slam_add_frames(h, lp, rp, C.c_size_t(row_stride or self.in_w), int(on_device), self._pp, self._ps,
                                              self._pk, self._pi))
        return self.poses

    def add_frames(self, left, right):
        """left/right: uint8 [n_streams, in_h, in_w] host arrays."""
        left, right = _u8(left), _u8(right)
        assert left.shape == (self.n, self.in_h, self.in_w) == right.shape
        sz = self.in_w * self.in_h
        return self.add_frames_ptrs([left.ctypes.data + i * sz for i in range(self.n)],
                                    [right.ctypes.data + i * sz for i in range(self.n)])

    def features(self, stream, right=False, cap=4096):
        xy = np.zeros((cap, 2), np.float32)
        ids = np.zeros(cap, np.int64)
        valid = np.zeros(cap, np.uint8)
        n = C.c_int(0)
        self.ctx._chk(self.ctx.lib.svs_slam_get_features(C.c_void_p(self.h), stream, int(right), _p(xy), _p(ids), _p(valid), cap, C.byref(n)))
        return xy[:n.value].copy(), ids[:n.value].copy(), valid[:n.value].copy()

    def keyframes(self, stream, active_only=False, cap=100000):
        kid = np.zeros(cap, np.int64); fid = np.zeros(cap, np.int64); poses = np.zeros((cap, 7))
        n = C.c_int(0)
        self.ctx._chk(self.ctx.lib.svs_slam_get_keyframes(C.c_void_p(self.h), stream, int(active_only), _p(kid), _p(fid), _p(poses), cap, C.byref(n)))
        return kid[:n.value].copy(), fid[:n.value].copy(), poses[:n.value].copy()

    def landmarks(self, stream, active_only=False, cap=2000000):
        ids = np.zeros(cap, np.int64); xyz = np.zeros((cap, 3)); ot = np.zeros(cap, np.int32)
        n = C.c_int(0)
        self.ctx._chk(self.ctx.lib.svs_slam_get_landmarks(C.c_void_p(self.h), stream, int(active_only), _p(ids), _p(xyz), _p(ot), cap, C.byref(n)))
        return ids[:n.value].copy(), xyz[:n.value].copy(), ot[:n.value].copy()

    def counters(self):
        ph = np.zeros(8); cn = np.zeros(12, np.int64)
        self.ctx._chk(self.ctx.lib.svs_slam_get_counters(C.c_void_p(self.h), _p(ph), _p(cn)))
        names = ("push", "track_lk", "pose_lm", "detect", "right_lk", "triangulate", "ba", "host")
        cnames = ("frames", "keyframes", "ba_problems", "ba_iterations", "ba_trials", "ba_edges", "ba_lms", "ba_kfs", "lk_points", "pose_edges",
                  "h2d_image_bytes", "right_images")
        phases = dict(zip(names, ph.tolist()))
        hs = np.zeros(8)
        self.ctx._chk(self.ctx.lib.svs_slam_get_host_seconds(C.c_void_p(self.h), _p(hs)))
        hnames = ("host:begin+prep_track", "host:fin_track+prep_pose", "host:fin_pose+prep_detect", "host:fin_detect+prep_right",
                  "host:fin_right+prep_tri", "host:fin_tri+prep_ba", "host:fin_ba+end")
        phases.update(dict(zip(hnames, hs.tolist())))
        return phases, dict(zip(cnames, cn.tolist()))

    def set_threads(self, n):
        self.ctx._chk(self.ctx.lib.svs_slam_set_threads(C.c_void_p(self.h), int(n)))

    def close(self):
        if self.h:
            self.ctx.lib.svs_slam_destroy(C.c_void_p(self.h))
            self.h = None


Context.slam = lambda self, *a, **k: Slam(self, *a, **k)
