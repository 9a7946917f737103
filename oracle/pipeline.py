"""CPU oracle — the reference's per-frame pipeline restated in Python on top of the oracle stages.

TEST INFRASTRUCTURE ONLY (checker for tests/, and the timed CPU baseline of bench.py).

Mirrors, with the synchronous-BA schedule and ascending-id iteration orders:
  Frontend::AddFrame / StereoInit / Track / InsertKeyframe ...   src/frontend.cpp:36-721
  Backend::Optimize (graph construction, chi2 post-pass)         src/backend.cpp:39-246
  Map::InsertKeyFrame / RemoveOldKeyframe / CleanMap             src/map.cpp:21-181
  MapPoint::AddObservation / RemoveObservation                   src/mappoint.cpp:22-78
  Camera projections                                             src/camera.cpp:28-86

stages = "cv2":    the OpenCV stages run through cv2 itself with the reference's exact arguments
                   (the real third-party code; this is what the CPU baseline times)
stages = "oracle": the OpenCV stages run through the integer-exact restatements (cv_stages.py / geom.c),
                   which is what the GPU engine is bit-identical to.
The g2o blocks always run through oracle/geom.c (g2o is not installable here; parity unpinned).
"""
import numpy as np

from . import cv_stages as cvs
from . import geom

F32 = np.float32


class Cfg:
    num_features = 150
    num_features_init = 50
    num_features_tracking = 50
    num_features_tracking_bad = 20
    num_features_needed_for_keyframe = 80
    max_triangulation_depth = 300.0
    num_active_keyframes = 10
    backend_on = 1
    chi2_th = 5.991
    gftt_quality = 0.01
    gftt_min_distance = 20.0
    lk_win = 11
    lk_max_level = 3
    lk_max_iter = 30
    lk_eps = 0.01
    ba_max_iter = 10
    ba_jacobian_mode = 0
    granule = 32

    def __init__(self, **kw):
        for k, v in kw.items():
            if not hasattr(Cfg, k):
                raise KeyError(k)
            setattr(self, k, v)


class Feature:
    __slots__ = ("x", "y", "mp", "outlier", "left", "frame")

    def __init__(self, frame, x, y, left=True):
        self.frame, self.x, self.y, self.left = frame, F32(x), F32(y), left
        self.mp = None
        self.outlier = False


class Frame:
    def __init__(self, fid, left, right):
        self.id, self.kf_id, self.is_kf = fid, 0, False
        self.pose = np.array([0, 0, 0, 1, 0, 0, 0.0])
        self.left, self.right = left, right
        self.fl, self.fr = [], []
        self.prev_kf = None
        self.pyr_l = self.pyr_r = None


class MapPoint:
    def __init__(self, mid):
        self.id, self.pos, self.obs, self.observed_times, self.is_outlier = mid, np.zeros(3), [], 0, False

    def add_obs(self, f):
        self.obs.append(f)
        self.observed_times += 1

    def remove_obs(self, f):
        for i, o in enumerate(self.obs):
            if o is f:
                del self.obs[i]
                if f.outlier:
                    f.mp = None
                self.observed_times -= 1
                break


class Pipeline:
    def __init__(self, K, baseline, cfg=None, stages="oracle", cv2=None, half=True):
        self.cfg = cfg or Cfg()
        self.K, self.b = np.asarray(K, float), float(baseline)
        self.ext_l = np.array([0, 0, 0, 1, 0, 0, 0.0])
        self.ext_r = np.array([0, 0, 0, 1, -self.b, 0, 0.0])
        self.stages, self.cv2, self.half = stages, cv2, half
        self.status = 0         # 0 INITING 1 GOOD 2 BAD 3 LOST
        self.cur = self.last = None
        self.rel_motion = np.array([0, 0, 0, 1, 0, 0, 0.0])
        self.kfs, self.active_kfs = {}, {}
        self.lms, self.active_lms = {}, {}
        self.n_frames = self.n_kf = self.n_mp = 0
        self.cur_kf = self.prev_kf = None
        self.tracking_inliers = 0
        self.is_kf = False
        self.stats = dict(ba_calls=0, ba_iterations=0)

    # ---- Camera (src/camera.cpp)
    def world2pixel(self, p, T, right):
        pc = geom.se3_act(self.ext_r if right else self.ext_l, geom.se3_act(T, p))
        return np.array([self.K[0] * pc[0] / pc[2] + self.K[2], self.K[1] * pc[1] / pc[2] + self.K[3]])

    # ---- stages
    def _resize(self, img):
        if not self.half:
            return img
        if self.stages == "cv2":
            return self.cv2.resize(img, None, fx=0.5, fy=0.5, interpolation=self.cv2.INTER_NEAREST)
        return cvs.half_nearest(img)

    def _detect(self, img, occupied):
        c = self.cfg
        mask = cvs.feature_mask(img.shape, occupied) if len(occupied) else None
        if self.stages == "cv2":
            if mask is None:
                mask = np.full(img.shape, 255, np.uint8)
            kps = self.cv2.GFTTDetector_create(c.num_features, c.gftt_quality, c.gftt_min_distance).detect(img, mask)
            return np.array([k.pt for k in kps], F32).reshape(-1, 2)
        return cvs.gftt_detect(img, mask, c.num_features, c.gftt_quality, c.gftt_min_distance, c.granule)[0]

    def _lk(self, a, b, pa, pb, prev_xy, init_xy):
        c = self.cfg
        if len(prev_xy) == 0:
            return np.zeros((0, 2), F32), np.zeros(0, np.uint8)
        if self.stages == "cv2":
            cv2 = self.cv2
            nxt, st, _ = cv2.calcOpticalFlowPyrLK(
                a, b, np.asarray(prev_xy, F32), np.array(init_xy, F32), winSize=(c.lk_win, c.lk_win), maxLevel=c.lk_max_level,
                criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, c.lk_max_iter, c.lk_eps), flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
            return nxt, st.ravel()
        nxt, st, _ = geom.lk_track(pa, pb, prev_xy, init_xy, c.lk_win, c.lk_max_iter, c.lk_eps)
        return nxt, st

    def _pyr(self, img):
        if self.stages == "cv2":
            return None
        return cvs.build_pyramid(img, self.cfg.lk_win, self.cfg.lk_max_level)

    # ---- Frontend::AddFrame
    def add_frame(self, left_full, right_full):
        f = Frame(self.n_frames, self._resize(left_full), self._resize(right_full))
        self.n_frames += 1
        f.pyr_l = self._pyr(f.left)
        self.cur = f
        self.is_kf = False
        if self.status == 0:
            self._stereo_init()
        elif self.status in (1, 2):
            self._track()
        if self.last is not None and self.last is not f and not self.last.is_kf:
            self.last.left = self.last.right = self.last.pyr_l = self.last.pyr_r = None
        self._before = self.last if self.status in (1, 2, 3) and self.last is not None and self.n_frames > 1 else None
        self.last = f
        return f.pose.copy()

    def _detect_features(self):
        f = self.cur
        occ = np.array([[p.x, p.y] for p in f.fl], F32).reshape(-1, 2)
        xy = self._detect(f.left, occ)
        for (x, y) in xy:
            f.fl.append(Feature(f, x, y))
        return len(xy)

    def _find_in_right(self):
        f = self.cur
        if f.pyr_r is None:
            f.pyr_r = self._pyr(f.right)
        prev = np.array([[p.x, p.y] for p in f.fl], F32).reshape(-1, 2)
        init = prev.copy()
        idx = [i for i, p in enumerate(f.fl) if p.mp is not None]
        if idx:
            init[idx] = geom.world2pixel_batch(f.pose, self.ext_r, self.K, np.array([f.fl[i].mp.pos for i in idx])).astype(F32)
        nxt, st = self._lk(f.left, f.right, f.pyr_l, f.pyr_r, prev, init)
        H, W = f.right.shape
        good = 0
        for i in range(len(f.fl)):
            x, y = nxt[i]
            if st[i] and y >= 0 and y < H and x >= 0 and x < W:
                f.fr.append(Feature(f, x, y, left=False))
                good += 1
            else:
                f.fr.append(None)
        return good

    def _triangulate(self, init):
        f = self.cur
        idx = [i for i in range(len(f.fl)) if f.fr[i] is not None and (init or f.fl[i].mp is None)]
        if not idx:
            return 0
        l = np.array([[f.fl[i].x, f.fl[i].y] for i in idx], F32)
        r = np.array([[f.fr[i].x, f.fr[i].y] for i in idx], F32)
        xyz, ok = geom.triangulate(l, r, self.K, self.K, self.b)
        Twc = geom.se3_inv(f.pose)
        cnt = 0
        for k, i in enumerate(idx):
            p = xyz[k]
            good = ok[k] and p[2] > 0 and (init or p[2] <= self.cfg.max_triangulation_depth)
            if not good:
                continue
            mp = MapPoint(self.n_mp)
            self.n_mp += 1
            mp.pos = p.copy() if init else geom.se3_act(Twc, p)
            mp.add_obs(f.fl[i]); mp.add_obs(f.fr[i])
            f.fl[i].mp = mp; f.fr[i].mp = mp
            self.lms[mp.id] = mp; self.active_lms[mp.id] = mp
            cnt += 1
        return cnt

    def _set_keyframe(self, f):
        f.is_kf = True
        f.kf_id = self.n_kf
        self.n_kf += 1
        self.is_kf = True

    def _stereo_init(self):
        self._detect_features()
        if self._find_in_right() < self.cfg.num_features_init:
            return False
        self.cur_kf = self.cur
        self._triangulate(True)
        self._set_keyframe(self.cur)
        self._map_insert_keyframe(self.cur)
        if self.cfg.backend_on:
            self._optimize()
        self.status = 1
        return True

    def _track(self):
        f, last, c = self.cur, self.last, self.cfg
        f.pose = geom.se3_mul(self.rel_motion, last.pose)
        # TrackLastFrame
        prev = np.array([[p.x, p.y] for p in last.fl], F32).reshape(-1, 2)
        init = prev.copy()
        idx = [i for i, p in enumerate(last.fl) if p.mp is not None]
        if idx:     # one batched projection (same arithmetic as world2pixel per point)
            init[idx] = geom.world2pixel_batch(f.pose, self.ext_l, self.K, np.array([last.fl[i].mp.pos for i in idx])).astype(F32)
        nxt, st = self._lk(last.left, f.left, last.pyr_l, f.pyr_l, prev, init)
        H, W = f.left.shape
        for i in range(len(last.fl)):
            if not st[i]:
                continue
            x, y = nxt[i]
            if y < 0 or y >= H or x < 0 or x >= W:
                continue
            nf = Feature(f, x, y)
            nf.mp = last.fl[i].mp
            f.fl.append(nf)
        # EstimateCurrentPose
        feats = [p for p in f.fl if p.mp is not None]
        pts = np.array([p.mp.pos for p in feats]).reshape(-1, 3)
        uv = np.array([[p.x, p.y] for p in feats], np.float64).reshape(-1, 2)
        T, outl, ninl, _ = geom.pose_only_lm(pts, uv, self.K, f.pose)
        f.pose = T
        for p, o in zip(feats, outl):
            if o:
                p.mp = None
                p.outlier = False
        self.tracking_inliers = ninl
        self.status = 1 if ninl > c.num_features_tracking else (2 if ninl > c.num_features_tracking_bad else 3)
        # InsertKeyframe
        if ninl < c.num_features_needed_for_keyframe:
            self._set_keyframe(f)
            self._map_insert_keyframe(f)
            self.prev_kf, self.cur_kf = self.cur_kf, f
            f.prev_kf = self.prev_kf
            for p in f.fl:
                if p.mp is not None:
                    p.mp.add_obs(p)
            self._detect_features()
            self._find_in_right()
            self._triangulate(False)
            if c.backend_on:
                self._optimize()
        self.rel_motion = geom.se3_mul(f.pose, geom.se3_inv(last.pose))

    # ---- Map
    def _map_insert_keyframe(self, f):
        self.kfs[f.kf_id] = f
        self.active_kfs[f.kf_id] = f
        if len(self.active_kfs) > self.cfg.num_active_keyframes:
            self._remove_old_keyframe(f)

    def _remove_old_keyframe(self, cur):
        max_dis, min_dis, max_id, min_id = 0.0, 999999.0, 0, 0
        Twc = geom.se3_inv(cur.pose)
        for kid in sorted(self.active_kfs):
            kf = self.active_kfs[kid]
            if kf is cur:
                continue
            dis = float(np.linalg.norm(geom.se3_log(geom.se3_mul(kf.pose, Twc))))
            if dis > max_dis:
                max_dis, max_id = dis, kid
            if dis < min_dis:
                min_dis, min_id = dis, kid
        rem = self.active_kfs[min_id if min_dis < 0.2 else max_id]
        del self.active_kfs[rem.kf_id]
        for p in rem.fl:
            if p.mp is not None:
                p.mp.remove_obs(p)
        for p in rem.fr:
            if p is not None and p.mp is not None:
                p.mp.remove_obs(p)
        for mid in [m for m in self.active_lms if self.active_lms[m].observed_times == 0]:
            del self.active_lms[mid]
        rem.left = rem.right = rem.pyr_l = rem.pyr_r = None

    # ---- Backend::Optimize
    def build_ba_problem(self):
        kf_ids = sorted(self.active_kfs)
        kmap = {k: i for i, k in enumerate(kf_ids)}
        poses = np.array([self.active_kfs[k].pose for k in kf_ids]).reshape(-1, 7)
        lm_ids, lms, ekf, elm, ecam, euv, efeat = [], [], [], [], [], [], []
        for mid in sorted(self.active_lms):
            mp = self.active_lms[mid]
            if mp.is_outlier:
                continue
            li = -1
            for ft in mp.obs:
                if ft.outlier:
                    continue
                if li < 0:
                    li = len(lm_ids)
                    lm_ids.append(mid)
                    lms.append(mp.pos.copy())
                if not ft.frame.is_kf or ft.frame.kf_id not in kmap or self.active_kfs.get(ft.frame.kf_id) is not ft.frame:
                    continue
                ekf.append(kmap[ft.frame.kf_id]); elm.append(li); ecam.append(0 if ft.left else 1)
                euv.append([float(ft.x), float(ft.y)]); efeat.append(ft)
        return dict(kf_ids=kf_ids, lm_ids=lm_ids, poses=poses, lms=np.array(lms).reshape(-1, 3),
                    edge_kf=np.array(ekf, np.int32), edge_lm=np.array(elm, np.int32), edge_cam=np.array(ecam, np.uint8),
                    edge_uv=np.array(euv).reshape(-1, 2), feats=efeat)

    def _optimize(self):
        pr = self.build_ba_problem()
        c = self.cfg
        if len(pr["edge_kf"]):
            poses, lms, chi2, st = geom.ba_optimize(pr["poses"], pr["lms"], pr["edge_kf"], pr["edge_lm"], pr["edge_cam"],
                                                    pr["edge_uv"], self.K, self.K, self.ext_l, self.ext_r, c.chi2_th,
                                                    c.ba_max_iter, c.ba_jacobian_mode)
            self.stats["ba_calls"] += 1
            self.stats["ba_iterations"] += st.iterations
        else:
            poses, lms, chi2 = pr["poses"], pr["lms"], np.zeros(0)
        th, it = c.chi2_th, 0
        while it < 5:
            n_out = int((chi2 > th).sum())
            n_in = len(chi2) - n_out
            ratio = n_in / float(n_in + n_out) if (n_in + n_out) else float("nan")
            if ratio > 0.5:
                break
            th *= 2
            it += 1
        for ft, c2 in zip(pr["feats"], chi2):
            if c2 > th:
                ft.outlier = True
                if ft.mp is not None:
                    ft.mp.remove_obs(ft)
            else:
                ft.outlier = False
        for k, kid in enumerate(pr["kf_ids"]):
            self.active_kfs[kid].pose = poses[k].copy()
        for l, mid in enumerate(pr["lm_ids"]):
            self.lms[mid].pos = lms[l].copy()

    # ---- introspection for parity tests
    def current_features(self):
        f = self.cur
        xy = np.array([[p.x, p.y] for p in f.fl], F32).reshape(-1, 2)
        ids = np.array([p.mp.id if p.mp is not None else -1 for p in f.fl], np.int64)
        return xy, ids

    def force_state(self, cur_pose, left_xy, right_xy, right_valid, kf_poses, lm_xyz):
        """Teacher forcing for lock-step parity tests: overwrite the floating-point state (never the discrete
        structure) with another implementation's values so that the next frame starts from identical numbers."""
        f = self.cur
        f.pose = np.array(cur_pose, float)
        assert len(left_xy) == len(f.fl)
        for p, (x, y) in zip(f.fl, left_xy):
            p.x, p.y = F32(x), F32(y)
        if len(f.fr):
            assert len(right_xy) == len(f.fr)
            for p, (x, y), v in zip(f.fr, right_xy, right_valid):
                assert (p is not None) == bool(v)
                if p is not None:
                    p.x, p.y = F32(x), F32(y)
        for kid, pose in kf_poses.items():
            self.kfs[kid].pose = np.array(pose, float)
        for mid, pos in lm_xyz.items():
            self.lms[mid].pos = np.array(pos, float)
        if getattr(self, "_before", None) is not None and self._before is not f:
            self.rel_motion = geom.se3_mul(f.pose, geom.se3_inv(self._before.pose))     # src/frontend.cpp:685
