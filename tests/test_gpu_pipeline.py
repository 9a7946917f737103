"""GPU parity of the whole per-frame path (rows a0-a9): the batched C++/CUDA pipeline (svs_slam_*) against the CPU
oracle pipeline on the same synthetic KITTI-shaped stereo frames.  north_star bar: bit-exact keypoint indices /
counts, poses and landmark XYZ within 1e-4 relative."""
import numpy as np
import pytest

from oracle import pipeline as op
from svslam import synth
from util import rel_to_norm

pytestmark = pytest.mark.gpu

N_FRAMES = 36


@pytest.fixture(scope="module")
def clip():
    cor = synth.Corridor("kitti05", seed=0, n_frames=80)
    L, R, T = cor.sequence(N_FRAMES)
    return cor, L, R, T


def _sync(slam, b, o):
    """Teacher forcing: copy the engine's floating-point state into the oracle (discrete structure must already agree)."""
    xy, _, _ = slam.features(b)
    rxy, _, rvalid = slam.features(b, right=True)
    kid, _, kposes = slam.keyframes(b)
    lid, lxyz, _ = slam.landmarks(b)
    o.force_state(slam.poses[b], xy, rxy, rvalid, dict(zip(kid.tolist(), kposes)), dict(zip(lid.tolist(), lxyz)))


@pytest.mark.parametrize("backend_on", [1, 0])
def test_pipeline_lockstep_matches_oracle(ctx, granule, clip, backend_on):
    """Every frame starts from a bitwise-identical state (teacher forcing, see _sync): discrete results must be
    identical, floating-point results within the per-frame tolerances below.  (Free-running, the reference's own
    algorithm amplifies 1e-14 summation-order noise chaotically: LM without a convergence test, LK stopping at
    0.01 px — see test_pipeline_free_running_accuracy and DESIGN.md §6.)"""
    cor, L, R, T = clip
    B = 3            # three streams: two in phase, one delayed by 5 frames (different keyframe timing)
    delay = [0, 0, 5]
    slam = ctx.slam(B, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, backend_on=backend_on, oracle_simd_granule=granule)
    oracles = [op.Pipeline(cor.K_half(), cor.baseline, op.Cfg(backend_on=backend_on, granule=granule), stages="oracle") for _ in range(B)]
    try:
        n_kf = np.zeros(B, int)
        exact = total = 0
        worst = dict(xy=0.0, pose=0.0, lm=0.0)
        for i in range(N_FRAMES - 5):
            idx = [i + d for d in delay]
            poses = slam.add_frames(L[idx], R[idx]).copy()
            for b in range(B):
                want = oracles[b].add_frame(L[idx[b]], R[idx[b]])
                o = oracles[b]
                assert slam.status[b] == o.status, (i, b)
                assert bool(slam.is_kf[b]) == o.is_kf, (i, b)
                n_kf[b] += o.is_kf
                if i > 0:
                    assert slam.inliers[b] == o.tracking_inliers, (i, b)
                xy, ids, _ = slam.features(b)
                wxy, wids = o.current_features()
                assert len(xy) == len(wxy), (i, b)
                assert np.array_equal(ids, wids), (i, b)                     # identical counts, order, landmark links
                exact += np.array_equal(xy.view(np.uint32), wxy.view(np.uint32)); total += 1
                worst["xy"] = max(worst["xy"], float(np.abs(xy - wxy).max()))
                worst["pose"] = max(worst["pose"], float(np.abs(poses[b] - want).max() / max(1.0, np.abs(want[4:]).max())))
                assert np.abs(xy - wxy).max() < 2e-2, (i, b)                 # LK stops at 0.01 px
                assert np.abs(poses[b] - want).max() < 1e-5 * max(1.0, np.abs(want[4:]).max()), (i, b)
                lid, lxyz, lot = slam.landmarks(b)
                assert list(lid) == sorted(o.lms)
                wl = np.array([o.lms[k].pos for k in sorted(o.lms)]).reshape(-1, 3)
                worst["lm"] = max(worst["lm"], float(rel_to_norm(lxyz, wl).max()))
                assert rel_to_norm(lxyz, wl).max() < 1e-4, (i, b)
                assert list(lot) == [o.lms[k].observed_times for k in sorted(o.lms)]
                kid, fid, kposes = slam.keyframes(b)
                assert list(kid) == sorted(o.kfs) and list(fid) == [o.kfs[k].id for k in sorted(o.kfs)]
                akid, _, _ = slam.keyframes(b, active_only=True)
                assert list(akid) == sorted(o.active_kfs)
                alid, _, _ = slam.landmarks(b, active_only=True)
                assert list(alid) == sorted(o.active_lms)
                _sync(slam, b, o)
        assert (n_kf >= 3).all()
        print("bit-identical keypoint frames %d/%d; worst per-frame diffs %s" % (exact, total, worst))
        ph, cn = slam.counters()
        assert cn["frames"] == B * (N_FRAMES - 5) and cn["keyframes"] == n_kf.sum()
        if backend_on:
            assert cn["ba_problems"] == n_kf.sum() and cn["ba_iterations"] > 0
    finally:
        slam.close()


@pytest.mark.parametrize("backend_on,kw", [(1, {}), (0, {}), (1, dict(num_active_keyframes=3, num_features_needed_for_keyframe=110))])
def test_device_tracking_is_bit_identical_to_host_tracking(ctx, granule, clip, backend_on, kw):
    """device_tracking = 1 (Track()'s per-frame arithmetic on device-resident state, csrc/track.cu; the host classes see a
    stream only at keyframes) against device_tracking = 0 (every seam a host round trip): free-running, three streams with
    different keyframe phases, every output bit for bit."""
    cor, L, R, T = clip
    B, delay = 3, [0, 2, 5]
    a = ctx.slam(B, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, backend_on=backend_on, oracle_simd_granule=granule, device_tracking=1, **kw)
    b = ctx.slam(B, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, backend_on=backend_on, oracle_simd_granule=granule, device_tracking=0, **kw)
    try:
        nk = 0
        for i in range(N_FRAMES - 5):
            idx = [i + d for d in delay]
            pa = a.add_frames(L[idx], R[idx]).copy()
            pb = b.add_frames(L[idx], R[idx]).copy()
            assert np.array_equal(pa.view(np.uint64), pb.view(np.uint64)), i
            assert np.array_equal(a.status, b.status) and np.array_equal(a.is_kf, b.is_kf) and np.array_equal(a.inliers, b.inliers), i
            nk += int(a.is_kf.sum())
            for s in range(B):
                for right in (False, True):
                    fa, fb = a.features(s, right=right), b.features(s, right=right)
                    assert np.array_equal(fa[0].view(np.uint32), fb[0].view(np.uint32)) and np.array_equal(fa[1], fb[1]) and np.array_equal(fa[2], fb[2]), (i, s, right)
        assert nk >= 9
        for s in range(B):
            for x, y in zip(a.landmarks(s), b.landmarks(s)):
                assert np.array_equal(x, y)
            for x, y in zip(a.keyframes(s), b.keyframes(s)):
                assert np.array_equal(x, y)
        ca, cb = a.counters()[1], b.counters()[1]
        for k in ("frames", "keyframes", "ba_problems", "ba_iterations", "ba_trials", "ba_edges", "lk_points", "pose_edges"):
            assert ca[k] == cb[k], k
    finally:
        a.close(); b.close()


def test_pipeline_window_eviction_lockstep(ctx, granule, clip):
    """A small window (num_active_keyframes = 3) forces Map::RemoveOldKeyframe / CleanMap on every keyframe."""
    cor, L, R, T = clip
    slam = ctx.slam(1, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, num_active_keyframes=3, oracle_simd_granule=granule)
    o = op.Pipeline(cor.K_half(), cor.baseline, op.Cfg(num_active_keyframes=3, granule=granule), stages="oracle")
    try:
        for i in range(N_FRAMES):
            est = slam.add_frames(L[i:i + 1], R[i:i + 1])[0].copy()
            west = o.add_frame(L[i], R[i])
            assert slam.status[0] == o.status and bool(slam.is_kf[0]) == o.is_kf and (i == 0 or slam.inliers[0] == o.tracking_inliers), i
            assert np.abs(est - west).max() < 1e-5 * max(1.0, np.abs(west).max()), i
            akid, _, _ = slam.keyframes(0, active_only=True)
            assert list(akid) == sorted(o.active_kfs) and len(akid) <= 3
            alid, _, aot = slam.landmarks(0, active_only=True)
            assert list(alid) == sorted(o.active_lms)
            assert list(aot) == [o.active_lms[k].observed_times for k in sorted(o.active_lms)]
            _sync(slam, 0, o)
        assert len(slam.keyframes(0)[0]) >= 5
    finally:
        slam.close()


def test_pipeline_window20_lockstep(ctx, granule, clip):
    """BASELINE config 3 (num_active_keyframes = 20, an override of config/stereo_slam_configs/default.yaml:27): the
    whole pipeline, teacher-forced.  num_features_needed_for_keyframe is raised so that EVERY frame is a keyframe: the
    20-keyframe window is full after 20 frames and evicts on each of the remaining 16 (BA over ~10 k edges, a 120x120
    reduced system)."""
    cor, L, R, T = clip
    kw = dict(num_active_keyframes=20, num_features_needed_for_keyframe=1000)
    slam = ctx.slam(1, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, oracle_simd_granule=granule, **kw)
    o = op.Pipeline(cor.K_half(), cor.baseline, op.Cfg(granule=granule, **kw), stages="oracle")
    try:
        worst = 0.0
        for i in range(N_FRAMES):
            est = slam.add_frames(L[i:i + 1], R[i:i + 1])[0].copy()
            west = o.add_frame(L[i], R[i])
            assert slam.status[0] == o.status and bool(slam.is_kf[0]) == o.is_kf and (i == 0 or slam.inliers[0] == o.tracking_inliers), i
            assert np.abs(est - west).max() < 1e-5 * max(1.0, np.abs(west).max()), i
            xy, ids, _ = slam.features(0)
            wxy, wids = o.current_features()
            assert len(xy) == len(wxy) and np.array_equal(ids, wids), i
            akid, _, _ = slam.keyframes(0, active_only=True)
            assert list(akid) == sorted(o.active_kfs) and len(akid) <= 20
            alid, alxyz, aot = slam.landmarks(0, active_only=True)
            assert list(alid) == sorted(o.active_lms)
            assert list(aot) == [o.active_lms[k].observed_times for k in sorted(o.active_lms)]
            wl = np.array([o.active_lms[k].pos for k in sorted(o.active_lms)]).reshape(-1, 3)
            worst = max(worst, float(rel_to_norm(alxyz, wl).max()))
            assert rel_to_norm(alxyz, wl).max() < 1e-4, i
            _sync(slam, 0, o)
        assert len(slam.keyframes(0)[0]) == N_FRAMES and len(slam.keyframes(0, active_only=True)[0]) == 20
        ph, cn = slam.counters()
        assert cn["ba_problems"] == N_FRAMES and cn["ba_kfs"] >= 20 * 16
        print("window-20 lock-step: worst landmark rel-to-norm diff %.3g, BA edges per window at the end %d" % (worst, cn["ba_edges"] // N_FRAMES))
    finally:
        slam.close()


def test_pipeline_free_running_accuracy(ctx, granule, clip):
    """Without teacher forcing the two implementations drift apart chaotically but must be equally accurate:
    ATE against the generator's ground truth and the keyframe rate agree."""
    cor, L, R, T = clip
    slam = ctx.slam(1, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, oracle_simd_granule=granule)
    o = op.Pipeline(cor.K_half(), cor.baseline, op.Cfg(granule=granule), stages="oracle")
    try:
        est, west, nk, wnk = [], [], 0, 0
        for i in range(N_FRAMES):
            est.append(slam.add_frames(L[i:i + 1], R[i:i + 1])[0].copy()); nk += int(slam.is_kf[0])
            west.append(o.add_frame(L[i], R[i])); wnk += int(o.is_kf)
        ate, wate = synth.ate_rmse(est, T), synth.ate_rmse(west, T)
        print("ATE engine %.4f m, oracle %.4f m over %.1f m; keyframes %d / %d" % (ate, wate, 0.8 * N_FRAMES, nk, wnk))
        assert ate < 0.10 and wate < 0.10 and abs(ate - wate) < 0.03
        assert abs(nk - wnk) <= 1 and slam.status[0] == 1
    finally:
        slam.close()


def test_prefetch_hint_gives_identical_results(ctx, granule, clip):
    """svs_slam_hint_next (double-buffered ingest on the second stream) and the lazy right-eye ingest must not change a
    single bit of the results: two pipelines over the same pinned frames, one with hints + lazy right images, one that
    pushes both eyes of every frame without hints."""
    import torch
    cor, L, R, T = clip
    n = 14
    Lt, Rt = torch.from_numpy(L[:n + 1].copy()).pin_memory(), torch.from_numpy(R[:n + 1].copy()).pin_memory()
    img = cor.W * cor.H
    a = ctx.slam(2, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, oracle_simd_granule=granule)
    b = ctx.slam(2, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, oracle_simd_granule=granule, lazy_right_ingest=0)
    try:
        for i in range(n):
            # stream 1 lags one frame behind stream 0; mode alternates between zero-copy (2) and staged DMA (0)
            mode = 2 if (i // 4) % 2 == 0 else 0
            lp = [Lt.data_ptr() + i * img, Lt.data_ptr() + max(i - 1, 0) * img]
            rp = [Rt.data_ptr() + i * img, Rt.data_ptr() + max(i - 1, 0) * img]
            nl = [Lt.data_ptr() + (i + 1) * img, Lt.data_ptr() + i * img]
            nr = [Rt.data_ptr() + (i + 1) * img, Rt.data_ptr() + i * img]
            pa = a.add_frames_ptrs(lp, rp, on_device=mode, next_left_ptrs=nl, next_right_ptrs=nr).copy()
            pb = b.add_frames_ptrs(lp, rp, on_device=mode).copy()
            assert np.array_equal(pa, pb), i
            assert np.array_equal(a.status, b.status) and np.array_equal(a.is_kf, b.is_kf) and np.array_equal(a.inliers, b.inliers)
        for s in range(2):
            assert np.array_equal(a.features(s)[0], b.features(s)[0])
            assert np.array_equal(a.landmarks(s)[1], b.landmarks(s)[1])
        ca, cb = a.counters()[1], b.counters()[1]
        assert cb["right_images"] == 2 * n and 2 <= ca["right_images"] < cb["right_images"]      # only init / keyframe steps
        assert ca["h2d_image_bytes"] <= 0.8 * cb["h2d_image_bytes"]      # left eyes (+ one unconsumed prefetch) + a few right eyes
    finally:
        a.close(); b.close()


def test_save_slam_output(ctx, granule, clip, tmp_path):
    """keyframes.txt / landmarks.pcd (reference src/visual_odometry.cpp:198-310) written from a finished run."""
    from svslam import kitti
    cor, L, R, T = clip
    slam = ctx.slam(1, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, oracle_simd_granule=granule)
    try:
        est = [slam.add_frames(L[i:i + 1], R[i:i + 1])[0].copy() for i in range(20)]
        nk, nl = kitti.save_slam_output(slam, 0, str(tmp_path), "/data/sequences/05", 0)
        kid, fid, kposes = slam.keyframes(0)
        lines = open(str(tmp_path / "keyframes.txt")).read().splitlines()
        assert nk == len(kid) >= 2 and len(lines) == 2 + nk and [int(l.split()[0]) for l in lines[2:]] == list(fid)
        Tcw = np.array([[float(v) for v in l.split()[1:]] for l in lines[2:]]).reshape(-1, 3, 4)
        assert np.allclose(Tcw[:, :, 3], kposes[:, 4:], rtol=1e-5, atol=1e-5)
        pcd = open(str(tmp_path / "landmarks.pcd")).read().splitlines()
        assert pcd[6] == "WIDTH %d" % nl and len(pcd) == 11 + nl and nl == len(slam.landmarks(0)[0])
        c, _ = kitti.pose7_to_Twc(np.array(est))
        gt_c, _ = kitti.pose7_to_Twc(np.asarray(T)[:20])
        ate = kitti.ate_rmse(c, gt_c)
        assert ate < 0.10 and abs(ate - synth.ate_rmse(est, T[:20])) < 1e-9
    finally:
        slam.close()


def test_pipeline_config4_shape_lockstep(ctx, granule, clip):
    """BASELINE config 4 frontend shape: FULL-resolution processing (1226x370, no half-resolution resize), 2000 requested
    features with minDistance 5 (the deviation SURVEY.md §0 documents), BA window 10.  Teacher-forced like the test
    above: identical discrete results, poses within 1e-5 relative."""
    cor, L, R, T = clip
    kw = dict(num_features=2000, gftt_min_distance=5.0, num_features_needed_for_keyframe=800, num_features_init=200,
              num_features_tracking=200, num_features_tracking_bad=80)
    slam = ctx.slam(1, cor.W, cor.H, cor.K_full(), cor.baseline, half=False, oracle_simd_granule=granule, **kw)
    o = op.Pipeline(cor.K_full(), cor.baseline, op.Cfg(granule=granule, **kw), stages="oracle", half=False)
    try:
        nk = 0
        for i in range(8):
            est = slam.add_frames(L[i:i + 1], R[i:i + 1])[0].copy()
            west = o.add_frame(L[i], R[i])
            assert slam.status[0] == o.status and bool(slam.is_kf[0]) == o.is_kf, i
            assert i == 0 or slam.inliers[0] == o.tracking_inliers, i
            xy, ids, _ = slam.features(0, cap=8192)
            wxy, wids = o.current_features()
            assert len(xy) == len(wxy) and np.array_equal(ids, wids), i
            assert np.abs(xy - wxy).max() < 2e-2, i
            assert np.abs(est - west).max() < 1e-5 * max(1.0, np.abs(west[4:]).max()), i
            nk += int(o.is_kf)
            _sync(slam, 0, o)
        assert len(xy) > 700 and nk >= 2
    finally:
        slam.close()


def test_kitti_directory_end_to_end(ctx, clip, tmp_path):
    """The data formats either side of the path, end to end: a KITTI-layout sequence directory (calib.txt, image_0/,
    image_1/ PNGs, ground-truth poses file) -> KittiSequence -> engine -> keyframes.txt / landmarks.pcd -> ATE."""
    import cv2
    import json
    import os
    import subprocess
    import sys
    from svslam import kitti
    cor, L, R, T = clip
    d = str(tmp_path / "05")
    for cam, imgs in ((0, L), (1, R)):
        os.makedirs(os.path.join(d, "image_%d" % cam))
        for i in range(24):
            cv2.imwrite(os.path.join(d, "image_%d" % cam, "%06d.png" % i), imgs[i])
    f, cx, cy, b = cor.f, cor.cx, cor.cy, cor.baseline
    with open(os.path.join(d, "calib.txt"), "w") as fh:
        for i, tx in enumerate((0.0, -f * b, 0.0, -f * b)):
            fh.write("P%d: %.12e 0 %.12e %.12e 0 %.12e %.12e 0 0 0 1 0\n" % (i, f, cx, tx, f, cy))
    _, Twc = kitti.pose7_to_Twc(np.asarray(T)[:24])
    np.savetxt(str(tmp_path / "05.txt"), Twc.reshape(24, 12))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "run_kitti.py"), d, "--poses", str(tmp_path / "05.txt"),
                        "--out", str(tmp_path / "out")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["frames"] == 24 and not out["lost"] and out["keyframes"] >= 2 and out["landmarks"] > 100
    assert out["ate_rmse_m"] < 0.10 and out["rpe_trans_m"] < 0.05
    assert os.path.exists(str(tmp_path / "out" / "keyframes.txt")) and os.path.exists(str(tmp_path / "out" / "landmarks.pcd"))
