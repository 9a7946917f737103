"""GPU parity for the FP64 geometry (rows a4, a5, a7) and the dense path (row a10) through the C ABI against the
CPU oracle.  Tolerances: north_star asks poses / landmark XYZ within 1e-4 relative; engine (analytic) vs oracle
(analytic) is held to 1e-9 (reduction order only), engine vs oracle with g2o's numeric Jacobians to the 1e-4
relative-to-norm contract."""
import cv2
import numpy as np
import pytest

from oracle import cv_stages as o
from oracle import geom
from util import (K05, BASELINE, EXT_L, EXT_R, pose_problem, ba_problem, rel_to_norm, quat_to_R, stereo_pair, texture)

pytestmark = pytest.mark.gpu


def test_triangulate(ctx):
    rng = np.random.RandomState(2)
    n = 1000
    u, v = rng.uniform(150, 600, n), rng.uniform(5, 180, n)
    ur, vr = u - rng.uniform(1.5, 90, n), v + rng.randn(n) * 1.2
    ur[:5] = u[:5] + 3                      # negative disparity -> point behind the cameras
    xyz, ok = ctx.triangulate(np.stack([u, v], 1), np.stack([ur, vr], 1), K05, K05, BASELINE)
    wxyz, wok = geom.triangulate(np.stack([u, v], 1), np.stack([ur, vr], 1), K05, K05, BASELINE)
    assert 0 < wok.sum() < n
    assert np.array_equal(ok, wok)
    assert rel_to_norm(xyz, wxyz).max() < 1e-9
    assert len(ctx.triangulate(np.zeros((0, 2)), np.zeros((0, 2)), K05, K05, BASELINE)[0]) == 0


def test_pose_only_lm_batch(ctx):
    probs = [pose_problem(s, m=m) for s, m in [(0, 150), (1, 190), (2, 75), (3, 33), (4, 1), (5, 0), (6, 400)]]
    bad = pose_problem(7, m=40)
    probs.append((bad[0], bad[1] + 500.0, bad[2], bad[3], bad[4]))          # every edge an outlier
    res = ctx.pose_only_lm([(p[0], p[1], p[2], p[3]) for p in probs])
    for (T, outl, ninl, st), p in zip(res, probs):
        wT, woutl, wninl, wst = geom.pose_only_lm(p[0], p[1], p[2], p[3])
        assert np.array_equal(outl, woutl) and ninl == wninl
        # trial counts may differ near convergence (g2o has no convergence test: once converged the sign of rho is
        # rounding noise), the pose must not
        assert st.solves == st.trials and st.iterations <= 40 and (st.iterations > 0) == (wst.iterations > 0)
        assert np.abs(T - wT).max() < 1e-8 * max(1.0, np.abs(wT).max())


@pytest.mark.parametrize("n_kf,n_lm", [(10, 300), (20, 600), (4, 40)])
def test_ba_window_vs_oracle(ctx, n_kf, n_lm):
    probs = [ba_problem(s, n_kf=n_kf, n_lm=n_lm)[0] for s in (0, 1)]
    probs.append(ba_problem(4, n_kf=6, n_lm=80, unused_kf=True, unused_lm=3)[0])
    for mode, tol in ((0, 1e-9), (1, None)):
        res = ctx.ba_optimize(probs, K05, K05, EXT_L, EXT_R, jac_mode=mode)
        for (P, L, chi2, st), pr in zip(res, probs):
            wP, wL, wchi2, wst = geom.ba_optimize(pr["poses"], pr["lms"], pr["edge_kf"], pr["edge_lm"], pr["edge_cam"],
                                                  pr["edge_uv"], K05, K05, EXT_L, EXT_R, jac_mode=mode)
            cen = lambda Q: np.array([-quat_to_R(q[:4]).T @ q[4:] for q in Q])
            if mode == 0:
                assert (st.iterations, st.trials) == (wst.iterations, wst.trials)
                assert abs(st.chi2 - wst.chi2) < 1e-9 * wst.chi2 and abs(st.chi2_init - wst.chi2_init) < 1e-10 * wst.chi2_init
                assert np.abs(P - wP).max() < tol * 10 and rel_to_norm(L, wL).max() < tol
                assert np.abs(chi2 - wchi2).max() < 1e-7 * max(1.0, wchi2.max())
            else:
                # numeric Jacobians (g2o's default for this edge, the reference's actual mode).  The delta = 1e-9 quotient
                # makes the result reproducible only to ~1e-4 relative-to-norm: that IS the north_star
                # contract, asserted here on the median and the 95 % quantile.  Measured over 6 seeds x 4 problem sizes
                # (tests/tools/diag_numeric_ba.py on a B200): windows of 10-20 keyframes: landmark median <= 1.8e-5,
                # q95 <= 7.5e-5, camera centres <= 1.1e-3 m over a 10-20 m window (<= 8e-5 of its extent), chi2 <= 9e-6.
                rl = rel_to_norm(L, wL)
                dc = np.linalg.norm(cen(P) - cen(wP), axis=1).max() / max(1.0, np.linalg.norm(cen(wP), axis=1).max())
                if len(pr["poses"]) >= 10:
                    assert np.median(rl) < 1e-4 and np.quantile(rl, 0.95) < 1e-4, (np.median(rl), np.quantile(rl, 0.95))
                    assert dc < 1e-4, dc
                    assert abs(st.chi2 - wst.chi2) < 1e-4 * wst.chi2
                else:
                    # 4-6 keyframes, 40-80 landmarks and no gauge fixed: the oracle's OWN numeric and analytic results
                    # differ by up to 5e-4 here, i.e. the reference's output is not defined more finely than that — the
                    # engine must sit within 3x that spread
                    aP, aL, _, _ = geom.ba_optimize(pr["poses"], pr["lms"], pr["edge_kf"], pr["edge_lm"], pr["edge_cam"],
                                                    pr["edge_uv"], K05, K05, EXT_L, EXT_R, jac_mode=0)
                    spread = np.quantile(rel_to_norm(wL, aL), 0.95)
                    assert np.quantile(rl, 0.95) < 3 * spread + 1e-4, (np.quantile(rl, 0.95), spread)
                    assert abs(st.chi2 - wst.chi2) < 1e-3 * wst.chi2
    # inactive vertices stay untouched
    P, L, _, _ = res[2]
    assert np.array_equal(P[0], probs[2]["poses"][0]) and np.array_equal(L[-3:], probs[2]["lms"][-3:])


def test_ba_structure_built_on_device_equals_host_build(ctx, monkeypatch):
    """k_ba_build (window structure on the device, the default for landmark-major edge lists) and the host construction give
    the same lists, hence bit-identical results; an edge list in another order takes the host construction."""
    probs = [ba_problem(s, n_kf=n, n_lm=m)[0] for s, n, m in ((0, 10, 300), (1, 20, 500), (2, 3, 20), (3, 10, 2200))]
    probs.append(ba_problem(4, n_kf=6, n_lm=80, unused_kf=True, unused_lm=3)[0])
    dev = ctx.ba_optimize(probs, K05, K05, EXT_L, EXT_R)
    monkeypatch.setenv("SVS_BA_HOST_BUILD", "1")
    host = ctx.ba_optimize(probs, K05, K05, EXT_L, EXT_R)
    monkeypatch.delenv("SVS_BA_HOST_BUILD")
    for (P, L, chi2, st), (hP, hL, hchi2, hst) in zip(dev, host):
        assert np.array_equal(P, hP) and np.array_equal(L, hL) and np.array_equal(chi2, hchi2)
        assert (st.iterations, st.trials, st.chi2) == (hst.iterations, hst.trials, hst.chi2)
    # a shuffled edge list (same graph): host construction, same minimum
    pr = dict(probs[0])
    perm = np.random.RandomState(0).permutation(len(pr["edge_kf"]))
    for k in ("edge_kf", "edge_lm", "edge_cam", "edge_uv"):
        pr[k] = np.ascontiguousarray(np.asarray(pr[k])[perm])
    (P, L, chi2, st), = ctx.ba_optimize([pr], K05, K05, EXT_L, EXT_R)
    assert np.abs(P - dev[0][0]).max() < 1e-8 and rel_to_norm(L, dev[0][1]).max() < 1e-8
    assert np.abs(chi2 - dev[0][2][perm]).max() < 1e-6 * max(1.0, dev[0][2].max())


def test_ba_schedule_and_headroom_do_not_change_results(ctx):
    """svs_set_ba_schedule (own high-priority stream, forced CTA size) and svs_reserve_headroom (scratch buffers grown once to
    factor x their largest request, contents kept) are scheduling / allocation knobs: bit-identical results, and after the
    head room a call that needs up to factor x the memory of the largest earlier call does not reallocate."""
    import ctypes as C
    lib = ctx.lib
    lib.svs_buffer_regrowths.restype = C.c_longlong
    small = [ba_problem(s, n_kf=10, n_lm=300)[0] for s in (0, 1, 2)]
    base = ctx.ba_optimize(small, K05, K05, EXT_L, EXT_R)
    for prio, threads in ((1, 0), (1, 256), (0, 512), (0, 256)):
        ctx.set_ba_schedule(prio, threads)
        got = ctx.ba_optimize(small, K05, K05, EXT_L, EXT_R)
        for (P, L, chi2, st), (bP, bL, bchi2, bst) in zip(got, base):
            assert (st.iterations, st.trials) == (bst.iterations, bst.trials), (prio, threads)
            if threads == 256:      # another CTA size sums the per-thread chi2 partials in another order: last-bit differences
                assert np.abs(P - bP).max() < 1e-9 and rel_to_norm(L, bL).max() < 1e-9 and abs(st.chi2 - bst.chi2) <= 1e-9 * bst.chi2
            else:                   # the stream a kernel runs on cannot change its arithmetic
                assert np.array_equal(P, bP) and np.array_equal(L, bL) and np.array_equal(chi2, bchi2), (prio, threads)
                assert st.chi2 == bst.chi2
    ctx.set_ba_schedule(0, 0)
    with pytest.raises(Exception):
        ctx.set_ba_schedule(0, 300)
    ctx.reserve_headroom(2.5)
    r0 = lib.svs_buffer_regrowths()
    more = small + [ba_problem(s, n_kf=10, n_lm=300)[0] for s in (3, 4, 5)]      # twice the windows: inside the head room
    got = ctx.ba_optimize(more, K05, K05, EXT_L, EXT_R)
    assert lib.svs_buffer_regrowths() == r0
    for (P, L, chi2, st), (bP, bL, bchi2, bst) in zip(got[:3], base):
        assert np.array_equal(P, bP) and np.array_equal(L, bL) and np.array_equal(chi2, bchi2)
    with pytest.raises(Exception):
        ctx.reserve_headroom(0.5)


def test_ba_large_window_global_S(ctx):
    pr = ba_problem(9, n_kf=30, n_lm=500)[0]          # 180x180 reduced system: beyond shared memory
    (P, L, chi2, st), = ctx.ba_optimize([pr], K05, K05, EXT_L, EXT_R)
    wP, wL, wchi2, wst = geom.ba_optimize(pr["poses"], pr["lms"], pr["edge_kf"], pr["edge_lm"], pr["edge_cam"], pr["edge_uv"],
                                          K05, K05, EXT_L, EXT_R)
    assert (st.iterations, st.trials) == (wst.iterations, wst.trials)
    assert np.abs(P - wP).max() < 1e-8 and rel_to_norm(L, wL).max() < 1e-9


@pytest.mark.parametrize("h,w", [(188, 620), (185, 613), (64, 200), (370, 1226)])
def test_stereo_bm_bit_exact(ctx, h, w):
    pairs = [stereo_pair(h, w, s) for s in (h, h + 1)]
    l, r = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    got = ctx.stereo_bm(l, r, 128, 15)
    for i in range(2):
        assert np.array_equal(got[i], o.stereo_bm(l[i], r[i]))
        assert np.array_equal(got[i], cv2.StereoBM_create(128, 15).compute(l[i], r[i]))
    assert (got >= 0).sum() > 100


def test_stereo_bm_other_params(ctx):
    l, r = stereo_pair(100, 300, 3, dmin=1, dmax=50)
    for nd, bs in ((64, 9), (16, 5), (96, 21)):
        assert np.array_equal(ctx.stereo_bm(l, r, nd, bs), cv2.StereoBM_create(nd, bs).compute(l, r)), (nd, bs)
    tiny = texture(20, 100, 1)
    assert (ctx.stereo_bm(tiny, tiny, 128, 15) == -16).all()      # narrower than the valid region


def test_bgr2gray_and_backproject(ctx):
    rng = np.random.RandomState(0)
    bgr = rng.randint(0, 256, (2, 60, 200, 3), np.uint8)
    g = ctx.bgr2gray(bgr)
    for i in range(2):
        assert np.array_equal(g[i], cv2.cvtColor(bgr[i], cv2.COLOR_BGR2GRAY)) and np.array_equal(g[i], o.bgr2gray(bgr[i]))
    l, r = stereo_pair(64, 200, 5)
    disp = cv2.StereoBM_create(128, 15).compute(l, r)
    col = np.stack([l, l // 2, 255 - l], -1)
    T = geom.se3_exp(np.array([0.3, -0.1, 2.0, 0.02, -0.1, 0.01])).astype(np.float32).astype(np.float64)
    xyz, rgb = ctx.backproject(disp, col, K05, BASELINE, EXT_L, T)
    wxyz, wrgb = o.backproject(disp, col, K05, BASELINE, EXT_L, T, geom.se3_act, geom.se3_inv)
    assert len(wxyz) > 100 and xyz.shape == wxyz.shape
    assert np.array_equal(rgb, wrgb)
    assert np.abs(xyz - wxyz).max() <= 1e-5 * np.abs(wxyz).max()
