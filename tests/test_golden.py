"""The oracle (CPU) and the CUDA path (GPU) against the committed golden vectors in tests/golden/golden_cv.npz
(generated from cv2 with the reference's arguments by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import cv_stages as o
from oracle import geom

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cv.npz"))


def _granule_of_fixture():
    for g in (32, 16, 8, 4, 64, 0):
        if np.array_equal(o.min_eig_map(G["img"], g).view(np.uint32), G["min_eig"].view(np.uint32)):
            return g
    raise AssertionError("no SIMD granule reproduces the golden min-eigenvalue map")


def test_oracle_matches_golden():
    g = _granule_of_fixture()
    assert np.array_equal(o.half_nearest(G["img"]), G["half"])
    assert np.array_equal(o.feature_mask(G["img"].shape, G["occ"]), G["mask"])
    xy, resp = o.gftt_detect(G["img"], G["mask"], 150, 0.01, 20, g)
    assert np.array_equal(xy, G["gftt_xy"]) and np.array_equal(resp, G["gftt_resp"])
    assert np.array_equal(o.pyr_down(G["img"]), G["pyr1"])
    q, st, _ = geom.lk_track(o.build_pyramid(G["lk_a"]), o.build_pyramid(G["lk_b"]), G["lk_p0"], G["lk_init"])
    assert np.array_equal(st, G["lk_status"])
    ok = st == 1
    assert np.abs(q[ok] - G["lk_p1"][ok]).max() < 1e-3
    assert np.array_equal(o.stereo_bm(G["bm_l"], G["bm_r"]), G["bm_disp"])
    assert np.array_equal(o.bgr2gray(G["bgr"]), G["gray"])


@pytest.mark.gpu
def test_cuda_matches_golden(ctx):
    g = _granule_of_fixture()
    assert np.array_equal(ctx.half_nearest(G["img"]), G["half"])
    assert np.array_equal(ctx.corner_min_eig(G["img"], g).view(np.uint32), G["min_eig"].view(np.uint32))
    xy, resp = ctx.gftt_detect(G["img"], occupied_xy=G["occ"], max_corners=150, granule=g)
    assert np.array_equal(xy, G["gftt_xy"]) and np.array_equal(resp, G["gftt_resp"])
    q, st = ctx.lk_track(G["lk_a"], G["lk_b"], G["lk_p0"], G["lk_init"])
    assert np.array_equal(st, G["lk_status"])
    ok = st == 1
    assert np.abs(q[ok] - G["lk_p1"][ok]).max() < 1e-3
    assert np.array_equal(ctx.stereo_bm(G["bm_l"], G["bm_r"], 128, 15), G["bm_disp"])
    assert np.array_equal(ctx.bgr2gray(G["bgr"]), G["gray"])


# ---------------------------------------------------------------------------------------------------------------------
# FP64 geometry: committed outputs of the C oracle (tests/golden/make_golden_geom.py).  g2o cannot be imported, so this
# pins the oracle build itself (any box, any compiler must reproduce the vectors) and the CUDA path against fixed numbers.
GG = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_geom.npz"))
_K05 = np.array([707.0912 * 0.5, 707.0912 * 0.5, 601.8873 * 0.5, 183.1104 * 0.5])
_EXT_L, _EXT_R, _BASE = np.array([0, 0, 0, 1, 0, 0, 0.0]), np.array([0, 0, 0, 1, -0.5371657, 0, 0.0]), 0.5371657


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def test_oracle_geometry_matches_golden():
    xyz, ok = geom.triangulate(GG["tri_l"], GG["tri_r"], _K05, _K05, _BASE)
    assert np.array_equal(ok, GG["tri_ok"]) and _rel(xyz[ok == 1], GG["tri_xyz"][ok == 1]) < 1e-12
    T, outl, ninl, st = geom.pose_only_lm(GG["po_pts"], GG["po_uv"], GG["po_K"], GG["po_T0"])
    assert np.array_equal(outl, GG["po_outlier"]) and ninl == int(GG["po_ninl"]) and _rel(T, GG["po_T"]) < 1e-10
    P, L, chi2, sb = geom.ba_optimize(GG["ba_poses"], GG["ba_lms"], GG["ba_edge_kf"], GG["ba_edge_lm"], GG["ba_edge_cam"], GG["ba_edge_uv"],
                                      _K05, _K05, _EXT_L, _EXT_R)
    assert (sb.iterations, sb.trials) == (int(GG["ba_stats"][0]), int(GG["ba_stats"][1]))
    assert abs(sb.chi2 - GG["ba_stats"][2]) < 1e-9 * GG["ba_stats"][2] and _rel(P, GG["ba_P"]) < 1e-9 and _rel(L, GG["ba_L"]) < 1e-9


@pytest.mark.gpu
def test_cuda_geometry_matches_golden(ctx):
    xyz, ok = ctx.triangulate(GG["tri_l"], GG["tri_r"], _K05, _K05, _BASE)
    assert np.array_equal(ok, GG["tri_ok"]) and _rel(xyz[ok == 1], GG["tri_xyz"][ok == 1]) < 1e-9
    (T, outl, ninl, st), = ctx.pose_only_lm([(GG["po_pts"], GG["po_uv"], GG["po_K"], GG["po_T0"])])
    assert np.array_equal(outl, GG["po_outlier"]) and ninl == int(GG["po_ninl"]) and _rel(T, GG["po_T"]) < 1e-8
    prob = {k[3:]: GG[k] for k in ("ba_poses", "ba_lms", "ba_edge_kf", "ba_edge_lm", "ba_edge_cam", "ba_edge_uv")}
    (P, L, chi2, sb), = ctx.ba_optimize([prob], _K05, _K05, _EXT_L, _EXT_R)
    assert (sb.iterations, sb.trials) == (int(GG["ba_stats"][0]), int(GG["ba_stats"][1]))
    assert abs(sb.chi2 - GG["ba_stats"][2]) < 1e-8 * GG["ba_stats"][2] and _rel(P, GG["ba_P"]) < 1e-7 and _rel(L, GG["ba_L"]) < 1e-7
    assert np.abs(chi2 - GG["ba_chi2"]).max() < 1e-6 * max(1.0, GG["ba_chi2"].max())


# ---------------------------------------------------------------------------------------------------------------------
# The REAL g2o: tests/golden/golden_g2o.npz is produced by tests/golden/g2o/make_golden_g2o.cpp, which runs the committed
# problems through the reference's own g2o_types.h on top of an installed g2o / Sophus / Eigen.  Those libraries do not exist
# in the build image, so the file is absent there and the g2o half of the oracle stays "parity unpinned" (DESIGN.md §3); where
# somebody has run the recipe the comparisons below turn that into a pin.
_G2O = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_g2o.npz")
_need_g2o = pytest.mark.skipif(not os.path.exists(_G2O), reason="golden_g2o.npz not generated (needs the real g2o: tests/golden/g2o/CMakeLists.txt)")


def _pg_problem():
    from util import pose_graph_problem
    return pose_graph_problem(40, n=40, loops=((-1, 2),))


@_need_g2o
def test_oracle_matches_real_g2o():
    R = np.load(_G2O)
    T, outl, ninl, st = geom.pose_only_lm(GG["po_pts"], GG["po_uv"], GG["po_K"], GG["po_T0"])
    assert np.array_equal(outl, R["po_outlier"]) and _rel(T, R["po_T"]) < 1e-8
    # Backend::Optimize: g2o differentiates EdgeProjection numerically (delta 1e-9) -> 1e-4 relative-to-norm is what is defined
    P, L, chi2, sb = geom.ba_optimize(GG["ba_poses"], GG["ba_lms"], GG["ba_edge_kf"], GG["ba_edge_lm"], GG["ba_edge_cam"], GG["ba_edge_uv"],
                                      _K05, _K05, _EXT_L, _EXT_R, jac_mode=1)
    rl = np.linalg.norm(L - R["ba_L"], axis=1) / np.linalg.norm(R["ba_L"], axis=1)
    assert np.median(rl) < 1e-4 and np.quantile(rl, 0.95) < 1e-4 and _rel(P, R["ba_P"]) < 1e-4
    pr = _pg_problem()
    Pg, _ = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 1)
    assert _rel(Pg, R["pg_P"]) < 1e-5


@_need_g2o
@pytest.mark.gpu
def test_cuda_matches_real_g2o(ctx):
    R = np.load(_G2O)
    (T, outl, ninl, st), = ctx.pose_only_lm([(GG["po_pts"], GG["po_uv"], GG["po_K"], GG["po_T0"])])
    assert np.array_equal(outl, R["po_outlier"]) and _rel(T, R["po_T"]) < 1e-7
    prob = {k[3:]: GG[k] for k in ("ba_poses", "ba_lms", "ba_edge_kf", "ba_edge_lm", "ba_edge_cam", "ba_edge_uv")}
    (P, L, chi2, sb), = ctx.ba_optimize([prob], _K05, _K05, _EXT_L, _EXT_R, jac_mode=1)
    rl = np.linalg.norm(L - R["ba_L"], axis=1) / np.linalg.norm(R["ba_L"], axis=1)
    assert np.median(rl) < 1e-4 and np.quantile(rl, 0.95) < 1e-4 and _rel(P, R["ba_P"]) < 1e-4
    pr = _pg_problem()
    Pg, _ = ctx.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 1)
    assert _rel(Pg, R["pg_P"]) < 1e-5
