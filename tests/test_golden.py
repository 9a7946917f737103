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
