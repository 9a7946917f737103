"""Pin the CPU oracle's OpenCV-stage restatements against the importable cv2 (the third-party
library the reference calls), with the reference's exact arguments.  CPU only."""
import cv2
import numpy as np
import pytest

from oracle import cv_stages as o
from oracle import geom
from util import texture, moved_pair, lk_points, stereo_pair

SIZES = [(188, 620), (185, 613), (47, 100), (64, 333)]


@pytest.mark.parametrize("h,w", SIZES + [(376, 1241)])
def test_half_nearest(h, w):
    img = texture(h, w, 1)
    ref = cv2.resize(img, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST)   # src/dataset.cpp:128
    assert np.array_equal(ref, o.half_nearest(img))


@pytest.mark.parametrize("h,w", SIZES)
def test_min_eig_bit_exact(h, w, granule):
    img = texture(h, w, h + w)
    ref = cv2.cornerMinEigenVal(img, 3, ksize=3)
    got = o.min_eig_map(img, granule)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


def test_min_eig_thread_independent(granule):
    img = texture(188, 620, 3)
    try:
        for nt in (1, 8):
            cv2.setNumThreads(nt)
            assert np.array_equal(cv2.cornerMinEigenVal(img, 3, ksize=3), o.min_eig_map(img, granule))
    finally:
        cv2.setNumThreads(0)


@pytest.mark.parametrize("h,w", [(188, 620), (185, 613)])
@pytest.mark.parametrize("min_dist,n", [(20, 150), (5, 2000), (1.4, 500), (0.5, 300)])
def test_gftt_identical_list_and_order(h, w, min_dist, n, granule):
    img = texture(h, w, 7)
    rng = np.random.RandomState(7)
    pts = np.stack([rng.rand(40) * w, rng.rand(40) * h], 1).astype(np.float32)
    mask = o.feature_mask((h, w), pts)
    m2 = np.full((h, w), 255, np.uint8)     # what Frontend::DetectFeatures draws, src/frontend.cpp:42-47
    for (x, y) in pts:
        p0 = (int(np.rint(x - np.float32(10))), int(np.rint(y - np.float32(10))))
        p1 = (int(np.rint(x + np.float32(10))), int(np.rint(y + np.float32(10))))
        cv2.rectangle(m2, p0, p1, 0, -1)
    assert np.array_equal(mask, m2)
    kps = cv2.GFTTDetector_create(n, 0.01, min_dist).detect(img, mask)   # src/frontend.cpp:24,51
    xy, resp = o.gftt_detect(img, mask, n, 0.01, min_dist, granule)
    rxy = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    rr = np.array([k.response for k in kps], np.float32)
    assert len(kps) > 0
    assert np.array_equal(rxy, xy) and np.array_equal(rr, resp)


def test_gftt_empty_mask(granule):
    img = texture(64, 100, 2)
    xy, resp = o.gftt_detect(img, np.zeros((64, 100), np.uint8), 50, 0.01, 20, granule)
    kps = cv2.GFTTDetector_create(50, 0.01, 20).detect(img, np.zeros((64, 100), np.uint8))
    assert len(xy) == 0 == len(kps)


@pytest.mark.parametrize("h,w", SIZES + [(376, 1241)])
def test_pyr_down(h, w):
    a = texture(h, w, 5)
    for _ in range(3):
        b = cv2.pyrDown(a)
        assert np.array_equal(b, o.pyr_down(a))
        a = b


@pytest.mark.parametrize("h,w,seed", [(188, 620, 1), (185, 613, 2), (376, 1241, 3)])
def test_lk_status_identical_positions_close(h, w, seed):
    a, b = moved_pair(h, w, seed)
    p0, init = lk_points(a, 5)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(                       # src/frontend.cpp:353-357
        a, b, p0, init.copy(), winSize=(11, 11), maxLevel=3,
        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01), flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    q1, s1, _ = geom.lk_track(o.build_pyramid(a), o.build_pyramid(b), p0, init)
    st = st.ravel()
    assert np.array_equal(st, s1)
    both = st == 1
    assert both.sum() > 200 and (st == 0).sum() > 0
    # OpenCV accumulates the integer window sums in f32 lanes, the oracle exactly: tolerance 1e-3 px
    assert np.abs(p1[both] - q1[both]).max() < 1e-3
    assert (p1[both] == q1[both]).all(1).mean() > 0.95


def test_lk_small_image_level_count():
    a = texture(30, 40, 9)
    b = np.roll(a, 1, 1)
    assert len(o.build_pyramid(a)) == 2
    p0 = np.array([[10, 10], [20, 15], [35, 25], [5, 28]], np.float32)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(a, b, p0, p0 + 1, winSize=(11, 11), maxLevel=3, criteria=(3, 30, 0.01), flags=4)
    q1, s1, _ = geom.lk_track(o.build_pyramid(a), o.build_pyramid(b), p0, p0 + 1)
    assert np.array_equal(st.ravel(), s1)
    assert np.abs(p1 - q1)[s1 == 1].max() < 1e-3


@pytest.mark.parametrize("h,w", [(188, 620), (185, 613), (64, 200)])
def test_stereo_bm_bit_exact(h, w):
    l, r = stereo_pair(h, w, h)
    ref = cv2.StereoBM_create(128, 15).compute(l, r)            # src/dense_reconstruction.cpp:89,114
    got = o.stereo_bm(l, r)
    assert (ref >= 0).sum() > 100
    assert np.array_equal(ref, got)
