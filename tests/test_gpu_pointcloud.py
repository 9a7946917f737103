"""GPU parity of the dense-map post-processing (reference src/dense_reconstruction.cpp:175-209):
statistical outlier removal (k = 50, 1.0 sigma) and the 0.02 m voxel grid against oracle/pointcloud.py."""
import cv2
import numpy as np
import pytest

from oracle import geom
from oracle import pointcloud as pc
from util import K05, BASELINE, EXT_L, stereo_pair

pytestmark = pytest.mark.gpu


def _cloud(n, seed):
    rng = np.random.RandomState(seed)
    a = np.stack([rng.uniform(-5, 5, n), np.full(n, 1.6) + 0.01 * rng.randn(n), rng.uniform(2, 30, n)], 1)
    b = np.stack([np.full(n // 2, 4.0) + 0.02 * rng.randn(n // 2), rng.uniform(-2, 1.6, n // 2), rng.uniform(2, 30, n // 2)], 1)
    o = rng.uniform(-8, 8, (n // 50, 3)) + np.array([0, -3, 15])
    return np.concatenate([a, b, o]).astype(np.float32)


@pytest.mark.parametrize("n,k", [(3000, 50), (20000, 50), (700, 10), (60, 50)])
def test_sor_matches_oracle(ctx, n, k):
    P = _cloud(n, n)
    keep, d = ctx.pointcloud_sor(P, k, 1.0)
    wkeep, wd, thr = pc.sor(P, k, 1.0)
    assert np.array_equal(d.view(np.uint32), wd.view(np.uint32))          # same float arithmetic: bit-identical mean distances
    assert np.array_equal(keep, wkeep)                                     # identical index set
    assert 0 < keep.sum() < len(P)


def test_sor_on_a_back_projected_keyframe(ctx):
    """The real producer: StereoBM -> back-projection (the step right before the filter in the reference), ~20 k points with
    the strongly varying density of a depth map, duplicates, and a few non-finite points."""
    l, r = stereo_pair(120, 400, 5)
    disp = cv2.StereoBM_create(128, 15).compute(l, r)
    col = np.stack([l, l // 2, 255 - l], -1)
    T = geom.se3_exp(np.array([0.3, -0.1, 2.0, 0.02, -0.1, 0.01]))
    xyz, rgb = ctx.backproject(disp, col, K05, BASELINE, EXT_L, T)
    P = xyz.astype(np.float32).copy()
    assert len(P) > 5000
    P[10] = P[11]                                            # a duplicate
    P[100, 1] = np.nan; P[200, 0] = np.inf
    keep, d = ctx.pointcloud_sor(P, 50, 1.0)
    wkeep, wd, _ = pc.sor(P, 50, 1.0)
    assert np.array_equal(d.view(np.uint32), wd.view(np.uint32)) and np.array_equal(keep, wkeep)
    assert d[100] == 0 and keep[100]


@pytest.mark.parametrize("leaf", [0.02, 0.05, 0.5])
def test_voxel_grid_matches_oracle(ctx, leaf):
    P = _cloud(8000, 3)
    P[:, 2] *= 0.2                                           # keep the 0.02 m grid below 2^31 voxels
    rng = np.random.RandomState(1)
    C = rng.randint(0, 256, (len(P), 3)).astype(np.uint8)
    out, oc = ctx.voxel_grid(P, C, leaf)
    wout, woc = pc.voxel_grid(P, C, leaf)
    assert out.shape == wout.shape and len(out) < len(P)
    assert np.array_equal(out.view(np.uint32), wout.view(np.uint32))      # same keys, same (index-order) float sums
    assert np.array_equal(oc, woc)
    out2, _ = ctx.voxel_grid(P, None, leaf)
    assert np.array_equal(out2, out)


def test_voxel_grid_overflow_passes_the_cloud_through(ctx):
    P = _cloud(2000, 4)
    P[0] = [-300, 0, 0]; P[1] = [300, 50, 200]              # 30001 x 2600 x 10000 voxels of 2 cm: PCL's int32 index overflows
    out, _ = ctx.voxel_grid(P, None, 0.02)
    assert np.array_equal(out, P)
