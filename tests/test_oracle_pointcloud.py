"""CPU checks of the dense-map filter oracle (oracle/pointcloud.py — PCL 1.12's StatisticalOutlierRemoval and VoxelGrid
restated; PCL is un-vendored and absent, parity unpinned): the k-NN mean distances against a brute-force float
evaluation, the threshold rule, voxel-grid known answers and PCL's index-overflow pass-through."""
import numpy as np

from oracle import pointcloud as pc


def _cloud(n, seed):
    rng = np.random.RandomState(seed)
    a = np.stack([rng.uniform(-5, 5, n), np.full(n, 1.6) + 0.01 * rng.randn(n), rng.uniform(2, 30, n)], 1)      # ground
    b = np.stack([np.full(n // 2, 4.0) + 0.02 * rng.randn(n // 2), rng.uniform(-2, 1.6, n // 2), rng.uniform(2, 30, n // 2)], 1)   # wall
    o = rng.uniform(-8, 8, (n // 50, 3)) + np.array([0, -3, 15])                                                  # stray points
    return np.concatenate([a, b, o]).astype(np.float32)


def test_sor_mean_distance_matches_brute_force():
    P = _cloud(600, 1)
    d, nv = pc.sor_mean_distances(P, 10)
    assert nv == len(P)
    for i in (0, 17, 599, len(P) - 1):
        diff = P[i] - P
        d2 = ((diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]).astype(np.float32) + (diff[:, 2] * diff[:, 2]).astype(np.float32)).astype(np.float32)
        s = np.sort(d2)[1:11]
        want = np.float32(np.sqrt(s.astype(np.float32)).astype(np.float64).sum() / 10)
        assert d[i] == want


def test_sor_removes_the_stray_points():
    P = _cloud(3000, 2)
    keep, d, thr = pc.sor(P, 50, 1.0)
    n_surface = 3000 + 1500
    assert keep[:n_surface].mean() > 0.9 and keep[n_surface:].mean() < 0.2
    assert np.array_equal(keep, ~(d.astype(np.float64) > thr))
    Q = P.copy(); Q[5] = np.nan                          # a non-finite point gets distance 0 and is kept (PCL's behaviour)
    keep2, d2, _ = pc.sor(Q, 50, 1.0)
    assert d2[5] == 0 and keep2[5]


def test_voxel_grid_known_answers():
    P = np.array([[0.001, 0.001, 0.001], [0.011, 0.012, 0.013], [0.05, 0.0, 0.0], [0.051, 0.001, 0.002], [-0.03, 0.0, 0.0]], np.float32)
    C = np.array([[10, 20, 30], [20, 40, 61], [1, 2, 3], [4, 4, 4], [9, 9, 9]], np.uint8)
    out, oc = pc.voxel_grid(P, C, 0.02)
    assert len(out) == 3
    # ascending voxel index = ascending x here: (-0.03), (0.001 & 0.011), (0.05 & 0.051)
    assert np.allclose(out[0], P[4]) and np.allclose(out[1], (P[0] + P[1]) / 2, atol=1e-7) and np.allclose(out[2], (P[2] + P[3]) / 2, atol=1e-7)
    assert list(oc[1]) == [15, 30, 45] and list(oc[2]) == [2, 3, 3] and list(oc[0]) == [9, 9, 9]
    big = np.array([[0, 0, 0], [100, 100, 100]], np.float32)          # 5001^3 voxels > 2^31: PCL warns and returns the input
    out, _ = pc.voxel_grid(big, None, 0.02)
    assert np.array_equal(out, big)
