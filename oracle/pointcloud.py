"""CPU oracle — PCL's two dense-map filters restated in NumPy / scipy (TEST INFRASTRUCTURE ONLY).

The reference calls them right after the back-projection (reference src/dense_reconstruction.cpp:175-209).  PCL is a
third-party dependency that is NOT vendored under /root/reference and absent from this image (README.md:33 names PCL 1.12;
CMakeLists.txt:45 `find_package(PCL)`), so this restates the published PCL 1.12 algorithms — parity unpinned:

  pcl::StatisticalOutlierRemoval::applyFilterIndices (filters/impl/statistical_outlier_removal.hpp): (mean_k + 1)-NN query per
    point (the point itself is neighbour 0), distances[i] = float(sum_k sqrt(nn_dists[k]) / mean_k) with FLANN's float squared
    L2 (((dx*dx) + dy*dy) + dz*dz), mean / variance of the distances in double, threshold = mean + std_mul * stddev, a point
    is REMOVED when distance > threshold.
  pcl::VoxelGrid::applyFilter (filters/impl/voxel_grid.hpp): float leaf and inverse leaf, min_b = floor(min * inv),
    ijk = int(floor(p * inv) - float(min_b)), idx = ijk . (1, div_x, div_x * div_y), points sorted by idx, one centroid per
    voxel (pcl::CentroidPoint: float sums / float count, colours truncated); if the index range overflows int32 the input
    is returned unchanged.
"""
import numpy as np

F32 = np.float32


def sor_mean_distances(xyz, mean_k=50):
    from scipy.spatial import cKDTree
    P = np.ascontiguousarray(xyz, F32).reshape(-1, 3)
    n = len(P)
    fin = np.isfinite(P).all(1)
    out = np.zeros(n, F32)
    idx = np.flatnonzero(fin)
    Q = P[idx]
    tree = cKDTree(Q.astype(np.float64))
    kq = min(len(Q), 2 * mean_k + 8)              # candidates in double; the float distances are re-ranked below
    _, nb = tree.query(Q.astype(np.float64), k=kq)
    d = Q[:, None, :] - Q[nb]                     # float32
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(F32) + (d[..., 2] * d[..., 2]).astype(F32)
    d2 = np.sort(d2.astype(F32), axis=1)[:, :mean_k + 1]
    s = np.sqrt(d2[:, 1:].astype(F32)).astype(np.float64).sum(1)          # ascending, neighbour 0 (the point itself) dropped
    out[idx] = (s / mean_k).astype(F32)
    return out, int(fin.sum())


def sor(xyz, mean_k=50, std_mul=1.0):
    d, n_valid = sor_mean_distances(xyz, mean_k)
    dd = d.astype(np.float64)
    s, sq = 0.0, 0.0
    s = float(dd.sum())
    sq = float((d * d).astype(np.float64).sum())            # float product, double accumulation
    mean = s / n_valid
    var = (sq - s * s / n_valid) / (n_valid - 1.0)
    thr = mean + std_mul * np.sqrt(var)
    return ~(dd > thr), d, thr


def voxel_grid(xyz, rgb=None, leaf=0.02):
    P = np.ascontiguousarray(xyz, F32).reshape(-1, 3)
    fin = np.isfinite(P).all(1)
    if not fin.any():
        return P.copy(), (None if rgb is None else np.asarray(rgb, np.uint8).copy())
    mn, mx = P[fin].min(0), P[fin].max(0)
    inv = F32(1.0) / F32(leaf)
    dxyz = [int(np.int64(F32(mx[a] - mn[a]) * inv)) + 1 for a in range(3)]
    if dxyz[0] * dxyz[1] * dxyz[2] > np.iinfo(np.int32).max:
        return P.copy(), (None if rgb is None else np.asarray(rgb, np.uint8).copy())
    minb = np.floor(mn * inv).astype(np.int32)
    maxb = np.floor(mx * inv).astype(np.int32)
    div = maxb - minb + 1
    mul = np.array([1, div[0], div[0] * div[1]], np.int64)
    ijk = (np.floor(P[fin] * inv).astype(F32) - minb.astype(F32)).astype(np.int32)
    key = (ijk.astype(np.int64) * mul).sum(1)
    order = np.argsort(key, kind="stable")
    Pf = P[fin][order]
    key = key[order]
    C = None if rgb is None else np.asarray(rgb, np.uint8).reshape(-1, 3)[fin][order]
    starts = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    ends = np.r_[starts[1:], len(key)]
    out = np.zeros((len(starts), 3), F32)
    oc = None if C is None else np.zeros((len(starts), 3), np.uint8)
    for v, (a, b) in enumerate(zip(starts, ends)):
        acc = np.zeros(3, F32)
        for j in range(a, b):
            acc = (acc + Pf[j]).astype(F32)
        out[v] = acc / F32(b - a)
        if C is not None:
            ca = np.zeros(3, F32)
            for j in range(a, b):
                ca = (ca + C[j].astype(F32)).astype(F32)
            oc[v] = (ca / F32(b - a)).astype(np.uint32).astype(np.uint8)
    return out, oc
