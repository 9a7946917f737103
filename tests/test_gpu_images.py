"""GPU parity for the image stages (rows a0, a1, a2/a3) through the C ABI, against the CPU oracle
and — where the third-party routine is importable — against cv2 itself."""
import cv2
import numpy as np
import pytest

from oracle import cv_stages as o
from oracle import geom
from util import texture, moved_pair, lk_points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w", [(376, 1241), (370, 1226), (47, 100), (65, 333), (6, 7)])
def test_half_nearest(ctx, h, w):
    imgs = np.stack([texture(h, w, s, rects=3) for s in range(3)])
    got = ctx.half_nearest(imgs)
    for i in range(3):
        assert np.array_equal(got[i], o.half_nearest(imgs[i]))
        assert np.array_equal(got[i], cv2.resize(imgs[i], None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST))


@pytest.mark.parametrize("h,w", [(188, 620), (185, 613), (47, 100), (64, 333), (376, 1241), (3, 3), (5, 29)])
@pytest.mark.parametrize("g", [None, 0, 16])
def test_min_eig_bit_exact(ctx, granule, h, w, g):
    gg = granule if g is None else g
    img = texture(h, w, h + w, rects=5)
    got = ctx.corner_min_eig(img, gg)
    want = o.min_eig_map(img, gg)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if g is None:
        assert np.array_equal(got.view(np.uint32), cv2.cornerMinEigenVal(img, 3, ksize=3).view(np.uint32))


@pytest.mark.parametrize("h,w", [(188, 620), (185, 613), (376, 1241)])
@pytest.mark.parametrize("min_dist,n", [(20, 150), (5, 2000), (1.4, 500), (0.5, 300), (7.3, 400)])
def test_gftt_identical(ctx, granule, h, w, min_dist, n):
    img = texture(h, w, 7)
    rng = np.random.RandomState(7)
    pts = np.stack([rng.rand(40) * w, rng.rand(40) * h], 1).astype(np.float32)
    pts[0] = (0.5, 0.5); pts[1] = (w - 1, h - 1); pts[2] = (w + 30, 5); pts[3] = (10.5, 20.5)
    mask = o.feature_mask((h, w), pts)
    # (1) mask given as box centres (the drop-in form), (2) mask given as an image
    xy1, r1 = ctx.gftt_detect(img, occupied_xy=pts, max_corners=n, min_distance=min_dist, granule=granule)
    xy2, r2 = ctx.gftt_detect(img, mask=mask, max_corners=n, min_distance=min_dist, granule=granule)
    wxy, wr = o.gftt_detect(img, mask, n, 0.01, min_dist, granule)
    assert len(wxy) > 0
    assert np.array_equal(xy1, wxy) and np.array_equal(r1, wr)
    assert np.array_equal(xy2, wxy) and np.array_equal(r2, wr)
    kps = cv2.GFTTDetector_create(n, 0.01, min_dist).detect(img, mask)
    assert np.array_equal(np.array([k.pt for k in kps], np.float32).reshape(-1, 2), xy1)
    assert np.array_equal(np.array([k.response for k in kps], np.float32), r1)


def test_gftt_edge_cases(ctx, granule):
    img = texture(64, 100, 2)
    xy, _ = ctx.gftt_detect(img, mask=np.zeros((64, 100), np.uint8), max_corners=50, granule=granule)
    assert len(xy) == 0                                   # everything masked
    flat = np.full((40, 60), 77, np.uint8)
    xy, _ = ctx.gftt_detect(flat, max_corners=50, granule=granule)
    assert len(xy) == len(o.gftt_detect(flat, None, 50, 0.01, 20, granule)[0]) == 0
    noise = np.random.RandomState(0).randint(0, 256, (120, 160), np.uint8)   # many candidates
    xy, r = ctx.gftt_detect(noise, max_corners=5000, min_distance=2, granule=granule)
    wxy, wr = o.gftt_detect(noise, None, 5000, 0.01, 2, granule)
    assert np.array_equal(xy, wxy) and np.array_equal(r, wr)
    big_noise = np.random.RandomState(1).randint(0, 256, (376, 1241), np.uint8)   # > smem sort capacity
    xy, r = ctx.gftt_detect(big_noise, max_corners=3000, min_distance=3, granule=granule)
    wxy, wr = o.gftt_detect(big_noise, None, 3000, 0.01, 3, granule)
    assert np.array_equal(xy, wxy) and np.array_equal(r, wr)


def test_frameset_pyramids_and_batched_gftt(ctx, granule):
    B, H, W = 5, 370, 1226
    left = np.stack([texture(H, W, 10 + s) for s in range(B)])
    right = np.stack([texture(H, W, 20 + s) for s in range(B)])
    fs = ctx.frameset(B, W, H, half=True)
    try:
        fs.push(left, right)
        assert (fs.w, fs.hgt, fs.n_levels) == (613, 185, 4)
        for b in (0, B - 1):
            for which, src in ((0, left), (2, right)):
                pyr = o.build_pyramid(o.half_nearest(src[b]))
                for lvl in range(4):
                    assert np.array_equal(fs.download(b, which, lvl), pyr[lvl]), (b, which, lvl)
        left2 = np.stack([texture(H, W, 30 + s) for s in range(B)])
        fs.push(left2, right)
        assert np.array_equal(fs.download(1, 1, 2), o.build_pyramid(o.half_nearest(left[1]))[2])   # previous
        assert np.array_equal(fs.download(1, 0, 0), o.half_nearest(left2[1]))                        # current
        rng = np.random.RandomState(3)
        sel = [3, 0, 4]
        occ = [np.stack([rng.rand(k) * 613, rng.rand(k) * 185], 1).astype(np.float32) for k in (25, 0, 60)]
        res = ctx.gftt_detect_batch(fs, sel, occ, max_corners=150, granule=granule)
        for (xy, r), s, oc in zip(res, sel, occ):
            img = o.half_nearest(left2[s])
            wxy, wr = o.gftt_detect(img, o.feature_mask(img.shape, oc), 150, 0.01, 20, granule)
            assert np.array_equal(xy, wxy) and np.array_equal(r, wr)
    finally:
        fs.close()


def test_tma_staged_kernels_over_many_tiles(ctx, granule):
    """The TMA-staged persistent kernels (k_pyr_down_tma, k_corner_response_tma) with more tiles than resident CTAs, so
    that every CTA walks several tiles through its two shared-memory buffers (mbarrier phase flips, buffer reuse), on
    full-resolution frames (processed at 1241x376: 20 x 24 corner tiles and 10 x 12 pyramid tiles per image) — every pyramid
    level and the GFTT result of every stream against the oracle."""
    B, H, W = 24, 376, 1241
    base = [texture(H, W, 70 + s) for s in range(4)]
    rng = np.random.RandomState(9)
    left = np.stack([np.roll(base[s % 4], (3 * s, 17 * s), (0, 1)) for s in range(B)])        # 24 distinct images, cheaply
    right = np.stack([np.roll(base[(s + 1) % 4], (5 * s, 11 * s), (0, 1)) for s in range(B)])
    fs = ctx.frameset(B, W, H, half=False)
    try:
        fs.push(left, right)
        fs.push(right, left)                                   # second push: the other buffers and tensor maps
        for b in range(B):
            for which, src in ((0, right), (1, left), (2, left)):
                if b % 5 and which:
                    continue
                pyr = o.build_pyramid(src[b])
                for lvl in range(fs.n_levels):
                    assert np.array_equal(fs.download(b, which, lvl), pyr[lvl]), (b, which, lvl)
        sel = list(range(B))
        occ = [np.stack([rng.rand(k) * W, rng.rand(k) * H], 1).astype(np.float32) for k in [0, 30, 80, 5] * (B // 4)]
        res = ctx.gftt_detect_batch(fs, sel, occ, max_corners=400, min_distance=12.0, granule=granule)
        for (xy, r), s_, oc in zip(res, sel, occ):
            if s_ % 3:
                continue
            wxy, wr = o.gftt_detect(right[s_], o.feature_mask((H, W), oc) if len(oc) else None, 400, 0.01, 12.0, granule)
            assert len(wxy) > 100 and np.array_equal(xy, wxy) and np.array_equal(r, wr), s_
    finally:
        fs.close()


@pytest.mark.parametrize("H,W,pad", [(370, 1226, 0), (376, 1241, 3), (47, 101, 1)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_frameset_ingest_modes(ctx, H, W, pad, mode):
    """svs_frameset_push_ptrs: staged DMA (0), device-resident (1) and zero-copy pinned-host (2, the persistent small-grid
    PCIe reader) ingest give the same half-resolution image and pyramid as the oracle, for aligned and unaligned rows."""
    import torch
    B = 6
    stride = W + pad
    buf = np.zeros((2, B, H, stride), np.uint8)
    for e in range(2):
        for b in range(B):
            buf[e, b, :, :W] = texture(H, W, 100 * e + b + H)
    t = torch.from_numpy(buf)
    t = t.cuda() if mode == 1 else (t.pin_memory() if mode == 2 else t)
    base = t.data_ptr()
    lp = [base + (0 * B + b) * H * stride for b in range(B)]
    rp = [base + (1 * B + b) * H * stride for b in range(B)]
    fs = ctx.frameset(B, W, H, half=True)
    try:
        for rep in range(2):     # twice: current/previous double buffering
            fs.push_ptrs(lp, rp, mode, row_stride=stride)
        for b in (0, 3, B - 1):
            for which, e in ((0, 0), (1, 0), (2, 1)):
                pyr = o.build_pyramid(o.half_nearest(buf[e, b, :, :W]))
                for lvl in range(fs.n_levels):
                    assert np.array_equal(fs.download(b, which, lvl), pyr[lvl]), (b, which, lvl)
    finally:
        fs.close()
    del t


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_frameset_prefetch(ctx, mode):
    """svs_frameset_prefetch_ptrs (double-buffered ingest on the second stream): a push of the prefetched pointers only
    rotates buffers; a push of other pointers discards the prefetch.  Current / previous / right buffers always hold what
    a plain sequence of pushes would give."""
    import torch
    B, H, W, T = 3, 370, 1226, 5
    buf = np.stack([np.stack([np.stack([texture(H, W, 1000 * e + 10 * t + b) for b in range(B)]) for t in range(T)]) for e in range(2)])
    t_ = torch.from_numpy(buf)
    t_ = t_.cuda() if mode == 1 else (t_.pin_memory() if mode == 2 else t_)
    base = t_.data_ptr()
    ptr = lambda e, t: [base + ((e * T + t) * B + b) * H * W for b in range(B)]
    want = lambda e, t, b, lvl: o.build_pyramid(o.half_nearest(buf[e, t, b]))[lvl]
    fs = ctx.frameset(B, W, H, half=True)
    try:
        fs.push_ptrs(ptr(0, 0), ptr(1, 0), mode)
        fs.prefetch_ptrs(ptr(0, 1), ptr(1, 1), mode)
        assert np.array_equal(fs.download(1, 0, 1), want(0, 0, 1, 1))       # current is still frame 0 while 1 is in flight
        fs.push_ptrs(ptr(0, 1), ptr(1, 1), mode)                            # hit
        fs.prefetch_ptrs(ptr(0, 2), ptr(1, 2), mode)
        for b in range(B):
            for lvl in range(fs.n_levels):
                assert np.array_equal(fs.download(b, 0, lvl), want(0, 1, b, lvl))
                assert np.array_equal(fs.download(b, 1, lvl), want(0, 0, b, lvl))
                assert np.array_equal(fs.download(b, 2, lvl), want(1, 1, b, lvl))
        fs.push_ptrs(ptr(0, 3), ptr(1, 3), mode)                            # miss: prefetch of frame 2 is discarded
        assert np.array_equal(fs.download(2, 0, 0), want(0, 3, 2, 0))
        assert np.array_equal(fs.download(2, 1, 2), want(0, 1, 2, 2))
        assert np.array_equal(fs.download(2, 2, 3), want(1, 3, 2, 3))
        fs.prefetch_ptrs(ptr(0, 4), ptr(1, 4), mode)
        fs.prefetch_ptrs(ptr(0, 2), ptr(1, 2), mode)                        # re-issued before being consumed
        fs.push_ptrs(ptr(0, 2), ptr(1, 2), mode)
        assert np.array_equal(fs.download(0, 0, 1), want(0, 2, 0, 1))
        assert np.array_equal(fs.download(0, 1, 1), want(0, 3, 0, 1))
        assert np.array_equal(fs.download(0, 2, 0), want(1, 2, 0, 0))
    finally:
        fs.close()
    del t_


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("half", [True, False])
def test_frameset_lazy_right(ctx, mode, half):
    """Left-only push / prefetch + svs_frameset_fetch_right_ptrs for a subset of the streams: the fetched right pyramids
    equal the oracle's, the other streams' right buffers are untouched, the left buffers behave as in a full push."""
    import torch
    B, H, W, T = 5, 370, 1226, 3
    buf = np.stack([np.stack([np.stack([texture(H, W, 500 * e + 7 * t + b) for b in range(B)]) for t in range(T)]) for e in range(2)])
    t_ = torch.from_numpy(buf)
    t_ = t_.cuda() if mode == 1 else (t_.pin_memory() if mode == 2 else t_)
    base = t_.data_ptr()
    ptr = lambda e, t: [base + ((e * T + t) * B + b) * H * W for b in range(B)]
    prep = (lambda im: o.half_nearest(im)) if half else (lambda im: im)
    want = lambda e, t, b, lvl: o.build_pyramid(prep(buf[e, t, b]))[lvl]
    fs = ctx.frameset(B, W, H, half=half)
    try:
        fs.push_ptrs(ptr(0, 0), ptr(1, 0), mode)                     # full push: both eyes of frame 0
        fs.push_ptrs(ptr(0, 1), None, mode)                          # left only
        fs.prefetch_ptrs(ptr(0, 2), None, mode)
        sel = [3, 0]
        fs.fetch_right_ptrs(sel, [ptr(1, 1)[b] for b in sel], mode)  # right images of frame 1 for streams 3 and 0
        for b in range(B):
            for lvl in range(fs.n_levels):
                assert np.array_equal(fs.download(b, 0, lvl), want(0, 1, b, lvl)), (b, lvl)
                assert np.array_equal(fs.download(b, 1, lvl), want(0, 0, b, lvl)), (b, lvl)
        for b in sel:
            for lvl in range(fs.n_levels):
                assert np.array_equal(fs.download(b, 2, lvl), want(1, 1, b, lvl)), (b, lvl)
        fs.push_ptrs(ptr(0, 2), None, mode)                          # prefetch hit (left only)
        fs.fetch_right_ptrs([4], [ptr(1, 2)[4]], mode)
        assert np.array_equal(fs.download(4, 0, 1), want(0, 2, 4, 1))
        assert np.array_equal(fs.download(4, 1, 0), want(0, 1, 4, 0))
        assert np.array_equal(fs.download(4, 2, 2), want(1, 2, 4, 2))
    finally:
        fs.close()
    del t_


@pytest.mark.parametrize("h,w,seed", [(188, 620, 1), (185, 613, 2), (376, 1241, 3), (30, 40, 4)])
def test_lk_bit_identical_to_oracle(ctx, h, w, seed):
    a, b = moved_pair(h, w, seed)
    p0, init = lk_points(a, 5) if h > 40 else (np.array([[10, 10], [20, 15], [35, 25], [5, 28]], np.float32),
                                               np.array([[11, 11], [21, 16], [36, 26], [6, 29]], np.float32))
    q, s = ctx.lk_track(a, b, p0, init)
    wq, ws, _ = geom.lk_track(o.build_pyramid(a), o.build_pyramid(b), p0, init)
    assert np.array_equal(s, ws)
    assert np.array_equal(q.view(np.uint32), wq.view(np.uint32))        # integer-exact restatement: bit-identical
    p1, st, _ = cv2.calcOpticalFlowPyrLK(a, b, p0, init.copy(), winSize=(11, 11), maxLevel=3,
                                         criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                         flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    assert np.array_equal(st.ravel(), s)
    ok = s == 1
    # cv2 sums the same integers in f32 SIMD lanes, the engine sums them exactly.  Measured on these inputs (the engine is
    # bit-identical to the oracle, so this is also oracle vs cv2): 96-99 % of the points bit-identical, median and q95 of
    # the per-point difference 0, q99 <= 2.7e-5 px (the survey's 2e-5 figure), maximum 1.8e-4 px (one flipped
    # convergence decision amplified towards LK's own 0.01-px stopping threshold).  Asserted with 2x head room.
    d = np.abs(p1[ok] - q[ok]).max(1)
    if h > 40:
        assert (d == 0).mean() >= 0.93, (d == 0).mean()
        assert np.quantile(d, 0.95) == 0.0
        assert np.quantile(d, 0.99) <= 6e-5, np.quantile(d, 0.99)
    assert d.max() < 4e-4, d.max()


def test_lk_nonfinite_points_are_lost(ctx):
    """cvFloor(NaN) is INT_MIN in OpenCV: a non-finite previous point or initial guess (world2pixel with z == 0, a
    degenerate triangulation) fails every bounds test and comes back with status 0 and its input value."""
    a, b = moved_pair(188, 620, 1)
    p0 = np.array([[100, 100], [np.nan, 50], [200, np.nan], [300, 80], [np.inf, 60], [50, 60], [-np.inf, 20], [1e20, 30],
                   [150, 90]], np.float32)
    init = p0.copy()
    init[3] = [np.nan, 80]; init[5] = [50, np.inf]
    q, s = ctx.lk_track(a, b, p0, init)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(a, b, p0, init.copy(), winSize=(11, 11), maxLevel=3,
                                         criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                         flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    wq, ws, _ = geom.lk_track(o.build_pyramid(a), o.build_pyramid(b), p0, init)
    assert list(s) == [1, 0, 0, 0, 0, 0, 0, 0, 1]
    assert np.array_equal(s, st.ravel()) and np.array_equal(s, ws)
    assert np.array_equal(q[s == 1].view(np.uint32), wq[s == 1].view(np.uint32))
    lost = s == 0                       # lost points keep their input value, like cv2 (NaN / inf / 1e20 unchanged)
    assert np.array_equal(np.isnan(q[lost]), np.isnan(p1[lost]))
    assert np.array_equal(q[lost][~np.isnan(q[lost])], p1[lost][~np.isnan(p1[lost])])


def test_lk_batch_pairs(ctx):
    B, H, W = 3, 188, 620
    frames = [moved_pair(H, W, 40 + s) for s in range(B)]
    rights = [moved_pair(H, W, 40 + s)[0][:, ::1] for s in range(B)]
    fs = ctx.frameset(B, W, H, half=False)
    try:
        fs.push(np.stack([f[0] for f in frames]), np.stack(rights))
        fs.push(np.stack([f[1] for f in frames]), np.stack([np.roll(f[1], -4, 1) for f in frames]))
        pts = [lk_points(f[0], 5, n_corner=50 + 40 * i, n_rand=10) for i, f in enumerate(frames)]
        pts[1] = (pts[1][0][:0], pts[1][1][:0])      # a stream with no points
        res = ctx.lk_track_batch(fs, 0, [p[0] for p in pts], [p[1] for p in pts])
        for (q, s), f, p in zip(res, frames, pts):
            wq, ws, _ = geom.lk_track(o.build_pyramid(f[0]), o.build_pyramid(f[1]), p[0], p[1])
            assert np.array_equal(s, ws) and np.array_equal(q.view(np.uint32), wq.view(np.uint32))
        res = ctx.lk_track_batch(fs, 1, [p[0] for p in pts], [p[0] for p in pts])     # left -> right
        for (q, s), f, p in zip(res, frames, pts):
            wq, ws, _ = geom.lk_track(o.build_pyramid(f[1]), o.build_pyramid(np.roll(f[1], -4, 1)), p[0], p[0])
            assert np.array_equal(s, ws) and np.array_equal(q.view(np.uint32), wq.view(np.uint32))
    finally:
        fs.close()
