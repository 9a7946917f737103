"""CPU oracle — NumPy restatements of the OpenCV stages on the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs do, and only as the checker.

Each function restates what a third-party routine computes at one of the
reference's call sites (the routine's own source is NOT under /root/reference:
OpenCV is an un-vendored dependency, README.md:35 names 4.5.4).  Every
restatement is pinned against the importable ``cv2`` (4.13.0 in this image) by
tests/test_oracle_cv.py, with the reference's exact arguments:

* half_nearest      <- cv::resize(.., 0.5, 0.5, INTER_NEAREST)   src/dataset.cpp:128-129
* min_eig_map       <- cv::cornerMinEigenVal inside GFTTDetector  src/frontend.cpp:24,51
* gftt_select       <- cv::goodFeaturesToTrack selection          src/frontend.cpp:24,51
* feature_mask      <- cv::rectangle mask of tracked features     src/frontend.cpp:42-47
* pyr_down          <- cv::pyrDown inside calcOpticalFlowPyrLK    src/frontend.cpp:105,353
* stereo_bm         <- cv::StereoBM(128,15)::compute              src/dense_reconstruction.cpp:89,114
* bgr2gray          <- cv::cvtColor(BGR2GRAY)                    src/dense_reconstruction.cpp:111-113
* backproject       <- disparity -> depth -> pixel2world loop    src/dense_reconstruction.cpp:116-173

(LK itself, triangulation, pose-only LM and BA are restated in oracle/geom.c.)
"""
import numpy as np

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------
# a0: half-resolution nearest resize
# --------------------------------------------------------------------------
def half_nearest(img):
    """dst[y][x] = src[2y][2x], size (cvRound(W/2), cvRound(H/2)) (round-half-even).

    Restates cv::resize(src, dst, Size(), 0.5, 0.5, INTER_NEAREST) as used at
    src/dataset.cpp:128-129 / :164-165.
    """
    h, w = img.shape[:2]
    dw = int(np.rint(w * 0.5))
    dh = int(np.rint(h * 0.5))
    ys = np.minimum(np.arange(dh) * 2, h - 1)
    xs = np.minimum(np.arange(dw) * 2, w - 1)
    return np.ascontiguousarray(img[ys][:, xs])


# --------------------------------------------------------------------------
# a1: GFTT
# --------------------------------------------------------------------------
def _reflect101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def _fma32(a, b, c):
    """Correctly rounded f32 fma for the operand ranges used here (a*b exact in f64)."""
    return (a.astype(F64) * F64(b) + c.astype(F64)).astype(F32)


def min_eig_map(img, granule=32):
    """cv::cornerMinEigenVal(img, blockSize=3, ksize=3) bit-exactly.

    ``granule`` is the oracle host's SIMD granule for the Sobel-dy row filter
    (columns < granule*floor(W/granule) use the fused form, the tail the
    unfused one); 0 = unfused everywhere (cv2.setUseOptimized(False)).
    SURVEY.md Appendix A.1.
    """
    img = np.asarray(img, np.uint8)
    H, W = img.shape
    s = 1.0 / (4.0 * 3.0 * 255.0)
    ks = F32(s)
    k2 = F32(2.0 * s)
    yi = _reflect101(np.arange(-1, H + 1), H)
    xi = _reflect101(np.arange(-1, W + 1), W)
    p = img[yi][:, xi].astype(F32)  # (H+2, W+2), p[y+1][x+1] = image(y, x)

    # ---- dx: row [-1 0 1], column [1 2 1]*s
    rx = p[:, 2:] - p[:, :-2]  # (H+2, W) exact integers
    top, mid, bot = rx[:-2], rx[1:-1], rx[2:]
    dx = _fma32(top + bot, ks, (mid * k2).astype(F32))

    # ---- dy: row [1 2 1]*s, column [-1 0 1]
    l, c, r = p[:, :-2], p[:, 1:-1], p[:, 2:]
    t0 = (l * ks).astype(F32)
    fused = _fma32(r, ks, _fma32(c, k2, t0))
    t1 = (t0 + (c * k2).astype(F32)).astype(F32)
    unfused = (t1 + (r * ks).astype(F32)).astype(F32)
    nbody = (W // granule) * granule if granule > 0 else 0
    ry = unfused.copy()
    ry[:, :nbody] = fused[:, :nbody]
    dy = (ry[2:] - ry[:-2]).astype(F32)

    cov = np.stack([(dx * dx).astype(F32), (dx * dy).astype(F32), (dy * dy).astype(F32)], axis=-1)

    # ---- 3x3 un-normalised box filter, f64 accumulators, running column sum
    cpad = cov[:, xi].astype(F64)  # (H, W+2, 3)
    rows = (cpad[:, :-2] + cpad[:, 1:-1]) + cpad[:, 2:]  # (H, W, 3)
    rows = rows[yi]  # (H+2, W, 3) rows -1..H
    out = np.empty((H, W, 3), F32)
    SUM = (np.zeros((W, 3), F64) + rows[0]) + rows[1]
    for y in range(H):
        s0 = SUM + rows[y + 2]
        out[y] = s0.astype(F32)
        SUM = s0 - rows[y]

    a = (out[..., 0] * F32(0.5)).astype(F32)
    b = out[..., 1]
    cc = (out[..., 2] * F32(0.5)).astype(F32)
    t = (a - cc).astype(F32)
    rad = ((t * t).astype(F32) + (b * b).astype(F32)).astype(F32)
    lam = ((a + cc).astype(F32) - np.sqrt(rad).astype(F32)).astype(F32)
    return lam


def feature_mask(shape, pts_xy):
    """The mask Frontend::DetectFeatures builds (src/frontend.cpp:42-47):
    255 everywhere, 0 in the inclusive box [round(x-10), round(x+10)] x
    [round(y-10), round(y+10)] (Point2f -> Point rounds half-to-even; the
    subtraction is in f32), clipped to the image."""
    H, W = shape
    mask = np.full((H, W), 255, np.uint8)
    for (x, y) in np.asarray(pts_xy, F32).reshape(-1, 2):
        x0 = int(np.rint(F32(x) - F32(10)))
        x1 = int(np.rint(F32(x) + F32(10)))
        y0 = int(np.rint(F32(y) - F32(10)))
        y1 = int(np.rint(F32(y) + F32(10)))
        x0, x1 = min(x0, x1), max(x0, x1)
        y0, y1 = min(y0, y1), max(y0, y1)
        if x1 < 0 or y1 < 0 or x0 >= W or y0 >= H:
            continue
        mask[max(y0, 0):min(y1, H - 1) + 1, max(x0, 0):min(x1, W - 1) + 1] = 0
    return mask


def gftt_candidates(lam, mask, quality):
    """Thresholded 3x3 local maxima, sorted (value desc, linear index desc)."""
    H, W = lam.shape
    if mask is None:
        mask = np.full((H, W), 255, np.uint8)
    m = mask != 0
    if not m.any():
        return np.zeros(0, np.int64), F32(0)
    max_val = lam[m].max()
    thr = F32(F64(max_val) * F64(quality))
    e = np.where(lam > thr, lam, F32(0))
    d = e.copy()
    pad = np.full((H + 2, W + 2), -np.inf, F32)
    pad[1:-1, 1:-1] = e
    for dy in range(3):
        for dx in range(3):
            d = np.maximum(d, pad[dy:dy + H, dx:dx + W])
    cand = (e != 0) & (e == d) & m
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    idx = np.flatnonzero(cand)
    vals = lam.ravel()[idx]
    order = np.lexsort((-idx, -vals.astype(F64)))
    return idx[order], thr


def gftt_select(lam, mask, max_corners, quality, min_distance):
    """cv::goodFeaturesToTrack's selection.

    Returns (xy float32 [n,2], response float32 [n]).
    """
    H, W = lam.shape
    idx, _ = gftt_candidates(lam, mask, quality)
    vals = lam.ravel()[idx]
    out_xy, out_v = [], []
    if min_distance >= 1:
        cell = int(np.rint(min_distance))
        gw = (W + cell - 1) // cell
        gh = (H + cell - 1) // cell
        grid = {}
        md2 = float(min_distance) * float(min_distance)
        for i, v in zip(idx.tolist(), vals.tolist()):
            y, x = divmod(i, W)
            xc, yc = x // cell, y // cell
            x1, y1 = max(xc - 1, 0), max(yc - 1, 0)
            x2, y2 = min(xc + 1, gw - 1), min(yc + 1, gh - 1)
            good = True
            for yy in range(y1, y2 + 1):
                for xx in range(x1, x2 + 1):
                    for (qx, qy) in grid.get((yy, xx), ()):
                        ddx, ddy = x - qx, y - qy
                        if ddx * ddx + ddy * ddy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid.setdefault((yc, xc), []).append((x, y))
                out_xy.append((x, y))
                out_v.append(v)
                if max_corners > 0 and len(out_xy) == max_corners:
                    break
    else:
        for i, v in zip(idx.tolist(), vals.tolist()):
            y, x = divmod(i, W)
            out_xy.append((x, y))
            out_v.append(v)
            if max_corners > 0 and len(out_xy) == max_corners:
                break
    return (np.asarray(out_xy, F32).reshape(-1, 2), np.asarray(out_v, F32))


def gftt_detect(img, mask, max_corners, quality=0.01, min_distance=20.0, granule=32):
    """cv::GFTTDetector::create(n, 0.01, 20)->detect(img, kps, mask), src/frontend.cpp:24,51."""
    lam = min_eig_map(img, granule)
    return gftt_select(lam, mask, max_corners, quality, min_distance)


def calibrate_granule(cv2):
    """Find the SIMD granule of THIS host's OpenCV build."""
    rng = np.random.RandomState(1234)
    img = rng.randint(0, 256, (40, 203), np.uint8)
    ref = cv2.cornerMinEigenVal(img, 3, ksize=3)
    for g in (32, 16, 8, 4, 64, 0):
        if np.array_equal(min_eig_map(img, g).view(np.uint32), ref.view(np.uint32)):
            return g
    raise RuntimeError("no SIMD granule reproduces cv2.cornerMinEigenVal on this host")


# --------------------------------------------------------------------------
# a2/a3: pyramid
# --------------------------------------------------------------------------
def pyr_down(img):
    """cv::pyrDown: separable [1 4 6 4 1], reflect-101,
    even rows/cols, (sum + 128) >> 8, size ((W+1)/2, (H+1)/2)."""
    img = np.asarray(img, np.uint8)
    H, W = img.shape
    dh, dw = (H + 1) // 2, (W + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int32)
    xs = np.arange(dw) * 2
    ys = np.arange(dh) * 2
    src = img.astype(np.int32)
    rowf = np.zeros((H, dw), np.int32)
    for t in range(5):
        rowf += k[t] * src[:, _reflect101(xs + t - 2, W)]
    out = np.zeros((dh, dw), np.int32)
    for t in range(5):
        out += k[t] * rowf[_reflect101(ys + t - 2, H)]
    return ((out + 128) >> 8).astype(np.uint8)


def lk_num_levels(w, h, win, max_level):
    """How many pyramid levels buildOpticalFlowPyramid keeps: stop before a level
    whose width or height is <= the window."""
    n = 0
    for lvl in range(1, max_level + 1):
        w, h = (w + 1) // 2, (h + 1) // 2
        if w <= win or h <= win:
            break
        n = lvl
    return n


def build_pyramid(img, win=11, max_level=3):
    levels = [np.ascontiguousarray(img)]
    for _ in range(lk_num_levels(img.shape[1], img.shape[0], win, max_level)):
        levels.append(pyr_down(levels[-1]))
    return levels


# --------------------------------------------------------------------------
# a10: StereoBM
# --------------------------------------------------------------------------
def bm_prefilter_xsobel(img, cap=31):
    """StereoBM's PREFILTER_XSOBEL."""
    img = np.asarray(img, np.uint8).astype(np.int32)
    H, W = img.shape
    out = np.full((H, W), cap, np.int32)
    d = np.zeros((H, W), np.int32)
    d[:, 1:-1] = img[:, 2:] - img[:, :-2]
    ym = np.arange(H) - 1
    ym[0] = 1 if H > 1 else 0
    yp = np.arange(H) + 1
    yp[-1] = H - 2 if H > 1 else 0
    v = d[ym] + 2 * d + d[yp]
    out[:, 1:-1] = np.clip(v[:, 1:-1], -cap, cap) + cap
    if H % 2 == 1:
        out[H - 1, :] = cap
    return out.astype(np.uint8)


def stereo_bm(left, right, ndisp=128, block=15, cap=31, texture=10, uniq=15):
    """cv::StereoBM::create(ndisp, block)->compute(left, right) -> int16 (16*d, invalid = -16)."""
    H, W = left.shape
    L = bm_prefilter_xsobel(left, cap).astype(np.int32)
    R = bm_prefilter_xsobel(right, cap).astype(np.int32)
    r = block // 2
    disp = np.full((H, W), -16, np.int16)
    x_lo, x_hi = ndisp - 1 + r, W - r
    y_lo, y_hi = r, H - r
    if x_hi <= x_lo or y_hi <= y_lo:
        return disp

    def box(a):  # window sum over (2r+1)^2, valid region only, via integral image
        ii = np.zeros((a.shape[0] + 1, a.shape[1] + 1), np.int64)
        ii[1:, 1:] = a.cumsum(0).cumsum(1)
        n = 2 * r + 1
        return ii[n:, n:] - ii[:-n, n:] - ii[n:, :-n] + ii[:-n, :-n]

    nx, ny = x_hi - x_lo, y_hi - y_lo
    sad = np.empty((ndisp, ny, nx), np.int64)  # indexed by search index d = ndisp-1-D
    for D in range(ndisp):
        ad = np.zeros((H, W), np.int32)
        ad[:, D:] = np.abs(L[:, D:] - R[:, :W - D])
        b = box(ad)  # b[y0, x0] = window centred (y0+r, x0+r)
        sad[ndisp - 1 - D] = b[y_lo - r:y_hi - r, x_lo - r:x_hi - r]
    tex = box(np.abs(L - cap))[y_lo - r:y_hi - r, x_lo - r:x_hi - r]

    mind = sad.argmin(0)  # first minimum in ascending d
    minsad = np.take_along_axis(sad, mind[None], 0)[0]
    valid = tex >= texture
    thresh = minsad + (minsad * uniq) // 100
    dd = np.arange(ndisp)[:, None, None]
    viol = (sad <= thresh[None]) & (np.abs(dd - mind[None]) > 1)
    valid &= ~viol.any(0)
    sp = np.concatenate([sad[1:2], sad, sad[ndisp - 2:ndisp - 1]], 0)  # sad[-1]=sad[1], sad[n]=sad[n-2]
    p = np.take_along_axis(sp, (mind + 2)[None], 0)[0]
    n = np.take_along_axis(sp, mind[None], 0)[0]
    den = p + n - 2 * minsad + np.abs(p - n)
    num = (p - n) * 256
    q = np.where(den != 0, np.sign(num) * (np.abs(num) // np.where(den != 0, den, 1)), 0)
    val = (((ndisp - 1 - mind) * 256 + q + 15) >> 4).astype(np.int16)
    disp[y_lo:y_hi, x_lo:x_hi] = np.where(valid, val, np.int16(-16))
    return disp


def bgr2gray(bgr):
    """cv::cvtColor(BGR2GRAY) for 8-bit images (src/dense_reconstruction.cpp:111-113): fixed point, shift 15."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def backproject(disp16, bgr, K4, baseline, cam_pose_inv, T_cw, se3_act, se3_inv):
    """src/dense_reconstruction.cpp:116-173: disparity/16 (f32) -> depth = fx*b/d (f32) -> for x outer, y inner,
    every pixel with depth >= 1: Camera::pixel2world (src/camera.cpp:58-86) in f64, stored as f32 + RGB."""
    h, w = disp16.shape
    d = disp16.astype(F32) * F32(1.0 / 16.0)
    fb = F32(F32(K4[0]) * F32(baseline))
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(d > 0, (fb / np.where(d > 0, d, F32(1))).astype(F32), F32(0))
    Twc = se3_inv(T_cw)
    pts, cols = [], []
    for x in range(w):
        for y in range(h):
            if z[y, x] < 1:
                continue
            depth = float(z[y, x])
            pc = np.array([(x - K4[2]) * depth / K4[0], (y - K4[3]) * depth / K4[1], depth])
            pw = se3_act(Twc, se3_act(cam_pose_inv, pc))
            pts.append(pw.astype(F32))
            cols.append(bgr[y, x, ::-1])
    return (np.asarray(pts, F32).reshape(-1, 3), np.asarray(cols, np.uint8).reshape(-1, 3))
