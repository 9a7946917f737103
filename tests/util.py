"""Shared synthetic inputs for the tests (SURVEY.md §8d value distributions)."""
import cv2
import numpy as np


def texture(h, w, seed, rects=None):
    """Gaussian-blurred (sigma 2) uniform noise normalised to 0..255 plus random filled rectangles."""
    rng = np.random.RandomState(seed)
    a = rng.rand(h, w).astype(np.float32)
    a = cv2.GaussianBlur(a, (0, 0), 2.0)
    a = (a - a.min()) / (a.max() - a.min()) * 255
    img = a.astype(np.uint8)
    if rects is None:
        rects = max(4, int(40 * h * w / (620 * 188)))
    for _ in range(rects):
        x, y = rng.randint(0, w), rng.randint(0, h)
        ww, hh = rng.randint(5, 60), rng.randint(5, 40)
        cv2.rectangle(img, (x, y), (x + ww, y + hh), int(rng.randint(0, 256)), -1)
    return img


def moved_pair(h, w, seed):
    """An image and a slightly affinely-warped copy of it (for LK)."""
    big = texture(h + 40, w + 40, seed)
    a = big[20:20 + h, 20:20 + w].copy()
    m = np.array([[1.01, 0.005, 3.3], [-0.004, 0.995, -2.1]], np.float32)
    b = cv2.warpAffine(big, m, (w + 40, h + 40))[20:20 + h, 20:20 + w].copy()
    return a, b


def lk_points(a, seed, n_corner=300, n_rand=60):
    h, w = a.shape
    kps = cv2.GFTTDetector_create(n_corner, 0.01, 10).detect(a, None)
    p0 = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    rng = np.random.RandomState(seed)
    extra = np.stack([rng.rand(n_rand) * w, rng.rand(n_rand) * h], 1).astype(np.float32)
    edge = np.array([[0.5, 0.5], [w - 1.2, h - 1.1], [2, h - 2], [w - 3, 3], [w - 0.5, 10], [5, h - 0.6]], np.float32)
    p0 = np.concatenate([p0, extra, edge])
    init = p0 + rng.randn(*p0.shape).astype(np.float32) * 2.0 + np.array([3, -2], np.float32)
    init[-3:] += np.array([30, 30], np.float32)
    return p0, init.astype(np.float32)


def stereo_pair(h, w, seed, dmin=2.0, dmax=120.0):
    """Rectified pair with a horizontal disparity ramp dmin..dmax: R(x) = L(x + d(x))."""
    big = texture(h, w + 130, seed)
    disp = (np.linspace(dmin, dmax, w)[None, :] * np.ones((h, 1))).astype(np.float32)
    xs = np.arange(w)[None, :] + disp
    ys = (np.arange(h)[:, None] * np.ones((1, w))).astype(np.float32)
    r = cv2.remap(big, xs.astype(np.float32), ys, cv2.INTER_LINEAR)
    return big[:, :w].copy(), r
