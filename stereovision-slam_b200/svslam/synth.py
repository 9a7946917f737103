"""Synthetic KITTI-shaped stereo sequences with ground truth: a textured corridor
(ground, ceiling, two walls, end cap) ray-cast per pixel for a rectified stereo rig moving along a smooth path.
Deterministic (NumPy + cv2.remap fixed point); used by tests and bench.py because no KITTI data is available.

Calibrations are the public KITTI odometry values, not files of the reference.
"""
import numpy as np
import cv2

CALIB = {
    # name: (width, height, f, cx, cy, baseline)
    "kitti00": (1241, 376, 718.856, 607.1928, 185.2157, 386.1448 / 718.856),
    "kitti05": (1226, 370, 707.0912, 601.8873, 183.1104, 379.8145 / 707.0912),
}
PPM = 32.0          # texels per metre at mip level 0
TEX = 2048          # texture size (periodic)
N_MIP = 7


def _texture(seed):
    rng = np.random.RandomState(seed)
    acc = np.zeros((TEX, TEX), np.float32)
    for sigma, amp in ((1.5, 1.0), (4, 1.0), (12, 1.2), (40, 1.5)):
        n = rng.rand(TEX, TEX).astype(np.float32)
        pad = int(4 * sigma) + 1          # periodic blur: wrap-pad, blur, crop
        n = cv2.copyMakeBorder(n, pad, pad, pad, pad, cv2.BORDER_WRAP)
        n = cv2.GaussianBlur(n, (0, 0), sigma)[pad:-pad, pad:-pad]
        n = (n - n.mean()) / (n.std() + 1e-9)
        acc += amp * n
    acc = (acc - acc.min()) / (acc.max() - acc.min()) * 255.0
    img = acc.astype(np.uint8)
    for _ in range(900):
        x, y = rng.randint(0, TEX - 80), rng.randint(0, TEX - 80)
        w, h = rng.randint(8, 80), rng.randint(8, 60)
        cv2.rectangle(img, (x, y), (x + w, y + h), int(rng.randint(0, 256)), -1)
    mips = [img]
    for _ in range(N_MIP - 1):
        mips.append(cv2.resize(mips[-1], None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA))
    return mips


class Corridor:
    """Planes (camera convention: x right, y down, z forward): ground y=1.65, ceiling y=-3.2, walls x=-5.5 / x=+6.5,
    end cap z=z_end."""

    def __init__(self, calib="kitti05", seed=0, speed=0.8, n_frames=64, noise_sigma=1.0):
        self.W, self.H, self.f, self.cx, self.cy, self.baseline = CALIB[calib]
        self.seed, self.speed, self.n_frames, self.noise_sigma = seed, speed, n_frames, noise_sigma
        # one atlas holding the 5 plane textures x N_MIP levels, each padded by a 2-texel wrap border, so that a
        # frame is rendered with a single cv2.remap
        PAD = 2
        texs = [_texture(seed * 10 + i) for i in range(5)]
        cell_w = TEX + 2 * PAD
        self.atlas = np.zeros((2 * cell_w, 5 * cell_w), np.uint8)
        self.off_x = np.zeros((5, N_MIP), np.float64)
        self.off_y = np.zeros((5, N_MIP), np.float64)
        for k in range(5):
            y0 = 0
            x0 = k * cell_w
            for l in range(N_MIP):
                t = cv2.copyMakeBorder(texs[k][l], PAD, PAD, PAD, PAD, cv2.BORDER_WRAP)
                if l == 1:
                    y0 = cell_w
                    xl = x0
                if l >= 1:
                    self.atlas[y0:y0 + t.shape[0], xl:xl + t.shape[1]] = t
                    self.off_x[k, l], self.off_y[k, l] = xl + PAD, y0 + PAD
                    xl += t.shape[1]
                else:
                    self.atlas[0:t.shape[0], x0:x0 + t.shape[1]] = t
                    self.off_x[k, l], self.off_y[k, l] = x0 + PAD, PAD
        self.z_end = speed * n_frames + 90.0
        u, v = np.meshgrid(np.arange(self.W, dtype=np.float64), np.arange(self.H, dtype=np.float64))
        self.dc = np.stack([(u - self.cx) / self.f, (v - self.cy) / self.f, np.ones_like(u)], -1)
        self.dc_norm = np.linalg.norm(self.dc, axis=-1)

    # ---- trajectory: gentle S-curve, yaw follows the tangent
    def center(self, i):
        z = self.speed * i
        return np.array([1.2 * np.sin(0.035 * z), 0.03 * np.sin(0.11 * z), z])

    def R_wc(self, i):
        z = self.speed * i
        yaw = np.arctan(1.2 * 0.035 * np.cos(0.035 * z))
        pitch = 0.004 * np.sin(0.09 * z)
        cy_, sy_ = np.cos(yaw), np.sin(yaw)
        cp, sp = np.cos(pitch), np.sin(pitch)
        Ry = np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        return Ry @ Rx

    def T_cw(self, i):
        """Ground-truth pose of frame i as [qx qy qz qw tx ty tz] (world -> left camera), relative to frame 0."""
        R0, c0 = self.R_wc(0), self.center(0)
        R, c = self.R_wc(i), self.center(i)
        Rcw = R.T @ R0                       # world' = frame-0 camera coordinates
        t = R.T @ (c0 - c)
        return np.concatenate([_R_to_quat(Rcw), t])

    def render(self, i, eye):
        R = self.R_wc(i)
        o = self.center(i) + (R @ np.array([self.baseline, 0, 0]) if eye else 0)
        d = self.dc @ R.T
        best_t = np.full((self.H, self.W), np.inf)
        plane = np.zeros((self.H, self.W), np.int8)
        for k, (axis, val) in enumerate(((1, 1.65), (1, -3.2), (0, -5.5), (0, 6.5), (2, self.z_end))):
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (val - o[axis]) / d[..., axis]
            t = np.where(t > 1e-6, t, np.inf)
            upd = t < best_t
            best_t = np.where(upd, t, best_t)
            plane = np.where(upd, k, plane)
        hit = o + d * best_t[..., None]
        foot = best_t * self.dc_norm / self.f                              # metres per pixel
        lvl = np.clip(np.rint(np.log2(np.maximum(foot * PPM, 1e-6))), 0, N_MIP - 1).astype(np.int64)
        ta = np.where(plane < 2, hit[..., 0], np.where(plane < 4, hit[..., 1], hit[..., 0])) * PPM
        tb = np.where(plane < 4, hit[..., 2], hit[..., 1]) * PPM
        sc = 1.0 / (1 << lvl).astype(np.float64)
        size = TEX * sc
        pk = plane.astype(np.int64)
        mx = (np.mod(ta * sc, size) + self.off_x[pk, lvl]).astype(np.float32)
        my = (np.mod(tb * sc, size) + self.off_y[pk, lvl]).astype(np.float32)
        smp = cv2.remap(self.atlas, mx, my, cv2.INTER_LINEAR)
        gain = np.array([1.0, 0.8, 0.9, 1.0, 0.7], np.float32)
        out = smp.astype(np.float32) * gain[pk]
        if self.noise_sigma > 0:
            rng = np.random.RandomState((self.seed * 100003 + i * 2 + eye) & 0x7FFFFFFF)
            out = out + rng.randn(self.H, self.W).astype(np.float32) * self.noise_sigma
        return np.clip(np.rint(out), 0, 255).astype(np.uint8)

    def sequence(self, n=None):
        """Returns (left [n,H,W] u8, right [n,H,W] u8, T_cw [n,7])."""
        n = self.n_frames if n is None else n
        L = np.stack([self.render(i, 0) for i in range(n)])
        R = np.stack([self.render(i, 1) for i in range(n)])
        T = np.stack([self.T_cw(i) for i in range(n)])
        return L, R, T

    def K_half(self):
        return np.array([self.f * 0.5, self.f * 0.5, self.cx * 0.5, self.cy * 0.5])

    def K_full(self):
        return np.array([self.f, self.f, self.cx, self.cy])


def _R_to_quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    q = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    return q / np.linalg.norm(q)


def ate_rmse(T_est, T_gt):
    """Absolute trajectory error: RMSE of camera-centre error after rigid (Kabsch, no scale) alignment."""
    def centers(T):
        out = []
        for t in T:
            x, y, z, w = t[:4]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            out.append(-R.T @ t[4:])
        return np.array(out)
    a, b = centers(np.asarray(T_est)), centers(np.asarray(T_gt))
    ma, mb = a.mean(0), b.mean(0)
    U, _, Vt = np.linalg.svd((a - ma).T @ (b - mb))
    D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    Rr = (U @ D @ Vt).T
    err = (Rr @ (a - ma).T).T + mb - b
    return float(np.sqrt((err ** 2).sum(1).mean()))
