"""A NumPy stand-in for svslam.ba_shard.Shard (TEST INFRASTRUCTURE): the same four device steps of the landmark-sharded
bundle adjustment — linearise, Schur partial sums, solve + trial state, accept — computed densely on the CPU, with the
buffer layouts of include/svslam.h (svs_ba_shard_*).  It lets the Levenberg-Marquardt driver svslam.ba_shard.lm_optimize
and its reductions (torch.distributed all-reduce over gloo) be tested without a GPU against the C oracle."""
import numpy as np
import torch

from oracle import geom


def _R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _skew(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


class _Ctx:
    def sync(self):
        pass


class MockShard:
    def __init__(self, poses, lms, edge_kf, edge_lm, edge_cam, edge_uv, K_left, K_right, ext_left, ext_right, huber_delta=5.991):
        self.ctx = _Ctx()
        self.poses, self.lms = np.array(poses, float).reshape(-1, 7), np.array(lms, float).reshape(-1, 3)
        self.ekf, self.elm, self.ecam = np.asarray(edge_kf), np.asarray(edge_lm), np.asarray(edge_cam)
        self.euv = np.asarray(edge_uv, float).reshape(-1, 2)
        self.K, self.ext, self.hd = [np.asarray(K_left, float), np.asarray(K_right, float)], [np.asarray(ext_left, float), np.asarray(ext_right, float)], huber_delta
        self.N, self.L, self.E = len(self.poses), len(self.lms), len(self.ekf)
        N = self.N
        self.lin = torch.zeros(42 * N + 2, dtype=torch.float64)
        self.maxdiag = torch.zeros(1, dtype=torch.float64)
        self.red = torch.zeros(36 * N * N + 6 * N, dtype=torch.float64)
        self.flag = torch.ones(1, dtype=torch.int32)
        self.tri = torch.zeros(4, dtype=torch.float64)
        self.poseT, self.lmT = self.poses.copy(), self.lms.copy()

    # ---- EdgeProjection (g2o_types.h:200-216) and its analytic Jacobians (left-multiplicative pose update)
    def _edge(self, T, p, cam, uv, jac):
        K, ext = self.K[cam], self.ext[cam]
        a = _R(T[:4]) @ p + T[4:]
        Re = _R(ext[:4])
        c = Re @ a + ext[4:]
        e = uv - np.array([K[0] * c[0] / c[2] + K[2], K[1] * c[1] / c[2] + K[3]])
        if not jac:
            return e, None, None
        Zi = 1.0 / c[2]
        D = np.array([[-K[0] * Zi, 0, K[0] * c[0] * Zi * Zi], [0, -K[1] * Zi, K[1] * c[1] * Zi * Zi]])
        Jp = D @ np.hstack([Re, -Re @ _skew(a)])
        Jl = D @ Re @ _R(T[:4])
        return e, Jp, Jl

    def _rho(self, e2):
        d2 = self.hd * self.hd
        if e2 <= d2:
            return e2, 1.0
        s = np.sqrt(e2)
        return 2 * s * self.hd - d2, self.hd / s

    def _chi2(self, P, Lm):
        return sum(self._rho(float(e @ e))[0] for e in (self._edge(P[self.ekf[k]], Lm[self.elm[k]], self.ecam[k], self.euv[k], False)[0]
                                                         for k in range(self.E)))

    def linearize(self):
        N, L = self.N, self.L
        self.Hpp, self.bp = np.zeros((N, 6, 6)), np.zeros((N, 6))
        self.Hll, self.bl, self.W = np.zeros((L, 3, 3)), np.zeros((L, 3)), np.zeros((self.E, 6, 3))
        chi = 0.0
        for k in range(self.E):
            i, l = self.ekf[k], self.elm[k]
            e, Jp, Jl = self._edge(self.poses[i], self.lms[l], self.ecam[k], self.euv[k], True)
            r0, r1 = self._rho(float(e @ e))
            chi += r0
            self.Hpp[i] += r1 * Jp.T @ Jp; self.bp[i] -= r1 * Jp.T @ e
            self.Hll[l] += r1 * Jl.T @ Jl; self.bl[l] -= r1 * Jl.T @ e
            self.W[k] = r1 * Jp.T @ Jl
        self.lin[:36 * N] = torch.from_numpy(self.Hpp.reshape(-1)); self.lin[36 * N:42 * N] = torch.from_numpy(self.bp.reshape(-1))
        self.lin[42 * N] = chi; self.lin[42 * N + 1] = 0.0
        used = np.bincount(self.elm, minlength=L) > 0
        self.maxdiag[0] = float(np.abs(np.diagonal(self.Hll[used], axis1=1, axis2=2)).max()) if used.any() else 0.0

    def schur(self, lam):
        N, np6 = self.N, 6 * self.N
        S, g = np.zeros((np6, np6)), np.zeros(np6)
        ok = 1
        self.Vinv = np.zeros_like(self.Hll)
        for l in np.unique(self.elm):
            V = self.Hll[l] + lam * np.eye(3)
            try:
                self.Vinv[l] = np.linalg.inv(V)
            except np.linalg.LinAlgError:
                ok = 0
                continue
            ks = np.flatnonzero(self.elm == l)
            for k1 in ks:
                X = self.W[k1] @ self.Vinv[l]
                i = self.ekf[k1]
                g[6 * i:6 * i + 6] -= X @ self.bl[l]
                for k2 in ks:
                    j = self.ekf[k2]
                    S[6 * i:6 * i + 6, 6 * j:6 * j + 6] -= X @ self.W[k2].T
        self.red[:np6 * np6] = torch.from_numpy(S.reshape(-1)); self.red[np6 * np6:] = torch.from_numpy(g)
        self.flag[0] = ok

    def try_step(self, lin, red, lam, flag_ok):
        N, np6 = self.N, 6 * self.N
        lin, red = lin.numpy(), red.numpy()
        S = red[:np6 * np6].reshape(np6, np6).copy()
        for a in range(N):
            S[6 * a:6 * a + 6, 6 * a:6 * a + 6] += lin[36 * a:36 * a + 36].reshape(6, 6) + lam * np.eye(6)
        bp = lin[36 * N:42 * N]
        g = bp + red[np6 * np6:]
        ok = bool(flag_ok)
        xp = np.zeros(np6)
        if ok:
            try:
                np.linalg.cholesky(S)            # Eigen::LDLT::isPositive for a symmetric matrix
                xp = np.linalg.solve(S, g)
            except np.linalg.LinAlgError:
                ok = False
        self.poseT = np.array([geom.se3_mul(geom.se3_exp(xp[6 * a:6 * a + 6]), self.poses[a]) for a in range(N)])
        self.lmT = self.lms.copy()
        sc = 0.0
        for l in np.unique(self.elm):
            x3 = np.zeros(3)
            if ok:
                c3 = self.bl[l].copy()
                for k in np.flatnonzero(self.elm == l):
                    c3 -= self.W[k].T @ xp[6 * self.ekf[k]:6 * self.ekf[k] + 6]
                x3 = self.Vinv[l] @ c3
            self.lmT[l] = self.lms[l] + x3
            sc += float(x3 @ (lam * x3 + self.bl[l]))
        self.tri[0] = self._chi2(self.poseT, self.lmT)
        self.tri[1] = sc
        self.tri[2] = float(xp @ (lam * xp + bp))
        self.tri[3] = 1.0 if ok else 0.0

    def accept(self):
        self.poses, self.lms = self.poseT.copy(), self.lmT.copy()

    def get(self):
        chi2 = np.zeros(self.E)
        for k in range(self.E):
            e = self._edge(self.poseT[self.ekf[k]], self.lmT[self.elm[k]], self.ecam[k], self.euv[k], False)[0]
            chi2[k] = float(e @ e)
        return self.poses.copy(), self.lms.copy(), chi2
