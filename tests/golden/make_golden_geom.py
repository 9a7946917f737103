"""Generates tests/golden/golden_geom.npz: inputs and OUTPUTS of the C oracle (oracle/geom.c) for one triangulation batch,
one pose-only LM problem and one window BA problem.  g2o / Sophus / Eigen cannot be imported (parity unpinned, DESIGN.md
§3), so these vectors pin what CAN be pinned: that the oracle build on any box — and the CUDA path — reproduce the numbers
whose minima were cross-checked against scipy (tests/test_oracle_geom.py).  Run:  python tests/golden/make_golden_geom.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import geom  # noqa: E402
from util import K05, BASELINE, EXT_L, EXT_R, pose_problem, ba_problem  # noqa: E402

out = {}
rng = np.random.RandomState(42)
n = 64
u, v = rng.uniform(150, 600, n), rng.uniform(5, 180, n)
lxy, rxy = np.stack([u, v], 1), np.stack([u - rng.uniform(1.5, 90, n), v + rng.randn(n) * 1.2], 1)
xyz, ok = geom.triangulate(lxy, rxy, K05, K05, BASELINE)
out.update(tri_l=lxy, tri_r=rxy, tri_xyz=xyz, tri_ok=ok)
pts, uv, K, T0, _ = pose_problem(31, m=120)
T, outl, ninl, st = geom.pose_only_lm(pts, uv, K, T0)
out.update(po_pts=pts, po_uv=uv, po_K=K, po_T0=T0, po_T=T, po_outlier=outl, po_ninl=np.array(ninl), po_iterations=np.array(st.iterations))
prob, _, _ = ba_problem(32, n_kf=8, n_lm=150)
P, L, chi2, sb = geom.ba_optimize(prob["poses"], prob["lms"], prob["edge_kf"], prob["edge_lm"], prob["edge_cam"], prob["edge_uv"],
                                  K05, K05, EXT_L, EXT_R)
out.update({"ba_" + k: np.asarray(v) for k, v in prob.items()})
out.update(ba_P=P, ba_L=L, ba_chi2=chi2, ba_stats=np.array([sb.iterations, sb.trials, sb.chi2, sb.chi2_init]))
np.savez_compressed(os.path.join(HERE, "golden_geom.npz"), **out)
print("wrote golden_geom.npz", {k: getattr(v, "shape", None) for k, v in out.items()})
