"""Diagnostic (GPU box): how closely do the engine and the CPU oracle agree in g2o's numeric-Jacobian mode?  Prints the
quantiles tests/test_gpu_geom.py asserts.  Lives under tests/ because it imports the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "stereovision-slam_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import svslam  # noqa: E402
from oracle import geom  # noqa: E402
from util import K05, EXT_L, EXT_R, ba_problem, quat_to_R, rel_to_norm  # noqa: E402


def main():
    ctx = svslam.Context(0)
    out = []
    cen = lambda Q: np.array([-quat_to_R(q[:4]).T @ q[4:] for q in Q])
    for (n_kf, n_lm) in ((10, 300), (20, 600), (4, 40), (10, 1000)):
        for seed in range(6):
            pr = ba_problem(seed, n_kf=n_kf, n_lm=n_lm)[0]
            (P, L, chi2, st), = ctx.ba_optimize([pr], K05, K05, EXT_L, EXT_R, jac_mode=1)
            wP, wL, wchi2, wst = geom.ba_optimize(pr["poses"], pr["lms"], pr["edge_kf"], pr["edge_lm"], pr["edge_cam"], pr["edge_uv"],
                                                  K05, K05, EXT_L, EXT_R, jac_mode=1)
            aP, aL, achi2, ast = geom.ba_optimize(pr["poses"], pr["lms"], pr["edge_kf"], pr["edge_lm"], pr["edge_cam"], pr["edge_uv"],
                                                  K05, K05, EXT_L, EXT_R, jac_mode=0)
            rl, rc = rel_to_norm(L, wL), rel_to_norm(cen(P), cen(wP))
            al = rel_to_norm(wL, aL)
            out.append(dict(n_kf=n_kf, n_lm=n_lm, seed=seed, it=(st.iterations, wst.iterations), tr=(st.trials, wst.trials),
                            lm_med=float(np.median(rl)), lm_q95=float(np.quantile(rl, .95)), lm_max=float(rl.max()),
                            cen_med=float(np.median(rc)), cen_q95=float(np.quantile(rc, .95)), cen_max=float(rc.max()),
                            cen_abs_max=float(np.abs(cen(P) - cen(wP)).max()), chi2_rel=float(abs(st.chi2 - wst.chi2) / wst.chi2),
                            oracle_numeric_vs_analytic_lm_q95=float(np.quantile(al, .95)), oracle_nva_lm_max=float(al.max())))
            print(json.dumps(out[-1]), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
