"""CPU checks of the pose-graph oracle (oracle/geom.c::orc_pose_graph_optimize, restating LoopClosure::PoseGraphOptimization,
reference src/loopclosure.cpp:641-799, with g2o's LM from SURVEY.md Appendix B — parity unpinned like the other g2o blocks):
closed-form Jacobians against finite differences, the minimum against scipy.optimize.least_squares on residuals written
independently in NumPy, the fixed vertex, and the landmark move."""
import numpy as np

from oracle import geom
from util import pose_graph_problem


def _err(M, A, B):
    return geom.se3_log(geom.se3_mul(geom.se3_mul(geom.se3_inv(M), A), geom.se3_inv(B)))


def test_closed_form_jacobians_match_finite_differences():
    rng = np.random.RandomState(1)
    rnd = lambda st, sr: geom.se3_exp(np.concatenate([rng.randn(3) * st, rng.randn(3) * sr]))
    worst = 0.0
    for t in range(60):
        A, B = rnd(3, 0.8), rnd(3, 0.8)
        M = geom.se3_mul(geom.se3_mul(A, geom.se3_inv(B)), rnd(0.5, 0.3 if t % 3 else 1e-8))
        Ja, Jb = geom.pg_edge_jac(M, A, B, 0)
        d = 1e-6
        for k in range(6):
            u = np.zeros(6); u[k] = d
            na = (_err(M, geom.se3_mul(geom.se3_exp(u), A), B) - _err(M, geom.se3_mul(geom.se3_exp(-u), A), B)) / (2 * d)
            nb = (_err(M, A, geom.se3_mul(geom.se3_exp(u), B)) - _err(M, A, geom.se3_mul(geom.se3_exp(-u), B))) / (2 * d)
            worst = max(worst, np.abs(Ja[:, k] - na).max(), np.abs(Jb[:, k] - nb).max())
        na, nb = geom.pg_edge_jac(M, A, B, 1)             # g2o's delta = 1e-9 differences: same up to their rounding noise
        assert np.abs(Ja - na).max() < 1e-4 and np.abs(Jb - nb).max() < 1e-4
    assert worst < 1e-7, worst


def test_minimum_matches_scipy():
    from scipy.optimize import least_squares
    pr = pose_graph_problem(4, n=12, loops=((-1, 1),))
    P, st = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 60, 0)
    assert st.chi2 < 0.3 * st.chi2_init and np.array_equal(P[0], pr["poses"][0])

    def resid(x):
        T = [pr["poses"][0]] + [geom.se3_mul(geom.se3_exp(x[6 * i:6 * i + 6]), pr["poses"][i + 1]) for i in range(11)]
        return np.concatenate([_err(pr["meas"][e], T[pr["edge_a"][e]], T[pr["edge_b"][e]]) for e in range(len(pr["edge_a"]))])
    sol = least_squares(resid, np.zeros(66), xtol=1e-15, ftol=1e-15, gtol=1e-15)
    assert abs(2 * sol.cost - st.chi2) < 1e-6 * st.chi2
    Ts = [pr["poses"][0]] + [geom.se3_mul(geom.se3_exp(sol.x[6 * i:6 * i + 6]), pr["poses"][i + 1]) for i in range(11)]
    assert np.abs(np.array(Ts) - P).max() < 1e-5


def test_numeric_mode_and_landmark_move():
    pr = pose_graph_problem(6, n=25)
    Pa, sa = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 0)
    Pn, sn = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 1)
    assert np.abs(Pa - Pn).max() < 1e-5 and abs(sa.chi2 - sn.chi2) < 1e-4 * sa.chi2
    lms = np.random.RandomState(0).randn(30, 3) * 5
    kf = np.arange(30) % 25
    moved = geom.move_landmarks(lms, kf, pr["poses"], Pa)
    for i in (0, 7, 29):          # the point keeps its coordinates in its keyframe's camera frame
        assert np.allclose(geom.se3_act(Pa[kf[i]], moved[i]), geom.se3_act(pr["poses"][kf[i]], lms[i]), atol=1e-10)
