"""Pose-graph optimisation (LoopClosure::PoseGraphOptimization, reference src/loopclosure.cpp:641-799)
on the GPU — one cooperative kernel, grid-wide blocked LDLT of the dense 6K x 6K system — against the C oracle."""
import numpy as np
import pytest

from oracle import geom
from util import pose_graph_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,loops", [(40, ((-1, 2),)), (12, ((-1, 0),)), (150, ((-1, 3), (100, 20))), (7, ())])
def test_pose_graph_matches_oracle(ctx, n, loops):
    pr = pose_graph_problem(n, n=n, loops=loops)
    # (1) before convergence the LM control flow is well defined: identical iteration / trial counts, poses to 1e-9
    wP, wst = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 3, 0)
    P, st = ctx.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 3, 0)
    assert (st.iterations, st.trials) == (wst.iterations, wst.trials)
    assert abs(st.chi2_init - wst.chi2_init) <= 1e-10 * wst.chi2_init + 1e-18
    assert abs(st.chi2 - wst.chi2) <= 1e-8 * wst.chi2 + 1e-15
    assert np.array_equal(P[0], pr["poses"][0])                       # the fixed vertex never moves
    assert np.abs(P - wP).max() < 1e-9 * max(1.0, np.abs(wP).max())
    # (2) optimize(22) as the reference calls it: g2o's LM has no convergence test, so once the minimum is reached the sign of
    # rho is rounding noise and the iteration count is not comparable — the minimum is
    wP, wst = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 0)
    P, st = ctx.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 0)
    assert abs(st.chi2 - wst.chi2) <= 1e-6 * wst.chi2 + 1e-13
    assert np.abs(P - wP).max() < 1e-6 * max(1.0, np.abs(wP).max())
    if loops:
        assert st.chi2 < 0.2 * st.chi2_init                           # the loop edge pulls the drifted chain together
    # the reference's mode: numeric Jacobians (delta 1e-9); the gauge is fixed, so the noise is not amplified much
    nP, nst = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 1)
    P1, st1 = ctx.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 1)
    assert np.abs(P1 - nP).max() < 1e-4 * max(1.0, np.abs(nP).max())
    assert abs(st1.chi2 - nst.chi2) <= 1e-3 * nst.chi2 + 1e-12


def test_pose_graph_large_dense_system(ctx):
    """400 keyframes -> a 2394 x 2394 dense system (75 panels of the cooperative LDLT, 64 x 64 trailing tiles)."""
    pr = pose_graph_problem(3, n=400, loops=((-1, 5), (300, 40), (200, 199)), step=0.05)
    wP, wst = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 6, 0)
    P, st = ctx.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 6, 0)
    assert (st.iterations, st.trials) == (wst.iterations, wst.trials)
    assert np.abs(P - wP).max() < 1e-7 * max(1.0, np.abs(wP).max())
    assert abs(st.chi2 - wst.chi2) <= 1e-7 * wst.chi2


def test_pose_graph_edge_cases(ctx):
    pr = pose_graph_problem(5, n=10)
    fixed_all = np.ones(10, np.uint8)
    P, st = ctx.pose_graph_optimize(pr["poses"], fixed_all, pr["edge_a"], pr["edge_b"], pr["meas"])
    assert np.array_equal(P, pr["poses"]) and st.iterations == 0      # nothing to optimise
    P, st = ctx.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"][:0], pr["edge_b"][:0], pr["meas"][:0])
    assert np.array_equal(P, pr["poses"])
    # two fixed vertices, and an isolated vertex that no edge touches
    fixed = pr["fixed"].copy(); fixed[5] = 1
    keep = pr["edge_a"] < 9
    wP, wst = geom.pose_graph_optimize(pr["poses"], fixed, pr["edge_a"][keep], pr["edge_b"][keep], pr["meas"][keep], 4, 0)
    P, st = ctx.pose_graph_optimize(pr["poses"], fixed, pr["edge_a"][keep], pr["edge_b"][keep], pr["meas"][keep], 4, 0)
    assert np.array_equal(P[[0, 5, 9]], pr["poses"][[0, 5, 9]])
    assert (st.iterations, st.trials) == (wst.iterations, wst.trials) and np.abs(P - wP).max() < 1e-8


def test_move_landmarks(ctx):
    pr = pose_graph_problem(2, n=20)
    rng = np.random.RandomState(0)
    new, _ = geom.pose_graph_optimize(pr["poses"], pr["fixed"], pr["edge_a"], pr["edge_b"], pr["meas"], 22, 0)
    lms = rng.randn(500, 3) * 10
    kf = rng.randint(-1, 20, 500).astype(np.int32)
    got = ctx.pose_graph_move_landmarks(lms, kf, pr["poses"], new)
    want = geom.move_landmarks(lms, kf, pr["poses"], new)
    assert np.abs(got - want).max() < 1e-10 * max(1.0, np.abs(want).max())
    assert np.array_equal(got[kf < 0], lms[kf < 0])
