"""Sharded / large-window BA (svs_ba_shard_*) on one GPU: the persistent cooperative solver with one shard,
and with several shards whose kernels run CONCURRENTLY on slices of the SMs and exchange their partial reduced systems
through each other's windows — the same in-kernel protocol that runs across GPUs over NVLink — against the CPU oracle."""
import numpy as np
import pytest

from oracle import geom
from util import K05, EXT_L, EXT_R, ba_problem, ba_problem_big, rel_to_norm

pytestmark = pytest.mark.gpu


def _run(ctx, prob, world, jac_mode=0, max_iter=10):
    import svslam
    from svslam import ba_shard
    parts = ba_shard.split_problem(prob, world)
    ctxs = [ctx] + [svslam.Context(0) for _ in range(world - 1)]      # one stream per shard: the kernels must overlap
    shards = [ba_shard.Shard(c, p["poses"], p["lms"], p["edge_kf"], p["edge_lm"], p["edge_cam"], p["edge_uv"],
                             K05, K05, EXT_L, EXT_R, 5.991, jac_mode) for c, (p, _) in zip(ctxs, parts)]
    try:
        if world > 1:
            ba_shard.wire_local(shards)
        st, sts = ba_shard.optimize_all(shards, max_iter)
        assert all(x == st for x in sts)                   # every rank reports the identical statistics
        lms = np.zeros((len(prob["lms"]), 3))
        chi2 = np.zeros(len(prob["edge_kf"]))
        poses = None
        elm = np.asarray(prob["edge_lm"])
        from svslam.dist import partition_by_weight
        owner = partition_by_weight(np.bincount(elm, minlength=len(prob["lms"])), world)
        for r, (s, (p, ids)) in enumerate(zip(shards, parts)):
            P, L, c2 = s.get()
            poses = P if poses is None else poses
            assert np.array_equal(P, poses)                # every shard holds the bitwise identical pose state
            lms[ids] = L
            chi2[owner[elm] == r] = c2
        return poses, lms, st, chi2
    finally:
        for s in shards:
            s.close()
        for c in ctxs[1:]:
            c.close()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_shards_match_oracle(ctx, world):
    prob, _, _ = ba_problem(5, n_kf=12, n_lm=400)
    wP, wL, wchi2, wst = geom.ba_optimize(prob["poses"], prob["lms"], prob["edge_kf"], prob["edge_lm"], prob["edge_cam"],
                                          prob["edge_uv"], K05, K05, EXT_L, EXT_R)
    P, L, st, chi2 = _run(ctx, prob, world)
    assert (st["iterations"], st["trials"]) == (wst.iterations, wst.trials)
    assert abs(st["chi2_init"] - wst.chi2_init) < 1e-10 * wst.chi2_init and abs(st["chi2"] - wst.chi2) < 1e-8 * wst.chi2
    assert np.abs(P - wP).max() < 1e-8 and rel_to_norm(L, wL).max() < 1e-8
    assert np.abs(chi2 - wchi2).max() < 1e-6 * max(1.0, wchi2.max())


def test_config4_scale_one_vs_four_shards(ctx):
    prob = ba_problem_big(4, n_kf=50, n_lm=20000)            # config-4 shape at 1/5 of the landmarks (test time)
    P1, L1, st1, _ = _run(ctx, prob, 1, max_iter=4)
    P4, L4, st4, _ = _run(ctx, prob, 4, max_iter=4)
    assert st1["chi2"] < 0.5 * st1["chi2_init"]
    assert (st1["iterations"], st1["trials"]) == (st4["iterations"], st4["trials"])
    assert np.abs(P1 - P4).max() < 1e-9 and rel_to_norm(L1, L4).max() < 1e-9     # 1-GPU result == sharded result


def test_numeric_jacobians_and_rejected_trials(ctx):
    """g2o's numeric Jacobians through the cooperative solver, and a start far enough from the optimum that trials are
    rejected (lambda grows, the stale-error edge chi2 semantics matter)."""
    prob, _, _ = ba_problem(7, n_kf=8, n_lm=200, pose_sigma=(0.3, 0.03), lm_sigma=1.0)
    wP, wL, wchi2, wst = geom.ba_optimize(prob["poses"], prob["lms"], prob["edge_kf"], prob["edge_lm"], prob["edge_cam"],
                                          prob["edge_uv"], K05, K05, EXT_L, EXT_R)
    P, L, st, chi2 = _run(ctx, prob, 2)
    assert (st["iterations"], st["trials"]) == (wst.iterations, wst.trials)
    assert abs(st["chi2"] - wst.chi2) < 1e-7 * wst.chi2 and np.abs(P - wP).max() < 1e-7
    P, L, st, _ = _run(ctx, prob, 1, jac_mode=1)
    nP, nL, _, nst = geom.ba_optimize(prob["poses"], prob["lms"], prob["edge_kf"], prob["edge_lm"], prob["edge_cam"],
                                      prob["edge_uv"], K05, K05, EXT_L, EXT_R, jac_mode=1)
    assert abs(st["chi2"] - nst.chi2) < 1e-3 * nst.chi2


def test_missing_peer_times_out_instead_of_hanging(ctx):
    """A rank whose peer never launches must come back with an error after the in-kernel timeout, not hang the GPU."""
    import time
    import svslam
    from svslam import ba_shard
    prob, _, _ = ba_problem(1, n_kf=4, n_lm=40)
    parts = ba_shard.split_problem(prob, 2)
    c2 = svslam.Context(0)
    shards = [ba_shard.Shard(c, p["poses"], p["lms"], p["edge_kf"], p["edge_lm"], p["edge_cam"], p["edge_uv"], K05, K05, EXT_L, EXT_R)
              for c, (p, _) in zip((ctx, c2), parts)]
    try:
        ba_shard.wire_local(shards)
        t0 = time.time()
        shards[0].launch(3)                                # rank 1 never launches
        with pytest.raises(svslam.SvsError):
            shards[0].finish()
        assert 2.0 < time.time() - t0 < 60.0
    finally:
        for s in shards:
            s.close()
        c2.close()
