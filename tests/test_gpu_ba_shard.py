"""Sharded / large-window BA (svs_ba_shard_*) on one GPU: one shard and several shards (their three
reductions summed on the device, exactly what the NCCL all-reduce does across GPUs) against the CPU oracle."""
import numpy as np
import pytest

from oracle import geom
from util import K05, EXT_L, EXT_R, ba_problem, ba_problem_big, rel_to_norm

pytestmark = pytest.mark.gpu


def _run(ctx, prob, world, jac_mode=0, max_iter=10):
    from svslam import ba_shard
    parts = ba_shard.split_problem(prob, world)
    shards = [ba_shard.Shard(ctx, p["poses"], p["lms"], p["edge_kf"], p["edge_lm"], p["edge_cam"], p["edge_uv"],
                             K05, K05, EXT_L, EXT_R, 5.991, jac_mode) for p, _ in parts]
    try:
        st = ba_shard.lm_optimize(shards, max_iter)
        lms = np.zeros((len(prob["lms"]), 3))
        chi2 = np.zeros(len(prob["edge_kf"]))
        owner_edge = np.zeros(len(prob["edge_kf"]), bool)
        poses = None
        for s, (p, ids) in zip(shards, parts):
            P, L, c2 = s.get()
            poses = P if poses is None else poses
            assert np.array_equal(P, poses)            # every shard holds the identical pose state
            lms[ids] = L
        return poses, lms, st
    finally:
        for s in shards:
            s.close()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_shards_match_oracle(ctx, world):
    prob, _, _ = ba_problem(5, n_kf=12, n_lm=400)
    wP, wL, wchi2, wst = geom.ba_optimize(prob["poses"], prob["lms"], prob["edge_kf"], prob["edge_lm"], prob["edge_cam"],
                                          prob["edge_uv"], K05, K05, EXT_L, EXT_R)
    P, L, st = _run(ctx, prob, world)
    assert (st["iterations"], st["trials"]) == (wst.iterations, wst.trials)
    assert abs(st["chi2_init"] - wst.chi2_init) < 1e-10 * wst.chi2_init and abs(st["chi2"] - wst.chi2) < 1e-8 * wst.chi2
    assert np.abs(P - wP).max() < 1e-8 and rel_to_norm(L, wL).max() < 1e-8


def test_config4_scale_one_vs_four_shards(ctx):
    prob = ba_problem_big(4, n_kf=50, n_lm=20000)            # config-4 shape at 1/5 of the landmarks (test time)
    P1, L1, st1 = _run(ctx, prob, 1, max_iter=4)
    P4, L4, st4 = _run(ctx, prob, 4, max_iter=4)
    assert st1["chi2"] < 0.5 * st1["chi2_init"]
    assert (st1["iterations"], st1["trials"]) == (st4["iterations"], st4["trials"])
    assert np.abs(P1 - P4).max() < 1e-9 and rel_to_norm(L1, L4).max() < 1e-9     # 1-GPU result == sharded result
