"""The Levenberg-Marquardt driver of the landmark-sharded BA (svslam.ba_shard.lm_optimize — the host side of
svs_ba_shard_*) without a GPU: NumPy stand-ins for the shards (tests/mock_shard.py), the real driver and
its reductions, checked against the single-problem C oracle — in one process with two shards, and with one shard per rank
over torch.distributed / gloo (world size 2)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from oracle import geom
from util import ba_problem, K05, EXT_L, EXT_R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    prob, _, _ = ba_problem(5, n_kf=5, n_lm=60)
    return prob


def _shards(prob, world, only=None):
    from mock_shard import MockShard
    from svslam import ba_shard
    parts = ba_shard.split_problem(prob, world)
    out = []
    for r, (p, ids) in enumerate(parts):
        if only is None or r == only:
            out.append((MockShard(p["poses"], p["lms"], p["edge_kf"], p["edge_lm"], p["edge_cam"], p["edge_uv"], K05, K05, EXT_L, EXT_R), ids))
    return out


def _check(st, poses, lms_by_id, prob):
    po, lo, co, so = geom.ba_optimize(prob["poses"], prob["lms"], prob["edge_kf"], prob["edge_lm"], prob["edge_cam"], prob["edge_uv"],
                                      K05, K05, EXT_L, EXT_R, max_iter=10)
    assert st["iterations"] == so.iterations and st["trials"] == so.trials
    assert abs(st["chi2_init"] - so.chi2_init) < 1e-9 * so.chi2_init and abs(st["chi2"] - so.chi2) < 1e-7 * so.chi2
    assert np.abs(poses - po).max() < 1e-7
    for l, x in lms_by_id.items():
        assert np.abs(x - lo[l]).max() < 1e-6


def test_lm_driver_two_shards_one_process():
    from svslam import ba_shard
    prob = _problem()
    sh = _shards(prob, 2)
    st = ba_shard.lm_optimize([s for s, _ in sh], 10, None)
    lms = {}
    for s, ids in sh:
        P, Lm, _ = s.get()
        lms.update({int(g): Lm[k] for k, g in enumerate(ids)})
    assert len(lms) == len(prob["lms"])
    _check(st, sh[0][0].get()[0], lms, prob)
    assert np.array_equal(sh[0][0].get()[0], sh[1][0].get()[0])          # every shard holds the same poses


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svslam import ba_shard
    prob = _problem()
    (s, ids), = _shards(prob, world, only=rank)
    st = ba_shard.lm_optimize([s], 10, dist)
    P, Lm, _ = s.get()
    out.put((rank, st, P, {int(g): Lm[k] for k, g in enumerate(ids)}))
    dist.barrier()
    dist.destroy_process_group()


def test_lm_driver_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, st0, P0, l0), (r1, st1, P1, l1) = res
    assert st0 == st1 and np.array_equal(P0, P1)              # the replicated part is bitwise identical on both ranks
    lms = dict(l0); lms.update(l1)
    _check(st0, P0, lms, _problem())
