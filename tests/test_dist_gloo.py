"""World-size-2 gloo test (CPU) of the multi-GPU host logic: replica de-phasing, max-over-ranks timing, and the
landmark partition + all-reduce identity the sharded BA relies on (sum of per-shard reduced systems == full system)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svslam import dist as sd
    from oracle import geom
    from util import ba_problem, K05, EXT_L, EXT_R
    r, w, _ = sd.env()
    assert (r, w) == (rank, world)
    ms = sd.max_over_ranks(10.0 + rank, dist)
    # landmark-sharded BA identity: each rank linearises only its landmarks' edges; the all-reduced chi2 and the
    # one-iteration result of the union must equal the single-process oracle
    prob, _, _ = ba_problem(3, n_kf=6, n_lm=120)
    deg = np.bincount(prob["edge_lm"], minlength=len(prob["lms"]))
    owner = sd.partition_by_weight(deg, world)
    mine = owner[prob["edge_lm"]] == rank
    sub = {k: (v[mine] if k.startswith("edge_") else v) for k, v in prob.items()}
    _, _, _, st = geom.ba_optimize(sub["poses"], sub["lms"], sub["edge_kf"], sub["edge_lm"], sub["edge_cam"], sub["edge_uv"],
                                   K05, K05, EXT_L, EXT_R, max_iter=1)
    tot = sd.sum_over_ranks([st.chi2_init, float(mine.sum())], dist)
    if rank == 0:
        _, _, _, full = geom.ba_optimize(prob["poses"], prob["lms"], prob["edge_kf"], prob["edge_lm"], prob["edge_cam"],
                                         prob["edge_uv"], K05, K05, EXT_L, EXT_R, max_iter=1)
        out.put((ms, tot.tolist(), full.chi2_init, len(prob["edge_kf"]), sd.clip_starts(4, 0, 48), owner.tolist()))
    else:
        out.put((ms, sd.clip_starts(4, 1, 48), owner.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = [x for x in res if len(x) == 6][0]
    r1 = [x for x in res if len(x) == 3][0]
    assert r0[0] == 11.0 and r1[0] == 11.0                       # max over ranks
    assert abs(r0[1][0] - r0[2]) < 1e-9 * r0[2]                  # sum of per-shard chi2 == full chi2
    assert r0[1][1] == r0[3]                                     # every edge owned exactly once
    assert r0[4] != r1[1]                                        # ranks are de-phased
    assert r0[5] == r1[2]                                        # identical partition on every rank
    loads = np.bincount(np.array(r0[5]), minlength=2)
    assert abs(int(loads[0]) - int(loads[1])) <= 60
