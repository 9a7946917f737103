"""World-size-2 gloo test (CPU) of the host plumbing of the landmark-sharded BA (svslam/ba_shard.py): the solver itself is
one cooperative kernel per GPU that talks to its peers through their exchange windows, so what the host does for N > 1 is
(1) split the landmarks, (2) carry every rank's 64-byte window handle to every other rank and hand the mapped pointers to
svs_ba_shard_set_peers in RANK ORDER with the rank's own window in its own slot.  A fake shard records the calls."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeShard:
    def __init__(self, rank):
        self.rank, self.peers, self.imported = rank, None, []

    def window(self):
        return 0x7000_0000 + 0x1000 * self.rank, 4096

    def export_handle(self):
        return bytes([self.rank]) * 64

    def import_handle(self, h):
        assert len(h) == 64 and h == bytes([h[0]]) * 64
        self.imported.append(h[0])
        return 0x9000_0000 + 0x1000 * h[0]            # where the peer's window is mapped in THIS process

    def set_peers(self, rank, ptrs):
        self.peers = (rank, list(ptrs))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svslam import ba_shard
    sh = FakeShard(rank)
    ba_shard.wire_distributed(sh, dist)
    out.put((rank, sh.peers, sh.imported))
    dist.barrier()
    dist.destroy_process_group()


def test_window_handles_reach_every_rank_in_rank_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r[0], r[1:]) for r in (q.get(timeout=120) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] == (0, [0x7000_0000, 0x9000_1000]) and res[0][1] == [1]      # own window in slot 0, peer 1 imported
    assert res[1][0] == (1, [0x9000_0000, 0x7000_1000]) and res[1][1] == [0]


def test_exchange_handles_rejects_a_broken_bootstrap():
    sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
    import pytest
    from svslam import ba_shard, SvsError
    mine = b"\x01" * 64
    assert ba_shard.exchange_handles(mine, 1, 2, lambda b: [b"\x00" * 64, b]) == [b"\x00" * 64, mine]
    with pytest.raises(SvsError):
        ba_shard.exchange_handles(mine, 0, 2, lambda b: [b"\x00" * 64, b])        # own handle in the wrong slot
    with pytest.raises(SvsError):
        ba_shard.exchange_handles(mine, 1, 2, lambda b: [b"\x00" * 10, b])        # truncated handle


def test_split_problem_covers_every_edge_once():
    sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from svslam import ba_shard
    from util import ba_problem
    prob, _, _ = ba_problem(3, n_kf=6, n_lm=120)
    for world in (1, 2, 3, 8):
        parts = ba_shard.split_problem(prob, world)
        assert sum(len(p["edge_kf"]) for p, _ in parts) == len(prob["edge_kf"])
        ids = np.concatenate([i for _, i in parts])
        assert sorted(ids.tolist()) == list(range(len(prob["lms"])))
        for p, i in parts:
            assert len(p["lms"]) == len(i) and (len(p["edge_lm"]) == 0 or p["edge_lm"].max() < len(i))
            assert np.array_equal(p["poses"], prob["poses"])             # every shard holds ALL poses
        loads = [len(p["edge_kf"]) for p, _ in parts]
        assert max(loads) - min(loads) <= 0.25 * max(loads) + 8
