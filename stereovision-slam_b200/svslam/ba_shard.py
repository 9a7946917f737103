"""Landmark-sharded bundle adjustment (svs_ba_shard_*): thin wrappers around the C entry points.  The Levenberg-Marquardt
loop — g2o's control flow, SURVEY.md Appendix B.2 — runs entirely inside one cooperative kernel per GPU
(svs_ba_shard_optimize), and the kernels of the ranks exchange their partial reduced systems themselves through peer
memory.  What is left here is plumbing: splitting a problem by landmarks, and handing every rank the other ranks' exchange
windows (CUDA IPC handles carried by torch.distributed, or plain pointers for several shards inside one process)."""
import ctypes as C

import numpy as np

from . import _p, _f64, _i32, _u8, SvsError, BaStats


class Shard:
    def __init__(self, ctx, poses, lms, edge_kf, edge_lm, edge_cam, edge_uv, K_left, K_right, ext_left, ext_right,
                 huber_delta=5.991, jac_mode=0):
        lib = ctx.lib
        lib.svs_ba_shard_create.restype = C.c_void_p
        self.ctx = ctx
        poses, lms = _f64(poses).reshape(-1, 7), _f64(lms).reshape(-1, 3)
        ekf, elm, ecam, euv = _i32(edge_kf), _i32(edge_lm), _u8(edge_cam), _f64(edge_uv).reshape(-1, 2)
        self.N, self.L, self.E = len(poses), len(lms), len(ekf)
        self.h = lib.svs_ba_shard_create(C.c_void_p(ctx.h), self.N, _p(poses), self.L, _p(lms), self.E, _p(ekf), _p(elm), _p(ecam),
                                         _p(euv), _p(_f64(K_left)), _p(_f64(K_right)), _p(_f64(ext_left)), _p(_f64(ext_right)),
                                         C.c_double(huber_delta), jac_mode)
        if not self.h:
            raise SvsError("svs_ba_shard_create failed: " + ctx.last_error())
        self._imported = []

    def window(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        self.ctx._chk(self.ctx.lib.svs_ba_shard_window(C.c_void_p(self.h), C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def set_peers(self, rank, ptrs):
        arr = (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
        self.ctx._chk(self.ctx.lib.svs_ba_shard_set_peers(C.c_void_p(self.ctx.h), C.c_void_p(self.h), len(ptrs), rank, arr))

    def set_grid_limit(self, n):
        self.ctx._chk(self.ctx.lib.svs_ba_shard_set_grid_limit(C.c_void_p(self.h), int(n)))

    def launch(self, max_iter=10):
        self.ctx._chk(self.ctx.lib.svs_ba_shard_launch(C.c_void_p(self.ctx.h), C.c_void_p(self.h), int(max_iter)))

    def finish(self):
        st = BaStats()
        self.ctx._chk(self.ctx.lib.svs_ba_shard_finish(C.c_void_p(self.ctx.h), C.c_void_p(self.h), C.byref(st)))
        return stats_dict(st)

    def optimize(self, max_iter=10):
        self.launch(max_iter)
        return self.finish()

    def export_handle(self):
        h = (C.c_ubyte * 64)()
        self.ctx._chk(self.ctx.lib.svs_ipc_export(C.c_void_p(self.ctx.h), C.c_void_p(self.window()[0]), h))
        return bytes(h)

    def import_handle(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        ptr = C.c_void_p()
        self.ctx._chk(self.ctx.lib.svs_ipc_import(C.c_void_p(self.ctx.h), buf, C.byref(ptr)))
        self._imported.append(ptr.value)
        return ptr.value

    PHASES = ("linearise", "hpp_reduce", "vinv_wd", "schur_chunks", "payload", "exchange", "permute", "ldlt", "trial_poses",
              "backsubst_chi2", "reduce", "exchange2", "accept", "commit")

    def phase_ns(self):
        out = np.zeros(16)
        self.ctx._chk(self.ctx.lib.svs_ba_shard_phase_ns(C.c_void_p(self.ctx.h), C.c_void_p(self.h), _p(out)))
        return dict(zip(self.PHASES, out.tolist()))

    def get(self):
        poses = np.zeros((self.N, 7)); lms = np.zeros((self.L, 3)); chi2 = np.zeros(max(self.E, 1))
        self.ctx._chk(self.ctx.lib.svs_ba_shard_get(C.c_void_p(self.ctx.h), C.c_void_p(self.h), _p(poses), _p(lms), _p(chi2)))
        return poses, lms, chi2[:self.E]

    def close(self):
        if self.h:
            self.ctx.sync()
            for p in self._imported:
                self.ctx.lib.svs_ipc_release(C.c_void_p(self.ctx.h), C.c_void_p(p))
            self._imported = []
            self.ctx.lib.svs_ba_shard_destroy(C.c_void_p(self.ctx.h), C.c_void_p(self.h))
            self.h = None


def stats_dict(st):
    return dict(iterations=int(st.iterations), trials=int(st.trials), linearizations=int(st.linearizations), chi2=float(st.chi2),
                chi2_init=float(st.chi2_init), lam=float(st.lambda_))


def split_problem(prob, world):
    """Partition landmarks (with all their edges) over `world` shards, balanced by edge count.  Returns a list of
    (sub-problem with LOCAL landmark indices, global landmark ids of the shard)."""
    from .dist import partition_by_weight
    elm = np.asarray(prob["edge_lm"])
    deg = np.bincount(elm, minlength=len(prob["lms"]))
    owner = partition_by_weight(deg, world)
    out = []
    for r in range(world):
        ids = np.flatnonzero(owner == r)
        remap = np.full(len(prob["lms"]), -1, np.int64)
        remap[ids] = np.arange(len(ids))
        m = owner[elm] == r
        out.append((dict(poses=prob["poses"], lms=np.asarray(prob["lms"])[ids], edge_kf=np.asarray(prob["edge_kf"])[m],
                         edge_lm=remap[elm[m]].astype(np.int32), edge_cam=np.asarray(prob["edge_cam"])[m],
                         edge_uv=np.asarray(prob["edge_uv"])[m]), ids))
    return out


def exchange_handles(my_handle, rank, world, all_gather):
    """Every rank's 64-byte window handle, in rank order.  all_gather(bytes) -> list of bytes is the only thing asked of the
    bootstrap (torch.distributed over gloo or NCCL, MPI, ...)."""
    handles = all_gather(my_handle)
    if len(handles) != world or handles[rank] != my_handle or any(len(h) != 64 for h in handles):
        raise SvsError("window handle exchange failed")
    return handles


def wire_distributed(shard, dist):
    """One shard per rank of a torch.distributed process group (one process per GPU): map the peers' windows (CUDA IPC)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return

    def all_gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    handles = exchange_handles(shard.export_handle(), rank, world, all_gather)
    own = shard.window()[0]
    shard.set_peers(rank, [own if r == rank else shard.import_handle(handles[r]) for r in range(world)])
    dist.barrier()


def wire_local(shards, sm_count=148):
    """Several shards inside ONE process on ONE GPU (tests): plain device pointers; every shard must sit on its own context
    (its own stream) and all kernels must be resident together, so each gets an equal slice of the SMs."""
    ptrs = [s.window()[0] for s in shards]
    for r, s in enumerate(shards):
        s.set_peers(r, ptrs)
        s.set_grid_limit(max(1, sm_count // len(shards)))


def optimize_all(shards, max_iter=10):
    """Launch the cooperative solvers of all local shards, then wait for all of them; returns the (identical) statistics."""
    for s in shards:
        s.launch(max_iter)
    sts = [s.finish() for s in shards]
    return sts[0], sts
