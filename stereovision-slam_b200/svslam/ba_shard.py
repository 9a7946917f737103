"""Driver of the landmark-sharded bundle adjustment (svs_ba_shard_*): g2o's Levenberg-Marquardt control loop
 in the caller, with the three per-trial reductions done by `allreduce` — NCCL through
torch.distributed for one shard per GPU, a plain sum for several shards on one GPU (tests), identity for one shard.
The buffers that are reduced live in torch CUDA tensors; the kernels write into them through their data_ptr."""
import ctypes as C

import numpy as np
import torch

from . import _p, _f64, _i32, _u8, SvsError


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
        dev = torch.device("cuda", torch.cuda.current_device())
        self.lin = torch.zeros(42 * self.N + 2, dtype=torch.float64, device=dev)
        self.maxdiag = torch.zeros(1, dtype=torch.float64, device=dev)
        self.red = torch.zeros(36 * self.N * self.N + 6 * self.N, dtype=torch.float64, device=dev)
        self.flag = torch.ones(1, dtype=torch.int32, device=dev)
        self.tri = torch.zeros(4, dtype=torch.float64, device=dev)
        # the fills above run on torch's current stream, the svs_ba_shard_* kernels on the context's own non-blocking stream:
        # nothing else orders the two, so the fills must have landed before the first kernel writes these buffers
        torch.cuda.synchronize(dev)

    def _v(self, t):
        return C.c_void_p(t.data_ptr())

    def linearize(self):
        self.ctx._chk(self.ctx.lib.svs_ba_shard_linearize(C.c_void_p(self.ctx.h), C.c_void_p(self.h), self._v(self.lin), self._v(self.maxdiag)))

    def schur(self, lam):
        self.ctx._chk(self.ctx.lib.svs_ba_shard_schur(C.c_void_p(self.ctx.h), C.c_void_p(self.h), C.c_double(lam), self._v(self.red), self._v(self.flag)))

    def try_step(self, lin, red, lam, flag_ok):
        self.ctx._chk(self.ctx.lib.svs_ba_shard_try(C.c_void_p(self.ctx.h), C.c_void_p(self.h), self._v(lin), self._v(red), C.c_double(lam),
                                                    int(flag_ok), self._v(self.tri)))

    def accept(self):
        self.ctx._chk(self.ctx.lib.svs_ba_shard_accept(C.c_void_p(self.ctx.h), C.c_void_p(self.h)))

    def get(self):
        poses = np.zeros((self.N, 7)); lms = np.zeros((self.L, 3)); chi2 = np.zeros(max(self.E, 1))
        self.ctx._chk(self.ctx.lib.svs_ba_shard_get(C.c_void_p(self.ctx.h), C.c_void_p(self.h), _p(poses), _p(lms), _p(chi2)))
        return poses, lms, chi2[:self.E]

    def close(self):
        if self.h:
            self.ctx.lib.svs_ba_shard_destroy(C.c_void_p(self.ctx.h), C.c_void_p(self.h))
            self.h = None


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


def lm_optimize(shards, max_iter=10, dist=None):
    """Levenberg-Marquardt over one or more shards.  `shards` = the Shard objects THIS process owns (several on one
    GPU in tests, exactly one per rank under torch.distributed).  Returns stats dict; state stays in the shards."""
    def allsum(tensors):
        t = tensors[0] if len(tensors) == 1 else torch.stack(tensors).sum(0)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def allmax(tensors):
        t = tensors[0] if len(tensors) == 1 else torch.stack(tensors).max(0).values
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def allmin(tensors):
        t = tensors[0] if len(tensors) == 1 else torch.stack(tensors).min(0).values
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return t

    ctx = shards[0].ctx
    N = shards[0].N
    lam, ni = 0.0, 2.0
    st = dict(iterations=0, trials=0, chi2=0.0, chi2_init=0.0, lam=0.0)
    for it in range(max_iter):
        for s in shards:
            s.linearize()
        ctx.sync()
        lin = allsum([s.lin for s in shards])
        cur = float(lin[42 * N].item())
        if it == 0:
            st["chi2_init"] = cur
            # the pose diagonal must be taken from the SUMMED Hpp, the landmark diagonals are local
            md_l = float(allmax([s.maxdiag for s in shards]).item())
            hpp = lin[:36 * N].view(N, 6, 6)
            md_p = float(torch.diagonal(hpp, dim1=1, dim2=2).abs().max().item())
            lam, ni = 1e-5 * max(md_l, md_p), 2.0
        rho, q = 0.0, 0
        while True:
            for s in shards:
                s.schur(lam)
            ctx.sync()
            red = allsum([s.red for s in shards])
            ok = int(allmin([s.flag for s in shards]).item())
            for s in shards:
                s.try_step(lin, red, lam, ok)
            ctx.sync()
            loc = allsum([s.tri[:2] for s in shards]).cpu().numpy()
            pose_scale, solved = float(shards[0].tri[2].item()), float(shards[0].tri[3].item()) != 0.0
            tmp = float(loc[0]) if solved else np.finfo(np.float64).max
            scale = pose_scale + float(loc[1]) + 1e-3
            rho = (cur - tmp) / scale
            q += 1
            st["trials"] += 1
            if rho > 0 and np.isfinite(tmp):
                alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                lam *= max(1.0 / 3.0, alpha)
                ni = 2.0
                cur = tmp
                for s in shards:
                    s.accept()
            else:
                lam *= ni
                ni *= 2
            if not (rho < 0 and q < 10):
                break
        st["iterations"] += 1
        st["chi2"], st["lam"] = cur, lam
        if q == 10 or rho == 0:
            break
    ctx.sync()
    return st
