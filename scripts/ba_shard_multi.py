"""torchrun script: landmark-sharded BA across GPUs, config-4 shape.  One cooperative solver kernel per GPU;
the kernels exchange their partial reduced systems through each other's device memory over NVLink (no NCCL in the loop;
torch.distributed only carries the 64-byte CUDA IPC handles of the exchange windows once).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/ba_shard_multi.py [n_lm]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
import numpy as np, torch, torch.distributed as dist
import svslam
from svslam import ba_shard
from svslam.problems import K05, EXT_L, EXT_R, ba_problem_big

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_lm = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
prob = ba_problem_big(4, n_kf=50, n_lm=n_lm)
ctx = svslam.Context(local)
p, ids = ba_shard.split_problem(prob, world)[rank]
res = {}
for rep in range(3):
    sh = ba_shard.Shard(ctx, p["poses"], p["lms"], p["edge_kf"], p["edge_lm"], p["edge_cam"], p["edge_uv"], K05, K05, EXT_L, EXT_R)
    if os.environ.get("SVS_BS_GRID"):
        sh.set_grid_limit(int(os.environ["SVS_BS_GRID"]))
    if world > 1:
        ba_shard.wire_distributed(sh, dist)
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = sh.optimize(10)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    P, L, c2 = sh.get()
    phases = sh.phase_ns()
    if world > 1:
        dist.barrier()          # nobody unmaps a window a peer may still read
    sh.close()
    res = dict(n_gpus=world, n_kf=50, n_lm=n_lm, n_edges=int(len(prob["edge_kf"])), local_edges=int(len(p["edge_kf"])), seconds=dt,
               lm_iterations=st["iterations"], trials=st["trials"], iters_per_sec=st["iterations"] / dt, chi2_init=st["chi2_init"],
               chi2=st["chi2"], pose_checksum=float(np.abs(P).sum()), phase_us_per_trial={k: round(v / 1e3 / max(1, st["trials"]), 1) for k, v in phases.items()})
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
