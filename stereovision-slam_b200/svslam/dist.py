"""torch.distributed plumbing for the multi-GPU runs (one process per GPU, replicas of independent streams — the
per-frame path has no data-path collective).  Works on gloo (CPU tests) and nccl."""
import os

import numpy as np


def env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def clip_starts(n_streams, rank, clip_len):
    """Start offsets of this rank's streams inside the clip: neighbouring streams and ranks are de-phased so that
    keyframes (and therefore GFTT / BA batches) are spread evenly over the steps."""
    return [((7 * b) % 24 + 5 * rank) % max(1, clip_len - 1) for b in range(n_streams)]


def partition_by_weight(weights, world):
    """Greedy longest-processing-time partition of items (e.g. landmarks weighted by their edge count) over ranks.
    Returns owner[i] in [0, world).  Deterministic; every rank computes the same answer."""
    weights = np.asarray(weights, np.int64)
    order = np.argsort(-weights, kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.zeros(len(weights), np.int32)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += weights[i]
    return owner


def max_over_ranks(value, dist=None, device="cpu"):
    """Elapsed time is reported as the MAX over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(arr, dist=None, device="cpu"):
    """Sum a float64 array over ranks (what the landmark-sharded BA does with the reduced camera system)."""
    import torch
    t = torch.as_tensor(np.asarray(arr, np.float64), device=device).clone()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
