"""Profiling aid: one batched StereoBM call (config 5 shape: 1241x376, 128 disparities, block 15) for an ncu capture."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
import svslam  # noqa: E402

rng = np.random.default_rng(0)
B, H, W = 16, 376, 1241
base = rng.integers(0, 255, (B, H, W + 64), dtype=np.uint8)
left = np.ascontiguousarray(base[:, :, 40:40 + W])
right = np.ascontiguousarray(base[:, :, 52:52 + W])          # constant 12-px disparity
ctx = svslam.Context(0)
for _ in range(2):
    d = ctx.stereo_bm(left, right, 128, 15)
print("valid pixels", int((d >= 0).sum()), "median disparity/16", float(np.median(d[d >= 0]) / 16.0))
ctx.close()
