#!/usr/bin/env python
"""Ingest micro-benchmark (diagnostic, not a bench line): aggregate PCIe throughput of the zero-copy half-resolution
ingest (svs_frameset_push_ptrs, on_device = 2) for G concurrent contexts, and of the staged-DMA mode.
Usage: python scripts/ingest_bench.py [streams] ; env SVS_ZC_CTAS sets the persistent grid."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
import torch
import svslam

W, H = 1226, 370
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
clip = torch.from_numpy(np.random.RandomState(0).randint(0, 255, (2, 48, H, W), dtype=np.uint8)).pin_memory()
clip_d = clip.cuda()
img = W * H
for mode in (2, 0, 1):
    for G in (1, 4, 16, 32):
        ctxs = [svslam.Context(0) for _ in range(G)]
        n = B // G
        fss = [c.frameset(n, W, H, half=True) for c in ctxs]
        base = (clip_d if mode == 1 else clip).data_ptr()

        def work(g, reps):
            for r in range(reps):
                lp = [base + ((g * n + b + r) % 48) * img for b in range(n)]
                rp = [base + (48 + (g * n + b + r) % 48) * img for b in range(n)]
                fss[g].push_ptrs(lp, rp, mode)
            ctxs[g].sync()

        def run(reps):
            th = [threading.Thread(target=work, args=(g, reps)) for g in range(G)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        run(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 6
        run(reps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nbytes = reps * B * 2 * W * ((H + 1) // 2)
        print("mode %d groups %2d streams %d: %.1f ms/step  %.1f GB/s (even rows)  %.0f frames/s" % (
            mode, G, B, 1e3 * dt / reps, nbytes / dt / 1e9, reps * B / dt), flush=True)
        for f in fss:
            f.close()
        for c in ctxs:
            c.close()
