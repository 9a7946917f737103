#!/usr/bin/env python
"""bench.py — stereo frames/s of the StereoVision-SLAM hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] — "KITTI seq-05 full pipeline (Frontend + Backend BA,
loop closure off), 1xB200": synthetic 1226x370 stereo pairs (no KITTI data is available offline) from the
ray-cast corridor generator with the seq-05 calibration, processed at the reference's half resolution
613x185, num_features 150, BA window 10, synchronous BA schedule.  One STEP = one Frontend::AddFrame for every
one of `--streams` independent stereo streams (a batch of stereo pairs), i.e. `streams` frames.

  value  frames/s with the input images already resident in HBM
  e2e    frames/s through the public C ABI with the images in pinned HOST memory (H2D copy of every frame
         inside the timed region; poses / statuses are read back to the host every step)

`--impl reference` times the reference's CPU path instead: the OpenCV stages through cv2 (the library the
reference calls, with its exact arguments) and the g2o blocks through the C restatement in oracle/geom.c,
one process per host core, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# OpenMP teams of several pipelines / ranks share the host cores: never spin-wait (must be set before libgomp loads)
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
os.environ.setdefault("GOMP_SPINCOUNT", "0")
# One hardware work queue per CUDA stream (16 context groups x (main + ingest stream)): with the default of 8 connections
# several streams share a queue and a group's kernels wait behind ANOTHER group's long PCIe-bound ingest kernel
# (head-of-line blocking).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))

METRIC = "stereo_frames_per_sec"
UNIT = "frames/s"
CALIB = "kitti05"
WORKLOAD = ("KITTI seq-05-shaped full pipeline (Frontend GFTT+LK+triangulation+pose-LM, Backend BA window 10, "
            "loop closure off), synthetic 1226x370 stereo pairs processed at 613x185")


def log(*a):
    if os.environ.get("SVS_BENCH_VERBOSE", "1") != "0":
        print("[bench %.1fs]" % (time.perf_counter() - _T0), *a, file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def pingpong(i, n):
    """Play a clip forward then backward (a physically valid camera motion) so any number of steps can be run."""
    p = i % (2 * n - 2)
    return p if p < n else 2 * n - 2 - p


def make_clip(n_frames):
    from svslam import synth
    cor = synth.Corridor(CALIB, seed=5, n_frames=n_frames)
    L, R, T = cor.sequence(n_frames)
    return cor, L, R, T


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  Default: NVML in-process
    (nvidia_ml_py) every 200 ms — the same counters `nvidia-smi --query-gpu=clocks.sm,...,clocks_event_reasons.*` prints,
    without a second process hammering the driver; `--sampler smi` uses the nvidia-smi command line of the recipe."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu, mode="nvml", period=0.2, uuid=None):
        self.gpu, self.mode, self.period, self.uuid = gpu, mode, period, uuid
        self.rows, self.proc, self.t, self.stop_flag = [], None, None, False
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []

    def start(self):
        if self.mode == "none":
            return
        if self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                self.nv = pynvml
                self.h = None
                if self.uuid:      # CUDA_VISIBLE_DEVICES may renumber the devices: NVML is addressed by UUID
                    try:
                        self.h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode() if isinstance(self.uuid, str) else self.uuid)
                    except Exception:
                        self.h = None
                if self.h is None:
                    self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
                self.t = threading.Thread(target=self._poll, daemon=True)
                self.t.start()
                return
            except Exception:
                self.mode = "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for n, bit in names:
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) < 9:
                continue
            try:
                self.sm.append(float(r[1])); self.mx.append(float(r[2])); self.power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        if self.mode == "none" or self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler unavailable" if self.mode != "none" else "sampler off"]}
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "power_w": float(np.median(self.power)) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "sampler": self.mode}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_stream_proc(idx, clip_path, prime, conn):
    """One independent stream on one host core: the reference's CPU path (persistent: the pipeline is primed once so
    that its BA window is full, like the GPU arm's streams, then it runs `n` more frames per request)."""
    import cv2
    cv2.setNumThreads(1)
    from oracle import pipeline as op
    from svslam import synth
    d = np.load(clip_path)
    L, R = d["L"], d["R"]
    cal = synth.CALIB[CALIB]
    K = np.array([cal[2] * 0.5, cal[2] * 0.5, cal[3] * 0.5, cal[4] * 0.5])
    p = op.Pipeline(K, cal[5], op.Cfg(backend_on=1), stages="cv2", cv2=cv2)
    nclip = len(L)
    cur = (7 * idx) % 24
    for _ in range(prime):
        j = pingpong(cur, nclip); cur += 1
        p.add_frame(L[j], R[j])
    conn.send(("ready", len(p.active_kfs)))
    while True:
        n = conn.recv()
        if n <= 0:
            break
        t0 = time.perf_counter()
        kfs = 0
        for _ in range(n):
            j = pingpong(cur, nclip); cur += 1
            p.add_frame(L[j], R[j])
            kfs += int(p.is_kf)
        conn.send((time.perf_counter() - t0, n, kfs, p.status))


class CpuStreams:
    """n_procs persistent CPU streams (one process per core)."""

    def __init__(self, clip_path, n_procs, prime):
        import multiprocessing as mp
        from oracle import geom
        geom.build()
        ctx = mp.get_context("spawn")
        self.procs, self.conns = [], []
        for i in range(n_procs):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_cpu_stream_proc, args=(i, clip_path, prime, b), daemon=True)
            pr.start()
            self.procs.append(pr); self.conns.append(a)
        self.window = [c.recv()[1] for c in self.conns]     # wait until every stream is primed

    def step(self, n):
        """Every stream runs n frames concurrently -> (aggregate frames/s, frames, keyframes)."""
        for c in self.conns:
            c.send(n)
        res = [c.recv() for c in self.conns]
        frames = sum(r[1] for r in res)
        return frames / max(r[0] for r in res), frames, sum(r[2] for r in res)

    def close(self):
        for c in self.conns:
            try:
                c.send(0)
            except Exception:
                pass
        for pr in self.procs:
            pr.join(timeout=5)


def save_clip(L, R):
    path = "/tmp/svslam_bench_clip_%d.npz" % os.getpid()
    np.savez(path, L=L, R=R)
    return path


CPU_PRIME = int(os.environ.get("SVS_CPU_PRIME", "150"))     # frames each CPU stream runs before it is timed (BA window full)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cor, L, R, T = make_clip(args.clip_frames)
    path = save_clip(L, R)
    # each step = every core runs `n` frames of its own (primed, steady-state) stream; n is sized so that the whole
    # warm-up + K steps run stays within a few minutes whatever K is
    n = args.cpu_frames if args.cpu_frames > 0 else max(2, min(40, 2400 // (args.warmup + args.steps)))
    cs = CpuStreams(path, cores, CPU_PRIME)
    res = []
    for s in range(args.warmup + args.steps):
        r = cs.step(n)
        if s >= args.warmup:
            res.append(r)
    cs.close()
    os.remove(path)
    frames = sum(r[1] for r in res)
    busy = sum(r[1] / r[0] for r in res)
    value = frames / busy
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * busy / max(1, len(res)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cpu_frames_per_stream_per_step": n, "processes": cores, "priming_frames": CPU_PRIME,
                   "active_keyframes_after_priming": int(np.median(cs.window))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d processes x %d frames x %d steps after %d priming frames each (BA window full); OpenCV stages "
                                   "through cv2 %s (the library the reference calls), g2o blocks through oracle/geom.c (g2o is not "
                                   "installable here)" % (cores, n, args.steps, CPU_PRIME, __import__("cv2").__version__)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ba_config4(ctx, dist, rank, world, dev):
    """BA LM iterations/s at BASELINE config-4 scale (N = 50 keyframes, L = 1e5 landmarks, ~5e5 edges), landmarks
    sharded over the ranks with an NCCL all-reduce of the reduced camera system per LM trial."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from svslam import ba_shard
    from util import K05, EXT_L, EXT_R, ba_problem_big
    prob = ba_problem_big(4, n_kf=50, n_lm=100000)
    p, _ = ba_shard.split_problem(prob, world)[rank]
    best = None
    for rep in range(3):
        sh = ba_shard.Shard(ctx, p["poses"], p["lms"], p["edge_kf"], p["edge_lm"], p["edge_cam"], p["edge_uv"], K05, K05, EXT_L, EXT_R)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        st = ba_shard.lm_optimize([sh], 10, dist)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        sh.close()
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    return {"n_kf": 50, "n_lm": 100000, "n_edges": int(len(prob["edge_kf"])), "shards": world, "lm_iterations": st["iterations"],
            "trials": st["trials"], "seconds": dt, "lm_iterations_per_sec": st["iterations"] / dt, "chi2_init": st["chi2_init"],
            "chi2": st["chi2"], "allreduce_bytes_per_trial": (36 * 50 * 50 + 6 * 50 + 2) * 8 if world > 1 else 0}


def run_gpu(args, rank, world, local_rank):
    import torch
    import svslam
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    # one context group per host core this rank can count on: a group's driver thread does the per-stream bookkeeping of its
    # streams and spin-waits on its CUDA stream in between, so more groups than cores only adds contention
    G = max(1, min(args.groups, args.streams, max(1, cores // max(1, local_world))))
    ctxs = [svslam.Context(dev) for _ in range(G)]   # raises if libsvslam.so / a B200 is missing: no fallback
    lib = ctxs[0].lib
    lib.svs_kernel_name.restype = C.c_char_p
    B = args.streams
    gsz = [B // G + (1 if g < B % G else 0) for g in range(G)]
    goff = np.concatenate([[0], np.cumsum(gsz)]).astype(int)
    host_threads = max(1, cores // (max(1, local_world) * G))
    log("rendering clip ...")
    cor, L, R, T = make_clip(args.clip_frames)
    log("clip ready")
    nclip = len(L)
    Kh = cor.K_half()
    img_bytes = cor.W * cor.H

    # inputs: the clip lives once in HBM (value) and once in pinned host memory (e2e)
    Ld = torch.from_numpy(L).cuda(dev); Rd = torch.from_numpy(R).cuda(dev)
    Lh = torch.from_numpy(L).pin_memory(); Rh = torch.from_numpy(R).pin_memory()
    from svslam import dist as sdist
    starts = sdist.clip_starts(B, rank, nclip)

    def ptrs(base_l, base_r, step, g):
        idx = [pingpong(starts[b] + step, nclip) for b in range(goff[g], goff[g + 1])]
        return [base_l + j * img_bytes for j in idx], [base_r + j * img_bytes for j in idx]

    # ctypes pointer arrays per (memory, group, clip phase), built once: the per-step Python work of the 16 driver threads
    # would otherwise serialise on the interpreter lock (measured: ~6 ms of a 30 ms step)
    period = 2 * nclip - 2
    ptr_cache = {}

    def ptr_arrays(on_device, step, g):
        key = (bool(on_device), g, step % period)
        v = ptr_cache.get(key)
        if v is None:
            bl, br = (Ld.data_ptr(), Rd.data_ptr()) if on_device else (Lh.data_ptr(), Rh.data_ptr())
            lp, rp = ptrs(bl, br, step, g)
            v = ptr_cache[key] = (svslam.Slam.ptr_array(lp), svslam.Slam.ptr_array(rp))
        return v

    for on_dev in (True, False):
        for g in range(G):
            for ph in range(period):
                ptr_arrays(on_dev, ph, g)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ONE set of pipelines, primed once (untimed) until every stream's sliding BA window is full, then timed in several
    # regions that only differ in where the frames live / whether launches are instrumented.  Every region does its own
    # W warm-up steps first; the stream state simply continues from region to region (steady state).
    slams = [ctxs[g].slam(gsz[g], cor.W, cor.H, Kh, cor.baseline, half=True, backend_on=1, lazy_right_ingest=0 if args.eager_right else 1)
             for g in range(G)]
    for s in slams:
        s.set_threads(host_threads)
    cursor = [0]
    # Independent streams do not insert keyframes in phase; the synthetic streams would (same age, same speed), which
    # makes the BA load arrive in bursts of one step in ~40.  Group g therefore runs gstag[g] extra untimed steps first,
    # spreading the groups' keyframe phases over one keyframe period.
    gstag = [(g * args.stagger) // G for g in range(G)]

    def run_steps(n, on_device, stagger=False):
        bl, br = (Ld.data_ptr(), Rd.data_ptr()) if on_device else (Lh.data_ptr(), Rh.data_ptr())
        # e2e ingest: 2 = zero-copy kernel reads of the pinned host frames, 0 = staged strided DMA copies,
        # 3 = mixed (even groups zero-copy, odd groups DMA) so SM-initiated reads and the copy engines share PCIe
        def mode_of(g):
            if on_device:
                return 1
            return args.h2d_mode if args.h2d_mode != 3 else (2 if g % 2 == 0 else 0)
        errors = []
        lo = cursor[0]

        def loop(g):
            try:
                first = lo + (0 if stagger else gstag[g])
                last = lo + gstag[g] + n
                nxt = ptr_arrays(on_device, first, g)
                mode = mode_of(g)
                for s in range(first, last):
                    lp, rp = nxt
                    nxt = ptr_arrays(on_device, s + 1, g)
                    if args.no_prefetch:
                        slams[g].add_frames_arrays(lp, rp, mode)
                    else:   # double-buffered ingest: frame s+1 crosses PCIe / is resized while frame s is tracked
                        slams[g].add_frames_arrays(lp, rp, mode, nxt[0], nxt[1])
            except Exception as e:      # surface worker failures in the main thread
                errors.append(e)

        th = [threading.Thread(target=loop, args=(g,)) for g in range(G)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        cursor[0] += n
        if errors:
            raise errors[0]

    def ba_host():
        tot = np.zeros(3)
        for c in ctxs:
            o = np.zeros(3)
            lib.svs_ba_host_seconds(C.c_void_p(c.h), o.ctypes.data_as(C.c_void_p))
            tot += o
        return tot

    def timed_region(on_device, timing, profile_window=False):
        run_steps(args.warmup, on_device)
        for c in ctxs:
            lib.svs_kernel_timing_reset(C.c_void_p(c.h))
            lib.svs_kernel_timing_enable(C.c_void_p(c.h), 1 if timing else 0)
        cn0 = [s.counters() for s in slams]
        bh0 = ba_host()
        l0 = sum(c.launch_count() for c in ctxs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile_window:     # `ncu --profile-from-start off`: only the steady-state steps below are captured
            torch.cuda.profiler.start()
        t0 = time.perf_counter()
        e0.record()
        run_steps(args.steps, on_device)
        torch.cuda.synchronize(dev)
        e1.record()
        if profile_window:
            torch.cuda.profiler.stop()
        e1.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = e0.elapsed_time(e1)
        log("timed region done (on_device=%s, instrumented=%s): %.1f ms/step" % (on_device, timing, ms / args.steps))
        launches = sum(c.launch_count() for c in ctxs) - l0
        cn1 = [s.counters() for s in slams]
        lost = int(sum((s.status == 3).sum() for s in slams))
        kern = {}
        for c in ctxs:
            nk = lib.svs_kernel_timing_get(C.c_void_p(c.h), None, None, 0)
            kms = np.zeros(nk); kcnt = np.zeros(nk, np.int64)
            lib.svs_kernel_timing_get(C.c_void_p(c.h), kms.ctypes.data_as(C.c_void_p), kcnt.ctypes.data_as(C.c_void_p), nk)
            lib.svs_kernel_timing_enable(C.c_void_p(c.h), 0)
            for i in range(nk):
                if kcnt[i]:
                    name = lib.svs_kernel_name(i).decode()
                    a = kern.get(name, (0.0, 0))
                    kern[name] = (a[0] + float(kms[i]), a[1] + int(kcnt[i]))
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        phases = {k: sum(c1[0][k] - c0[0][k] for c0, c1 in zip(cn0, cn1)) for k in cn1[0][0]}
        bh1 = ba_host()
        phases.update({"ba:host_build": bh1[0] - bh0[0], "ba:pack_enqueue": bh1[1] - bh0[1], "ba:device_wait_unpack": bh1[2] - bh0[2]})
        counts = {k: sum(c1[1][k] - c0[1][k] for c0, c1 in zip(cn0, cn1)) for k in cn1[0][1]}
        return dict(ms=ms, wall=wall, launches=launches, phases=phases, counts=counts, kern=kern, lost=lost)

    # priming (untimed, before any warm-up): run until the sliding BA window of every stream is full so that the timed
    # steps see steady-state problem sizes (10 keyframes); timing a cold pipeline would overstate.  The priming steps
    # also grow every grow-only device / pinned buffer to its steady-state size (regrowing is a device-wide sync).
    run_steps(args.priming, True, stagger=True)
    log("primed %d steps (+ up to %d per group to stagger the keyframe phases)" % (args.priming, max(gstag)))
    if args.profile_window:   # profiling aid (never a bench number): one device-resident region inside a profiler window
        timed_region(True, False, profile_window=True)
        log("profile window done")
        return
    run_steps(args.warmup, False)     # the e2e path's own buffers (pointer tables, staging) exist before anything is timed
    # three regions over the same workload: (1) frames resident in HBM -> value, (2) frames in pinned host memory -> e2e,
    # both with the clock sampler running and no per-kernel instrumentation; (3) device-resident again with every launch
    # bracketed by CUDA events on its stream -> per-kernel durations for the roofline block (not used for value / e2e)
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(dev, args.sampler, uuid=gpu_uuid)
    sampler.start()
    dev_pass = timed_region(True, False)
    e2e_pass = timed_region(False, False)
    clocks = sampler.stop()
    kern_pass = timed_region(True, True) if not args.no_kernel_pass else dev_pass
    diag = None
    if args.diag:      # repeat the two headline regions without the sampler: sampler / ordering sensitivity
        d2, e2 = timed_region(True, False), timed_region(False, False)
        diag = {"dev_ms_per_step_again": d2["ms"] / args.steps, "e2e_ms_per_step_again": e2["ms"] / args.steps,
                "e2e_phase_seconds_again": {k: round(v, 4) for k, v in e2["phases"].items()}}
    for s in slams:
        s.close()

    # accuracy beside the speed (BASELINE.json: "ATE vs reference"): one stream over the clip's forward pass, ATE against the
    # generator's ground truth (the oracle pipeline's ATE on the same frames is asserted equal within 3 cm in tests/)
    ate = None
    if rank == 0:
        try:
            from svslam import kitti
            one = ctxs[0].slam(1, cor.W, cor.H, Kh, cor.baseline, half=True, backend_on=1)
            est = [one.add_frames(L[i:i + 1], R[i:i + 1])[0].copy() for i in range(nclip)]
            lost = int(one.status[0] == 3)
            one.close()
            ce, _ = kitti.pose7_to_Twc(np.array(est))
            cg, _ = kitti.pose7_to_Twc(np.asarray(T)[:nclip])
            ate = {"ate_rmse_m": kitti.ate_rmse(ce, cg), "frames": nclip, "path_m": float(np.linalg.norm(np.diff(cg, axis=0), axis=1).sum()),
                   "lost": lost}
        except Exception as e:
            ate = {"error": repr(e)}
    ba4 = None
    if not args.no_ba4:
        ba4 = run_ba_config4(ctxs[0], dist, rank, world, dev)
    frames_total = B * args.steps * world
    value = frames_total / (dev_pass["ms"] * 1e-3)
    e2e = frames_total / (e2e_pass["ms"] * 1e-3)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of device time in the timed region)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:      # driver-written; tolerate a nested layout or a differently spelled key
            def find(o):
                if isinstance(o, dict):
                    for k, v in o.items():
                        if isinstance(v, (int, float)) and "hbm" in k.lower() and ("gb" in k.lower() or "bw" in k.lower() or "band" in k.lower()):
                            return float(v), k
                    for v in o.values():
                        r = find(v)
                        if r:
                            return r
                return None
            r = find(json.load(open(peaks_path)))
            if r and r[0] > 100.0:
                peak, peak_src = r[0], "measured (MEASURED_PEAKS.json %s)" % r[1]
        except Exception:
            pass
    kern = kern_pass["kern"]
    # dominant kernel = largest share of SERIALISED device time in the committed ncu launch list (profiles/); the summed
    # event-bracketed durations of the instrumented region are inflated by queueing behind other groups' kernels (they add up
    # to ~10x the wall time), so they only break ties when no launch list is present
    dom, ncu_share = None, {}
    try:
        for line in open(os.path.join(ROOT, "profiles", "r01_launches_steady_256streams.csv")):
            f = line.strip().split(",")
            if len(f) == 6 and f[0] != "kernel" and not line.startswith("#"):
                ncu_share[f[0].split("<")[0]] = float(f[5])
        dom = max((k for k in ncu_share if k in kern), key=lambda k: ncu_share[k], default=None)
    except Exception:
        pass
    if dom is None:
        dom = max(kern, key=lambda k: kern[k][0]) if kern else None
    cnt = kern_pass["counts"]
    P = 613 * 185
    frames = cnt["frames"]
    # ALGORITHMIC bytes moved by each kernel class over the whole timed region (DESIGN.md §4 / SURVEY.md §8d per-unit
    # figures x the units the region processed); achieved = that / (sum of the class's launch durations)
    alg_total = {
        "k_half_nearest": 3.0 * P * 2 * frames,
        "k_pyr_down": 1.64 * P * 2 * frames,
        "k_corner_response": 5.0 * P * cnt["keyframes"],
        "k_corner_select": 5.0 * P * cnt["keyframes"],
        "k_lk_track": 3700.0 * cnt["lk_points"],
        "k_pose_only_lm": 40.0 * cnt["pose_edges"] * 56.0,            # ~56 LM trials per problem (4 rounds x 10 it + retries)
        "k_ba_window": (316.0 * cnt["ba_edges"] + 216.0 * cnt["ba_lms"] + 576.0 * 10 * cnt["ba_kfs"]) *
                       (cnt["ba_trials"] / max(1, cnt["ba_problems"])),
        "k_triangulate": 41.0 * cnt["keyframes"] * 150,
    }
    roof = None
    if dom and dom in alg_total:
        ms_tot, n_l = kern[dom]
        ach = alg_total[dom] / (ms_tot * 1e-3) / 1e9
        limiter = {"k_ba_window": "dependent FP64 + L2 latency, one 512-thread CTA per window (ncu: 23 % issue-active, 18 % FP64 pipe, 0.02 % DRAM)",
                   "k_pose_only_lm": "dependent FP64 latency, one warp per problem (ncu: 22 % issue-active, 20 % FP64 pipe)",
                   "k_lk_track": "integer instruction issue (ncu: 79 % issue-active, 0.7 % DRAM)"}.get(dom, "HBM streaming")
        # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/), scaled by the units per launch
        units = {"k_ba_window": cnt["ba_problems"], "k_lk_track": cnt["lk_points"], "k_pose_only_lm": cnt["frames"]}.get(dom)
        traffic, traffic_src = None, None
        try:
            import csv
            for row in csv.DictReader(l for l in open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.csv")) if not l.startswith("#")):
                if row["kernel"].replace("void ", "").split("<")[0] == dom and units:
                    per_launch = (float(row["dram_rd_MB"]) + float(row["dram_wr_MB"])) * 1e6
                    grid = float(row["grid"])
                    per_unit = per_launch / (grid * {"k_ba_window": 1, "k_lk_track": 4, "k_pose_only_lm": 4}[dom])
                    traffic = per_unit * units / max(1, n_l)
                    traffic_src = "profiles/r01_ncu_full_summary.csv: %.0f B per %s x %.1f per launch" % (
                        per_unit, {"k_ba_window": "window", "k_lk_track": "keypoint", "k_pose_only_lm": "problem"}[dom], units / max(1, n_l))
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": ms_tot / max(1, n_l), "launches": n_l,
                "algorithmic_bytes_per_launch": alg_total[dom] / max(1, n_l), "peak_source": peak_src,
                "note": "launch durations overlap across %d context groups; actual limiter: %s (DESIGN.md §4)" % (G, limiter)}
    # per-kernel roofline view: (a) live — algorithmic bytes / summed event-bracketed durations of the instrumented region
    # (inflated by queueing behind the other groups' kernels), (b) stand-alone — the committed ncu capture of one context
    kernel_roofline = {}
    for k, (ms_k, n_k) in kern.items():
        if k in alg_total and ms_k > 0:
            a = alg_total[k] / (ms_k * 1e-3) / 1e9
            kernel_roofline[k] = {"live_gbs": round(a, 2), "live_frac": round(a / peak, 5), "launches": n_k}
    try:
        import csv as _csv
        for row in _csv.DictReader(l for l in open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.csv")) if not l.startswith("#")):
            k = row["kernel"].replace("void ", "").split("<")[0]
            kernel_roofline.setdefault(k, {}).setdefault("ncu_standalone", {
                "dur_us": float(row["dur_us"]), "dram_MB": round(float(row["dram_rd_MB"]) + float(row["dram_wr_MB"]), 3),
                "dram_pct": float(row["dram_pct"]), "issue_active_pct": float(row["issue_active_pct"]),
                "fp64_pipe_pct": float(row["fp64_pipe_pct"]), "registers": float(row["regs"])})
    except Exception:
        pass
    dev_total = sum(v[0] for v in kern.values())
    shares = {k: round(v[0] / dev_total, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])} if dev_total else {}

    # ---- CPU baseline on this box (bounded sample)
    cpu = None
    if not args.no_cpu_baseline:
        # in a clean subprocess (no CUDA / OpenMP state inherited), bounded by its own timeout
        log("cpu baseline subprocess ...")
        path = save_clip(L, R)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-clip", path, "--cpu-frames", str(args.cpu_frames)],
                               capture_output=True, text=True, timeout=240)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:      # the GPU line is still valid without it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "cpu baseline failed: %r" % (e,)}
        finally:
            os.remove(path)

    # image bytes the engine actually read from host memory in the e2e region (counted by the frame sets): the even rows of
    # every left frame + the even rows of the right frame of the streams that inserted a keyframe in that step
    h2d = int(e2e_pass["counts"]["h2d_image_bytes"] // args.steps)
    # image rows one step actually reads per GPU: the even rows of every left frame (+ a few right frames at keyframes)
    in_bytes = (1 if not args.eager_right else 2) * cor.W * ((cor.H + 1) // 2) * B
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_pass["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": B, "context_groups": G, "frames_per_step": B * world, "num_features": 150,
                   "num_active_keyframes": 10, "ba": "synchronous, analytic Jacobians",
                   "ingest": ("per-step push" if args.no_prefetch else "double-buffered: frame t+1 is ingested on a second stream during step t") +
                             ("; both eyes of every frame" if args.eager_right else
                              "; right images are ingested lazily, only for the streams that insert a keyframe in the step "
                              "(the frontend reads the right image nowhere else; results are bit-identical)"), "clip_frames": nclip,
                   "priming_steps": args.priming, "group_stagger_steps": args.stagger,
                   "l2": "per-step image rows read %.0f MB per GPU > 126 MB L2 (inputs larger than L2; plus ~0.6 GB of pyramids written "
                         "and re-read per step)" % (in_bytes / 1e6)
                   if in_bytes > 126e6 else "per-step image rows read %.0f MB per GPU (< L2; distinct frames every step)" % (in_bytes / 1e6)},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * (56 + 12),
                "right_images_per_step": e2e_pass["counts"]["right_images"] / args.steps,
                "h2d": {2: "zero-copy: the resize kernel reads the pinned host frames over PCIe", 0: "staged strided DMA copies of the even rows", 3: "mixed: even context groups zero-copy, odd groups strided DMA"}[args.h2d_mode],
                "ms_per_step": e2e_pass["ms"] / args.steps},
        "gpu_launches": int(dev_pass["launches"]),
        "roofline": roof,
        "cpu_baseline": cpu,
        "detail": {"phase_seconds": {k: round(v, 4) for k, v in dev_pass["phases"].items()},
                   "e2e_phase_seconds": {k: round(v, 4) for k, v in e2e_pass["phases"].items()},
                   "counts": dev_pass["counts"], "kernel_ms": {k: [round(v[0], 3), v[1]] for k, v in kern.items()},
                   "kernel_time_share": shares, "ncu_serialized_share": ncu_share, "kernel_roofline": kernel_roofline, "kernel_pass_ms_per_step": kern_pass["ms"] / args.steps,
                   "summed_kernel_ms_over_wall_ms": dev_total / kern_pass["ms"] if kern_pass["ms"] else None,
                   "lost_streams": dev_pass["lost"], "host_cores": cores, "host_threads_per_group": host_threads,
                   "ba_lm_iterations_per_sec": dev_pass["counts"]["ba_iterations"] * world / (dev_pass["ms"] * 1e-3),
                   "kernel_pass_phase_seconds": {k: round(v, 4) for k, v in kern_pass["phases"].items()},
                   "diag": diag, "accuracy": ate, "ba_config4": ba4},
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60,
                    help="timed steps per region (the 16 context groups run unsynchronised and a region ends with the slowest "
                         "group, so short regions measure the luckiest / unluckiest keyframe phase rather than the mean)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("SVS_BENCH_STREAMS", "4096")))
    ap.add_argument("--groups", type=int, default=int(os.environ.get("SVS_BENCH_GROUPS", "16")),
                    help="independent contexts (one CUDA stream + one host thread each) the streams are split over")
    ap.add_argument("--h2d-mode", type=int, default=2, choices=[0, 2, 3],
                    help="e2e transfer of the pinned host frames: 2 zero-copy kernel reads over PCIe, 0 staged DMA copies")
    ap.add_argument("--clip-frames", type=int, default=48)
    ap.add_argument("--priming", type=int, default=150, help="untimed steps before warm-up so the BA window is full")
    ap.add_argument("--stagger", type=int, default=40,
                    help="spread the context groups' stream ages over this many steps (about one keyframe period); 0 = all in phase")
    ap.add_argument("--eager-right", action="store_true",
                    help="ingest the right image of every frame (default: only for the streams that insert a keyframe in the step)")
    ap.add_argument("--no-prefetch", action="store_true", help="disable the double-buffered ingest (svs_slam_hint_next)")
    ap.add_argument("--diag", action="store_true", help="repeat the value / e2e regions a second time (detail.diag)")
    ap.add_argument("--cpu-frames", type=int, default=0,
                    help="frames per stream per step of the CPU arms (0 = automatic: 40 for the in-run baseline, sized by K for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba4", action="store_true", help="skip the config-4 sharded-BA detail block")
    ap.add_argument("--sampler", default="nvml", choices=["nvml", "smi", "none"], help="clock / throttle-reason sampler")
    ap.add_argument("--no-kernel-pass", action="store_true", help="skip the instrumented per-kernel timing pass")
    ap.add_argument("--profile-window", action="store_true",
                    help="profiling aid: run only a device-resident pass with cudaProfilerStart/Stop around the timed steps")
    ap.add_argument("--cpu-baseline-clip", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_baseline_clip:
        cores = os.cpu_count() or 1
        n = args.cpu_frames if args.cpu_frames > 0 else 40
        cs = CpuStreams(args.cpu_baseline_clip, cores, CPU_PRIME)
        fps, frames, kfs = cs.step(n)
        cs.close()
        c1 = CpuStreams(args.cpu_baseline_clip, 1, CPU_PRIME)
        fps1, _, _ = c1.step(n)
        c1.close()
        print(json.dumps({"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "single_core_value": fps1,
                          "sample": "%d processes x %d frames of the same workload after %d priming frames each (BA window full: %d "
                                    "active keyframes); OpenCV stages through cv2 %s = the library the reference calls, 1 thread each; "
                                    "g2o blocks through oracle/geom.c; one core alone: %.1f frames/s"
                                    % (cores, n, CPU_PRIME, int(np.median(cs.window)), __import__("cv2").__version__, fps1)}))
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
