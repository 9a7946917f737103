#!/usr/bin/env python
"""bench.py — stereo frames/s of the StereoVision-SLAM hot path on B200 (BASELINE.json metric).

Headline workload (config.workload): BASELINE.json configs[1] — "KITTI seq-05 full pipeline (Frontend + Backend BA,
loop closure off), 1xB200": synthetic 1226x370 stereo pairs (no KITTI data is available offline) from the ray-cast
corridor generator with the seq-05 calibration, processed at the reference's half resolution 613x185,
num_features 150, BA window 10, synchronous BA schedule.  One STEP = one Frontend::AddFrame for every one of
`--streams` independent stereo streams (a batch of stereo pairs), i.e. `streams` frames.

  value  frames/s with the input images already resident in HBM
  e2e    frames/s through the public C ABI with the images in pinned HOST memory (H2D transfer of every frame
         inside the timed region; poses / statuses are read back to the host every step)

detail.config_1 / _3 / _4 / _5 hold the other BASELINE configs (frontend only; BA window 20; full resolution with
2000 features beside the sharded N = 50 BA; StereoBM(128, 15)), each with its own roofline and CPU baseline, and
detail.latency the single-stream ms/frame.

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
# One hardware work queue per CUDA stream (context groups x (main + ingest stream)).  Must be set before the CUDA context
# exists; it is the application's setting, the library does not touch the environment.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))

METRIC = "stereo_frames_per_sec"
UNIT = "frames/s"

# The BASELINE.json configs as pipeline settings.  "cfg" = svs_slam_config overrides.
CONFIGS = {
    2: dict(calib="kitti05", half=1, cfg=dict(backend_on=1), priming=150, stagger=40, cpu_prime=150,
            workload="KITTI seq-05-shaped full pipeline (Frontend GFTT+LK+triangulation+pose-LM, Backend BA window 10, "
                     "loop closure off), synthetic 1226x370 stereo pairs processed at 613x185"),
    1: dict(calib="kitti00", half=1, cfg=dict(backend_on=0), priming=60, stagger=40, cpu_prime=60,
            workload="config 1: KITTI seq-00-shaped FRONTEND ONLY (backend_on 0), synthetic 1241x376 pairs processed at 620x188"),
    3: dict(calib="kitti00", half=1, cfg=dict(backend_on=1, num_active_keyframes=20), priming=800, stagger=40, cpu_prime=800,
            workload="config 3: seq-00 geometry, full pipeline with num_active_keyframes 20 (override of default.yaml:27)"),
    4: dict(calib="kitti00", half=0, cfg=dict(backend_on=0, num_features=2000, gftt_min_distance=5.0, num_features_needed_for_keyframe=800,
                                              num_features_init=200, num_features_tracking=200, num_features_tracking_bad=80),
            priming=30, stagger=8, cpu_prime=12,
            workload="config 4 frontend: FULL resolution 1241x376 (no half-resolution resize), 2000 requested features with "
                     "minDistance 5 (deviation from src/frontend.cpp:24, SURVEY.md §0), frontend only; the N = 50 / L = 1e5 BA of "
                     "this config is detail.ba_config4"),
}


_REAL_STDOUT = None


def guard_stdout():
    """Libraries (NCCL's version banner) print to fd 1; the contract is ONE JSON line there.  Route fd 1 to stderr and keep
    the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def log(*a):
    if os.environ.get("SVS_BENCH_VERBOSE", "1") != "0":
        print("[bench %.1fs]" % (time.perf_counter() - _T0), *a, file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def pingpong(i, n):
    """Play a clip forward then backward (a physically valid camera motion) so any number of steps can be run."""
    p = i % (2 * n - 2)
    return p if p < n else 2 * n - 2 - p


_COR = {}


def _render_one(a):
    from svslam import synth
    calib, seed, n_frames, i, eye = a
    key = (calib, seed, n_frames)
    if key not in _COR:          # building the textured corridor costs seconds: once per worker process
        _COR[key] = synth.Corridor(calib, seed=seed, n_frames=n_frames)
    return _COR[key].render(i, eye)


def make_clip(calib, n_frames, seed=5):
    """Ray-cast clip, rendered by a process pool (a frame takes ~0.15 s on one core)."""
    from svslam import synth
    import multiprocessing as mp
    cor = synth.Corridor(calib, seed=seed, n_frames=n_frames)
    jobs = [(calib, seed, n_frames, i, eye) for eye in (0, 1) for i in range(n_frames)]
    procs = max(1, min(len(jobs), (os.cpu_count() or 1)))
    with mp.get_context("spawn").Pool(procs) as pool:
        imgs = pool.map(_render_one, jobs, chunksize=max(1, (len(jobs) + procs - 1) // procs))
    L, R = np.stack(imgs[:n_frames]), np.stack(imgs[n_frames:])
    T = np.stack([cor.T_cw(i) for i in range(n_frames)])
    return cor, L, R, T


def variant(img, v):
    """Photometric variant v of a clip (v = 0: the clip itself): gain, offset and fresh per-pixel sensor noise, the same for
    both eyes' exposure but with independent noise.  Every stream of the bench reads its own (variant, phase) pair, so the
    frames of a step are distinct BYTES (no L2 / PCIe reuse between streams) and distinct trajectories."""
    if v == 0:
        return img
    rng = np.random.RandomState(1000 + v)
    gain, off = 1.0 + 0.12 * (rng.rand() - 0.5), 16.0 * (rng.rand() - 0.5)
    noise = rng.randint(-3, 4, img.shape).astype(np.int16)
    return np.clip(np.rint(img.astype(np.float32) * gain + off).astype(np.int16) + noise, 0, 255).astype(np.uint8)


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
def _cpu_stream_proc(idx, clip_path, spec, conn):
    """One independent stream on one host core: the reference's CPU path (persistent: the pipeline is primed once so
    that its BA window is full, like the GPU arm's streams, then it runs `n` more frames per request)."""
    import cv2
    cv2.setNumThreads(1)
    from oracle import pipeline as op
    from svslam import synth
    d = np.load(clip_path)
    L, R = variant(d["L"], idx % 24), variant(d["R"], (idx % 24) + 100 if idx % 24 else 0)
    cal = synth.CALIB[spec["calib"]]
    s = 0.5 if spec["half"] else 1.0
    K = np.array([cal[2] * s, cal[2] * s, cal[3] * s, cal[4] * s])
    p = op.Pipeline(K, cal[5], op.Cfg(**spec["cfg"]), stages="cv2", cv2=cv2, half=bool(spec["half"]))
    nclip = len(L)
    cur = (7 * idx) % (2 * nclip - 2)
    for _ in range(spec["prime"]):
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

    def __init__(self, clip_path, n_procs, spec):
        import multiprocessing as mp
        from oracle import geom
        geom.build()
        ctx = mp.get_context("spawn")
        self.procs, self.conns = [], []
        for i in range(n_procs):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_cpu_stream_proc, args=(i, clip_path, spec, b), daemon=True)
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


def save_clip(L, R, tag=""):
    path = "/tmp/svslam_bench_clip_%d%s.npz" % (os.getpid(), tag)
    np.savez(path, L=L, R=R)
    return path


def cpu_spec(cid):
    c = CONFIGS[cid]
    return dict(calib=c["calib"], half=c["half"], cfg=c["cfg"], prime=int(os.environ.get("SVS_CPU_PRIME", c["cpu_prime"])))


def cpu_sample_text(cores, n, steps, spec, window):
    return ("%d processes x %d frames x %d step(s) after %d priming frames each (%d active keyframes); OpenCV stages through cv2 %s "
            "(the library the reference calls, 1 thread per process), g2o blocks through oracle/geom.c (g2o is not installable here)"
            % (cores, n, steps, spec["prime"], window, __import__("cv2").__version__))


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cid = 2
    spec = cpu_spec(cid)
    cor, L, R, T = make_clip(spec["calib"], args.clip_frames)
    path = save_clip(L, R)
    # each step = every core runs `n` frames of its own (primed, steady-state) stream; n is sized so that the whole
    # warm-up + K steps run stays within a few minutes whatever K is
    n = args.cpu_frames if args.cpu_frames > 0 else max(2, min(40, 2400 // (args.warmup + args.steps)))
    cs = CpuStreams(path, cores, spec)
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
        "config": {"workload": CONFIGS[cid]["workload"], "cpu_frames_per_stream_per_step": n, "processes": cores, "priming_frames": spec["prime"],
                   "active_keyframes_after_priming": int(np.median(cs.window))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(cores, n, args.steps, spec, int(np.median(cs.window)))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def cpu_baseline_subprocess(clip_path, cid, frames, timeout=300):
    """CPU baseline of one config in a clean subprocess (no CUDA / OpenMP state inherited), bounded by its own timeout."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-clip", clip_path, "--cpu-config", str(cid),
                            "--cpu-frames", str(frames)], capture_output=True, text=True, timeout=timeout)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:      # the GPU line is still valid without it
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "cpu baseline failed: %r" % (e,)}


# ------------------------------------------------------------------------------------------------ GPU arm
def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            v = json.load(open(peaks_path)).get("hbm_gbs")
            if isinstance(v, (int, float)) and v > 100:
                peak, src = float(v), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return peak, src


def ncu_traffic():
    """Mean DRAM bytes per launch per kernel from the committed `ncu --set full` capture of the bench's own configuration."""
    out = {}
    try:
        import csv
        for row in csv.DictReader(l for l in open(os.path.join(ROOT, "profiles", "r02_ncu_full_summary.csv")) if not l.startswith("#")):
            k = row["kernel"].replace("void ", "").split("<")[0].split("(")[0].replace("_tma", "")
            if k in out:
                continue
            out[k] = dict(traffic=(float(row["dram_rd_MB"]) + float(row["dram_wr_MB"])) * 1e6, dur_us=float(row["dur_us"]),
                          dram_pct=float(row["dram_pct"]), issue_active_pct=float(row["issue_active_pct"]),
                          fp64_pipe_pct=float(row["fp64_pipe_pct"]), registers=float(row["regs"]), grid=row.get("grid"))
    except Exception:
        pass
    return out


def ncu_shares():
    """Share of the serialised step per kernel class from the committed ncu launch list of this configuration
    (profiles/r02_launches_summary.csv).  Event-bracketed durations cannot rank the kernels: the ingest-stream kernels run
    underneath the compute kernels of the other context group and their brackets include the time they wait for SMs."""
    out = {}
    try:
        import csv
        for row in csv.DictReader(l for l in open(os.path.join(ROOT, "profiles", "r02_launches_summary.csv")) if not l.startswith("#")):
            k = row["kernel"].replace("void ", "").split("<")[0].split("(")[0].replace("_tma", "")
            out[k] = out.get(k, 0.0) + float(row["share"])
    except Exception:
        pass
    return out


LIMITER = {
    "k_ba_window": "dependent FP64 + L2 latency, one CTA per window",
    "k_pose_only_lm": "dependent FP64 latency, one warp per problem",
    "k_lk_track": "integer instruction issue (warp per keypoint, smem-staged patches)",
    "k_pyr_down": "HBM / issue (TMA-staged tiles, packed 16-bit arithmetic)",
    "k_corner_response": "latency (sequential f64 column march forced by bit-exactness)",
    "k_half_nearest": "HBM", "k_corner_select": "HBM", "k_bm_sad": "INT32 ALU + shared memory", "k_bm_prefilter": "HBM",
}


def kernel_rooflines(kern, cnt, P, n_win, peak, feats_per_kf):
    """ALGORITHMIC bytes moved by each kernel class over the region (DESIGN.md §4 / SURVEY.md §8d per-unit figures x the units
    the region processed) / the summed event-bracketed launch durations of that class."""
    frames = cnt["frames"]
    alg = {
        "k_half_nearest": 3.0 * P * (frames + cnt["right_images"]),
        "k_pyr_down": 1.64 * P * (frames + cnt["right_images"]),
        "k_corner_response": 5.0 * P * cnt["keyframes"],
        "k_corner_select": 5.0 * P * cnt["keyframes"],
        "k_lk_track": 3700.0 * cnt["lk_points"],
        "k_pose_only_lm": 40.0 * cnt["pose_edges"] * 56.0,            # ~56 LM trials per problem (4 rounds x 10 it + retries)
        "k_ba_window": (316.0 * cnt["ba_edges"] + 216.0 * cnt["ba_lms"] + 576.0 * n_win * cnt["ba_kfs"]) *
                       (cnt["ba_trials"] / max(1, cnt["ba_problems"])),
        "k_triangulate": 41.0 * cnt["keyframes"] * feats_per_kf,
    }
    out = {}
    for k, (ms_k, n_k) in kern.items():
        if k in alg and ms_k > 0 and n_k > 0:
            a = alg[k] / (ms_k * 1e-3) / 1e9
            out[k] = {"achieved_gbs": a, "frac": a / peak, "launches": n_k, "avg_launch_ms": ms_k / n_k,
                      "algorithmic_bytes_per_launch": alg[k] / n_k}
    return out


class Rig:
    """`streams` independent pipelines over `groups` contexts, fed from photometric variants of one clip that live once in
    HBM (value) and once in pinned host memory (e2e)."""

    def __init__(self, svslam, torch, dev, rank, clip, spec, streams, groups, host_threads, variants, args):
        self.svslam, self.torch, self.dev, self.args, self.spec = svslam, torch, dev, args, spec
        cor, L, R, T = clip
        self.cor, self.nclip = cor, len(L)
        self.B, self.G = streams, max(1, min(groups, streams))
        V = max(1, min(variants, streams))
        self.V = V
        self.img_bytes = cor.W * cor.H
        Lv = np.stack([variant(L, v) for v in range(V)]); Rv = np.stack([variant(R, v + 100 if v else 0) for v in range(V)])
        self.Ld = torch.from_numpy(Lv).cuda(dev); self.Rd = torch.from_numpy(Rv).cuda(dev)
        self.Lh = torch.from_numpy(Lv).pin_memory(); self.Rh = torch.from_numpy(Rv).pin_memory()
        del Lv, Rv
        self.period = 2 * self.nclip - 2
        # stream b plays variant b % V starting at its own phase: V x period distinct (variant, phase) pairs
        self.var = [b % V for b in range(streams)]
        self.start = [((b // V) * 7 + (b % V) * 3 + rank * 11) % self.period for b in range(streams)]
        self.distinct = len(set(zip(self.var, self.start)))
        self.ctxs = [svslam.Context(dev) for _ in range(self.G)]   # raises if libsvslam.so / a B200 is missing: no fallback
        self.lib = self.ctxs[0].lib
        self.lib.svs_kernel_name.restype = C.c_char_p
        for c in self.ctxs:
            c.set_wait_mode(1 if getattr(args, "wait_block", False) else 0)
        self.gsz = [streams // self.G + (1 if g < streams % self.G else 0) for g in range(self.G)]
        self.goff = np.concatenate([[0], np.cumsum(self.gsz)]).astype(int)
        K = cor.K_half() if spec["half"] else cor.K_full()
        kw = dict(spec["cfg"])
        kw["lazy_right_ingest"] = 0 if args.eager_right else 1
        kw["device_tracking"] = 0 if args.host_tracking else 1
        self.slams = [self.ctxs[g].slam(self.gsz[g], cor.W, cor.H, K, cor.baseline, half=bool(spec["half"]), **kw) for g in range(self.G)]
        for s in self.slams:
            s.set_threads(host_threads)
        self.cursor = 0
        self.gstag = [(g * spec["stagger"]) // self.G for g in range(self.G)]
        self.ptr_cache = {}
        for on_dev in (True, False):        # ctypes pointer arrays per (memory, group, clip phase), built once
            for g in range(self.G):
                for ph in range(self.period):
                    self.ptr_arrays(on_dev, ph, g)

    def ptr_arrays(self, on_device, step, g):
        key = (bool(on_device), g, step % self.period)
        v = self.ptr_cache.get(key)
        if v is None:
            bl, br = (self.Ld.data_ptr(), self.Rd.data_ptr()) if on_device else (self.Lh.data_ptr(), self.Rh.data_ptr())
            lp, rp = [], []
            for b in range(self.goff[g], self.goff[g + 1]):
                j = self.var[b] * self.nclip + pingpong(self.start[b] + step, self.nclip)
                lp.append(bl + j * self.img_bytes); rp.append(br + j * self.img_bytes)
            v = self.ptr_cache[key] = (self.svslam.Slam.ptr_array(lp), self.svslam.Slam.ptr_array(rp))
        return v

    def run_steps(self, n, on_device, stagger=False, serial=False):
        args = self.args

        def mode_of(g):      # e2e ingest: 2 = zero-copy kernel reads of the pinned host frames, 0 = staged strided DMA copies
            if on_device:
                return 1
            return args.h2d_mode if args.h2d_mode != 3 else (2 if g % 2 == 0 else 0)
        errors = []
        lo = self.cursor

        def loop(g):
            try:
                first = lo + (0 if stagger else self.gstag[g])
                last = lo + self.gstag[g] + n
                nxt = self.ptr_arrays(on_device, first, g)
                mode = mode_of(g)
                for s in range(first, last):
                    lp, rp = nxt
                    nxt = self.ptr_arrays(on_device, s + 1, g)
                    if args.no_prefetch:
                        self.slams[g].add_frames_arrays(lp, rp, mode)
                    else:   # double-buffered ingest: frame s+1 crosses PCIe / is resized while frame s is tracked
                        self.slams[g].add_frames_arrays(lp, rp, mode, nxt[0], nxt[1])
            except Exception as e:      # surface worker failures in the main thread
                errors.append(e)

        if self.G == 1:
            loop(0)
        elif serial:        # one group after another: a launch's event bracket is then the kernel's own duration
            for g in range(self.G):
                loop(g)
        else:
            th = [threading.Thread(target=loop, args=(g,)) for g in range(self.G)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        self.cursor += n
        if errors:
            raise errors[0]

    def ba_host(self):
        tot = np.zeros(3)
        for c in self.ctxs:
            o = np.zeros(3)
            self.lib.svs_ba_host_seconds(C.c_void_p(c.h), o.ctypes.data_as(C.c_void_p))
            tot += o
        return tot

    def region(self, steps, warmup, on_device, timing, barrier, dist=None, profile_window=False, serial=False):
        torch, lib = self.torch, self.lib
        self.run_steps(warmup, on_device, serial=serial)
        # steady state must not reallocate: a scratch-buffer regrowth is a device-wide synchronisation (a single one inside a
        # 60-step region costs tens of milliseconds on every context).  Sizes fluctuate with the number of keyframe streams
        # per step, so after priming + warm-up every buffer is grown once to 2 x the largest size it has been asked for.
        for c in self.ctxs:
            c.reserve_headroom(2.0)
        for c in self.ctxs:
            lib.svs_kernel_timing_reset(C.c_void_p(c.h))
            lib.svs_kernel_timing_enable(C.c_void_p(c.h), 1 if timing else 0)
        cn0 = [s.counters() for s in self.slams]
        bh0 = self.ba_host()
        l0 = sum(c.launch_count() for c in self.ctxs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile_window:     # `ncu --profile-from-start off`: only the steady-state steps below are captured
            torch.cuda.profiler.start()
        lib.svs_buffer_regrowths.restype = C.c_longlong
        rg0 = lib.svs_buffer_regrowths()
        t0 = time.perf_counter()
        cpu0 = time.process_time()
        e0.record()
        self.run_steps(steps, on_device, serial=serial)
        torch.cuda.synchronize(self.dev)
        e1.record()
        cpu_s = time.process_time() - cpu0      # user + system time of every thread of this rank over the timed steps
        if profile_window:
            torch.cuda.profiler.stop()
        e1.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = e0.elapsed_time(e1)
        launches = sum(c.launch_count() for c in self.ctxs) - l0
        cn1 = [s.counters() for s in self.slams]
        lost = int(sum((s.status == 3).sum() for s in self.slams))
        kern = {}
        for c in self.ctxs:
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
        bh1 = self.ba_host()
        phases.update({"ba:host_build": bh1[0] - bh0[0], "ba:pack_enqueue": bh1[1] - bh0[1], "ba:device_wait_unpack": bh1[2] - bh0[2]})
        counts = {k: sum(c1[1][k] - c0[1][k] for c0, c1 in zip(cn0, cn1)) for k in cn1[0][1]}
        log("region (on_device=%s, instrumented=%s): %.2f ms/step" % (on_device, timing, ms / steps))
        regrowths = int(lib.svs_buffer_regrowths() - rg0)
        if dist is not None:      # every rank must take the same re-measure decision
            t = torch.tensor([regrowths], device="cuda", dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            regrowths = int(t.item())
        return dict(ms=ms, wall=wall, launches=launches, phases=phases, counts=counts, kern=kern, lost=lost, cpu_s=cpu_s,
                    regrowths=regrowths, steps=steps)

    def region_clean(self, *a, **k):
        """region(), re-measured once if a scratch buffer had to regrow inside the timed steps (the regrowth's device-wide
        synchronisation is an artefact of sizes not seen during warm-up, not steady-state behaviour)."""
        r = self.region(*a, **k)
        if r["regrowths"] > 0:
            log("%d buffer regrowth(s) inside the timed region: re-measuring once" % r["regrowths"])
            first = r["ms"]
            r = self.region(*a, **k)
            r["remeasured_after_regrowth_ms"] = first
        return r

    def close(self):
        for s in self.slams:
            s.close()
        for c in self.ctxs:
            c.close()
        del self.Ld, self.Rd, self.Lh, self.Rh
        self.torch.cuda.empty_cache()


def side_config(svslam, torch, dev, rank, cid, clip, clip_path, streams, steps, warmup, args, cores, peak, peak_src, traffic):
    """One of the non-headline BASELINE configs: value + e2e + roofline of its dominant kernel + CPU baseline."""
    spec = CONFIGS[cid]
    cor = clip[0]
    t_start = time.perf_counter()
    rig = Rig(svslam, torch, dev, rank, clip, spec, streams, 1, max(1, cores), min(16, streams), args)
    barrier = lambda: torch.cuda.synchronize(dev)
    rig.run_steps(spec["priming"], True, stagger=True)
    dev_pass = rig.region_clean(steps, warmup, True, False, barrier)
    e2e_pass = rig.region_clean(steps, warmup, False, False, barrier)
    kern_pass = rig.region(max(4, steps // 2), 1, True, True, barrier)
    P = (int(round(cor.W * 0.5)) * int(round(cor.H * 0.5))) if spec["half"] else cor.W * cor.H
    cnt = kern_pass["counts"]
    feats = spec["cfg"].get("num_features", 150)
    kr = kernel_rooflines(kern_pass["kern"], cnt, P, spec["cfg"].get("num_active_keyframes", 10), peak, feats)
    kern = kern_pass["kern"]
    dom = max((k for k in kern if k in kr), key=lambda k: kern[k][0], default=None)
    roof = None
    if dom:
        roof = {"bound": "hbm", "kernel": dom, "achieved": kr[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": kr[dom]["frac"],
                "traffic": None, "avg_launch_ms": kr[dom]["avg_launch_ms"], "launches": kr[dom]["launches"],
                "algorithmic_bytes_per_launch": kr[dom]["algorithmic_bytes_per_launch"], "peak_source": peak_src,
                "limiter": LIMITER.get(dom, "HBM streaming")}
    tot = sum(v[0] for v in kern.values())
    rig.close()
    frames = streams * steps
    out = {"workload": spec["workload"], "streams": streams, "steps": steps, "warmup": warmup, "priming_steps": spec["priming"],
           "value": frames / (dev_pass["ms"] * 1e-3), "unit": UNIT, "ms_per_step": dev_pass["ms"] / steps,
           "e2e": {"value": frames / (e2e_pass["ms"] * 1e-3), "unit": UNIT, "ms_per_step": e2e_pass["ms"] / steps,
                   "h2d_bytes_per_step": int(e2e_pass["counts"]["h2d_image_bytes"] // steps), "d2h_bytes_per_step": streams * (56 + 12)},
           "gpu_launches": int(dev_pass["launches"]), "lost_streams": dev_pass["lost"], "roofline": roof,
           "kernel_time_share": {k: round(v[0] / tot, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])} if tot else {},
           "kernel_roofline": kr, "counts": dev_pass["counts"],
           "phase_seconds": {k: round(v, 4) for k, v in dev_pass["phases"].items()}, "distinct_sequences": rig.distinct}
    if not args.no_cpu_baseline and rank == 0:
        out["cpu_baseline"] = cpu_baseline_subprocess(clip_path, cid, args.cpu_frames or (8 if cid == 4 else 30))
    out["seconds_spent"] = round(time.perf_counter() - t_start, 1)
    return out


def stereo_bm_config(svslam, torch, dev, args, peak, peak_src):
    """BASELINE config 5: cv::StereoBM(128, 15) (src/dense_reconstruction.cpp:89,114) on batches of rectified pairs."""
    import cv2
    res = {}
    ctx = svslam.Context(dev)
    lib = ctx.lib
    lib.svs_kernel_name.restype = C.c_char_p
    for (w, h, n) in ((620, 188, 512), (1241, 376, 128)):
        rng = np.random.RandomState(50)
        # rectified pairs with a disparity ramp 2..120 px: R(x) = L(x + d(x)), textured like the survey probes
        base = cv2.GaussianBlur(rng.randint(0, 256, (h, w + 130)).astype(np.uint8), (0, 0), 2)
        base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
        for _ in range(40 * (w * h) // (620 * 188)):
            x, y = rng.randint(0, w + 100), rng.randint(0, h - 10)
            cv2.rectangle(base, (x, y), (x + rng.randint(8, 60), y + rng.randint(6, 40)), int(rng.randint(0, 256)), -1)
        disp = np.linspace(2.0, 120.0, w, dtype=np.float32)[None, :] * np.ones((h, 1), np.float32)
        xs = np.arange(w, dtype=np.float32)[None, :] + disp
        ys = np.arange(h, dtype=np.float32)[:, None] * np.ones((1, w), np.float32)
        l0, r0 = base[:, :w].copy(), cv2.remap(base, xs, ys, cv2.INTER_LINEAR)
        # n distinct pairs: per-pair noise so that no two images share bytes
        Ls = np.stack([np.clip(l0.astype(np.int16) + np.random.RandomState(i).randint(-2, 3, l0.shape), 0, 255).astype(np.uint8) for i in range(n)])
        Rs = np.stack([np.clip(r0.astype(np.int16) + np.random.RandomState(1000 + i).randint(-2, 3, r0.shape), 0, 255).astype(np.uint8) for i in range(n)])
        Ld, Rd = torch.from_numpy(Ls).cuda(dev), torch.from_numpy(Rs).cuda(dev)
        Dd = torch.empty((n, h, w), dtype=torch.int16, device=dev)
        Lh, Rh = torch.from_numpy(Ls).pin_memory(), torch.from_numpy(Rs).pin_memory()
        Dh = torch.empty((n, h, w), dtype=torch.int16).pin_memory()

        def dev_call():
            ctx._chk(lib.svs_stereo_bm_dev(C.c_void_p(ctx.h), C.c_void_p(Ld.data_ptr()), C.c_void_p(Rd.data_ptr()), w, h, w, n,
                                           C.c_size_t(w * h), 128, 15, C.c_void_p(Dd.data_ptr())))

        def host_call():
            ctx._chk(lib.svs_stereo_bm(C.c_void_p(ctx.h), C.c_void_p(Lh.data_ptr()), C.c_void_p(Rh.data_ptr()), w, h, w, n,
                                       C.c_size_t(w * h), 128, 15, C.c_void_p(Dh.data_ptr())))
        reps = 5
        for _ in range(3):
            dev_call()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            dev_call()
        torch.cuda.synchronize(dev)
        t_dev = (time.perf_counter() - t0) / reps
        for _ in range(2):
            host_call()
        t0 = time.perf_counter()
        for _ in range(reps):
            host_call()
        t_e2e = (time.perf_counter() - t0) / reps
        # per-kernel durations (CUDA events on the context stream)
        lib.svs_kernel_timing_reset(C.c_void_p(ctx.h)); lib.svs_kernel_timing_enable(C.c_void_p(ctx.h), 1)
        for _ in range(reps):
            dev_call()
        nk = lib.svs_kernel_timing_get(C.c_void_p(ctx.h), None, None, 0)
        kms = np.zeros(nk); kcnt = np.zeros(nk, np.int64)
        lib.svs_kernel_timing_get(C.c_void_p(ctx.h), kms.ctypes.data_as(C.c_void_p), kcnt.ctypes.data_as(C.c_void_p), nk)
        lib.svs_kernel_timing_enable(C.c_void_p(ctx.h), 0)
        kern = {lib.svs_kernel_name(i).decode(): (float(kms[i]), int(kcnt[i])) for i in range(nk) if kcnt[i]}
        sad_ms = kern.get("k_bm_sad", (0.0, 1))[0] / max(1, kern.get("k_bm_sad", (0.0, 1))[1])
        P = w * h
        valid = max(0, (w - 7 - 134)) * max(0, (h - 14))                # columns [134, W-7), rows [7, H-7)
        got = Dh.numpy()[0]
        want = cv2.StereoBM_create(128, 15).compute(Ls[0], Rs[0])
        assert np.array_equal(got, want), "StereoBM parity lost in the bench"
        # CPU: cv2.StereoBM with all its threads on a bounded sample of the same pairs
        cv2.setNumThreads(os.cpu_count() or 1)
        bm = cv2.StereoBM_create(128, 15)
        m = min(n, 24 if w < 1000 else 8)
        bm.compute(Ls[0], Rs[0])
        t0 = time.perf_counter()
        for i in range(m):
            bm.compute(Ls[i], Rs[i])
        t_cpu = (time.perf_counter() - t0) / m
        ach = 4.0 * P * n / (sad_ms * 1e-3) / 1e9 if sad_ms > 0 else None
        res["%dx%d" % (w, h)] = {
            "pairs_per_call": n, "value": n / t_dev, "unit": "stereo pairs/s", "ms_per_call": 1e3 * t_dev,
            "e2e": {"value": n / t_e2e, "unit": "stereo pairs/s", "h2d_bytes_per_step": 2 * P * n, "d2h_bytes_per_step": 2 * P * n,
                    "ms_per_call": 1e3 * t_e2e},
            "roofline": {"bound": "hbm", "kernel": "k_bm_sad", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if ach else None,
                         "traffic": None, "avg_launch_ms": sad_ms, "algorithmic_bytes_per_launch": 4.0 * P * n, "peak_source": peak_src,
                         "limiter": LIMITER["k_bm_sad"],
                         "int_ops_per_sec": valid * 128 * 6.0 * n / (sad_ms * 1e-3) if sad_ms > 0 else None,
                         "int_ops_note": "valid pixels x 128 disparities x ~6 integer ops per cost cell (SURVEY.md §8d)"},
            "kernel_ms_per_call": {k: round(v[0] / reps, 4) for k, v in kern.items()},
            "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "stereo pairs/s", "cores": os.cpu_count(), "kind": "reference",
                             "sample": "cv2 %s StereoBM_create(128,15).compute on %d of the same pairs, cv2.setNumThreads(%d) (the library "
                                       "routine the reference calls at src/dense_reconstruction.cpp:114)" % (cv2.__version__, m, os.cpu_count() or 1)},
            "bit_exact_vs_cv2": True}
        del Ld, Rd, Dd, Lh, Rh, Dh
    ctx.close()
    return {"workload": "config 5: cv::StereoBM(128,15) on batches of distinct rectified pairs (disparities 2-120 px)", **res}


def latency_block(svslam, dev, clip, n_frames=90):
    """Single-stream latency (the reference is a one-camera real-time system, src/visual_odometry.cpp:126-153 times one frame):
    one stream, one frame per call through svs_slam_add_frames with host frames."""
    cor, L, R, T = clip
    ctx = svslam.Context(dev)
    one = ctx.slam(1, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, backend_on=1)
    n = len(L)
    ts, kf = [], []
    for i in range(n_frames):
        j = pingpong(i, n)
        t0 = time.perf_counter()
        one.add_frames(L[j:j + 1], R[j:j + 1])
        ts.append(time.perf_counter() - t0); kf.append(int(one.is_kf[0]))
    ph, cn = one.counters()
    lost = int(one.status[0] == 3)
    one.close(); ctx.close()
    ts, kf = np.array(ts[10:]), np.array(kf[10:], bool)
    return {"streams": 1, "frames": int(len(ts)), "ms_per_frame_mean": float(1e3 * ts.mean()), "ms_per_frame_median": float(1e3 * np.median(ts)),
            "ms_per_frame_p99": float(1e3 * np.quantile(ts, 0.99)),
            "ms_per_tracked_frame_median": float(1e3 * np.median(ts[~kf])) if (~kf).any() else None,
            "ms_per_keyframe_frame_median": float(1e3 * np.median(ts[kf])) if kf.any() else None, "keyframes": int(kf.sum()),
            "frames_per_sec": float(1.0 / ts.mean()), "lost": lost,
            "phase_ms_per_frame": {k: round(1e3 * v / n_frames, 4) for k, v in ph.items() if not k.startswith("host:")},
            "note": "pageable host frames through svs_slam_add_frames, one synchronous call per frame; compare with cpu_baseline.single_core_value"}


def run_ba_config4(ctx, dist, rank, world, dev):
    """BA LM iterations/s at BASELINE config-4 scale (N = 50 keyframes, L = 1e5 landmarks, ~5e5 edges), landmarks sharded over
    the ranks: one persistent cooperative solver kernel per GPU, the ranks' kernels exchange their partial
    reduced camera systems through peer memory (svs_ba_shard_optimize; no host round trip, no NCCL in the loop)."""
    import torch
    from svslam import ba_shard
    from svslam.problems import K05, EXT_L, EXT_R, ba_problem_big
    prob = ba_problem_big(4, n_kf=50, n_lm=100000)
    p, _ = ba_shard.split_problem(prob, world)[rank]
    best = None
    for rep in range(4):
        sh = ba_shard.Shard(ctx, p["poses"], p["lms"], p["edge_kf"], p["edge_lm"], p["edge_cam"], p["edge_uv"], K05, K05, EXT_L, EXT_R)
        if dist is not None:
            ba_shard.wire_distributed(sh, dist)
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st_ = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
        e0.record(st_)
        sh.launch(10)
        e1.record(st_)
        st = sh.finish()
        e1.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            dist.barrier()          # nobody unmaps a window a peer may still read
        sh.close()
        if rep > 0 and (best is None or dt < best[0]):      # the first repetition is the warm-up
            best = (dt, st)
    dt, st = best
    E, L, N = int(len(prob["edge_kf"])), 100000, 50
    alg = (316.0 * E + 216.0 * L + 576.0 * N * N) * st["trials"]
    peak, peak_src = hbm_peak()
    xbytes = (N * (N + 1) // 2 * 36 + 12 * N + 8) * 8
    return {"n_kf": N, "n_lm": L, "n_edges": E, "shards": world, "lm_iterations": st["iterations"],
            "trials": st["trials"], "seconds": dt, "lm_iterations_per_sec": st["iterations"] / dt, "chi2_init": st["chi2_init"],
            "chi2": st["chi2"], "kernel": "k_bs_lm (one persistent cooperative launch per optimize(10))",
            "roofline": {"bound": "hbm", "kernel": "k_bs_lm", "achieved": alg / dt / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / dt / 1e9 / peak, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                         "note": "316 E + 216 L + 576 N^2 bytes per LM trial (SURVEY.md §8d) x trials / the kernel's duration (CUDA events "
                                 "on the context stream); the solve of the 300 x 300 reduced system and ~25 grid barriers per trial are "
                                 "latency, not traffic"},
            "exchange": "none (one shard)" if world == 1 else "in-kernel all-gather + ordered local reduce through peer windows (NVLink)",
            "exchange_bytes_per_trial_per_rank": 0 if world == 1 else (world - 1) * (xbytes + 16)}


def build_report(args, world, B, G, cor, spec, dev_pass, e2e_pass, kern_pass, clocks, peak, peak_src, traffic, cpu, ate, ba4, lat,
                 detail_cfg, distinct, V, cores, my_cores, host_threads, nclip):
    """The JSON line of the GPU arm from the measured regions (pure function: tests/test_bench_cpu.py runs it on fabricated
    measurements)."""
    frames_total = B * args.steps * world
    value = frames_total / (dev_pass["ms"] * 1e-3)
    e2e = frames_total / (e2e_pass["ms"] * 1e-3)
    # ---- roofline of the dominant kernel (largest share of device time in the instrumented region)
    kern = kern_pass["kern"]
    cnt = kern_pass["counts"]
    P = int(round(cor.W * 0.5)) * int(round(cor.H * 0.5))
    kr = kernel_rooflines(kern, cnt, P, 10, peak, 150)
    shares_ncu = ncu_shares()
    ingest = ("k_half_nearest", "k_pyr_down")     # launched on the ingest stream, overlapped with the other kernels
    if any(k in shares_ncu for k in kr):
        dom = max((k for k in kr), key=lambda k: shares_ncu.get(k, 0.0), default=None)
        dom_src = ("largest share of the serialised step in profiles/r02_launches_summary.csv (%.0f %%): the ncu launch list of this "
                   "workload at 2 context groups, where every launch fills the machine and serialised time = SM time; at the bench's "
                   "%d groups the small BA launches (<= 11 CTAs, ~3 ms) are the largest SERIALISED share (kernel_time_share, "
                   "profiles/r02_launches_16groups_summary.csv) but run underneath the other groups' kernels"
                   % (100 * shares_ncu.get(dom, 0.0), G))
    else:
        dom = max((k for k in kern if k in kr and k not in ingest), key=lambda k: kern[k][0], default=None)
        dom_src = "largest summed event-bracketed time among the compute-stream kernels"
    roof = None
    if dom:
        tr = traffic.get(dom)
        # the ncu capture launched this kernel for 2048 streams at a time (2 context groups), the live run for B / G streams:
        # DRAM bytes per launch are scaled by the units per launch (keypoints = 4 per CTA for LK, windows = CTAs for BA)
        tr_bytes, tr_note = (tr["traffic"], "unscaled") if tr else (None, None)
        per_cta = {"k_lk_track": 4 * 3700.0}.get(dom)
        if tr and per_cta and tr.get("grid"):
            try:
                ncu_alg = float(tr["grid"]) * per_cta
                tr_bytes = tr["traffic"] * kr[dom]["algorithmic_bytes_per_launch"] / ncu_alg
                tr_note = "ncu bytes per launch x (live algorithmic bytes per launch / algorithmic bytes of the captured launch = %.3f)" % (
                    kr[dom]["algorithmic_bytes_per_launch"] / ncu_alg)
            except (ValueError, ZeroDivisionError):
                pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": kr[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": kr[dom]["frac"],
                "traffic": tr_bytes,
                "traffic_source": ("profiles/r02_ncu_full_summary.csv (ncu --set full, mean DRAM read + write bytes per launch; %s)" % tr_note) if tr else None,
                "avg_launch_ms": kr[dom]["avg_launch_ms"], "launches": kr[dom]["launches"],
                "algorithmic_bytes_per_launch": kr[dom]["algorithmic_bytes_per_launch"], "peak_source": peak_src,
                "dominant_by": dom_src,
                "note": "launch durations from CUDA events on the launching stream, %s; limiter: %s (DESIGN.md §4)"
                        % ("the %d context groups stepped one after another in this instrumented pass (a bracket is the kernel's own "
                           "duration)" % G if kern_pass.get("serial_groups") else "%d context group(s) overlapping" % G,
                           LIMITER.get(dom, "HBM streaming"))}
    for k, v in kr.items():
        if k in traffic:
            v["ncu_standalone"] = traffic[k]
    dev_total = sum(v[0] for v in kern.values())
    shares = {k: round(v[0] / dev_total, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])} if dev_total else {}

    h2d = int(e2e_pass["counts"]["h2d_image_bytes"] // args.steps)
    rows = (cor.H + 1) // 2
    distinct_bytes = min(distinct, B) * cor.W * rows
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_pass["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["workload"], "streams_per_gpu": B, "context_groups": G, "frames_per_step": B * world, "num_features": 150,
                   "num_active_keyframes": 10, "ba": "synchronous, analytic Jacobians",
                   "tracking": "host round trip per seam" if args.host_tracking else
                               "device-resident per-stream state (one host sync per step; host classes only at keyframes)",
                   "ingest": ("per-step push" if args.no_prefetch else "double-buffered: frame t+1 is ingested on a second stream during step t") +
                             ("; both eyes of every frame" if args.eager_right else
                              "; right images are ingested lazily, only for the streams that insert a keyframe in the step "
                              "(the frontend reads the right image nowhere else; results are bit-identical)"),
                   "clip_frames": nclip, "photometric_variants": V, "distinct_sequences_per_gpu": distinct,
                   "priming_steps": spec["priming"] if args.priming < 0 else args.priming, "group_stagger_steps": spec["stagger"],
                   "l2": "every stream reads its own (variant, phase) frame: %d distinct left frames = %.0f MB of even rows per step per GPU "
                         "(%s the 126 MB L2), plus ~0.6 GB of pyramids written and re-read per step"
                         % (min(distinct, B), distinct_bytes / 1e6, "larger than" if distinct_bytes > 126e6 else "NOT larger than")},
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
                   "kernel_time_share": shares, "kernel_roofline": kr, "kernel_pass_ms_per_step": kern_pass["ms"] / kern_pass.get("steps", args.steps),
                   "kernel_pass": "context groups stepped one after another" if kern_pass.get("serial_groups") else "context groups overlapping",
                   "summed_kernel_ms_over_wall_ms": dev_total / kern_pass["ms"] if kern_pass["ms"] else None,
                   "lost_streams": dev_pass["lost"], "host_cores": cores, "host_cores_this_rank": my_cores, "host_threads_per_group": host_threads,
                   "buffer_regrowths_in_timed_regions": {"value": dev_pass.get("regrowths"), "e2e": e2e_pass.get("regrowths")},
                   "host_wait_mode": "block" if getattr(args, "wait_block", False) else "spin",
                   "host_cpu_ms_per_step": {"value": round(1e3 * dev_pass.get("cpu_s", 0.0) / args.steps, 2), "e2e": round(1e3 * e2e_pass.get("cpu_s", 0.0) / args.steps, 2),
                                            "note": "user + system time of every thread of rank 0 per timed step (spinning waits count as busy)"},
                   "ba_lm_iterations_per_sec": dev_pass["counts"]["ba_iterations"] * world / (dev_pass["ms"] * 1e-3),
                   "accuracy": ate, "ba_config4": ba4, "latency": lat, **detail_cfg},
    }
    return out


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
    my_cores = max(1, cores // max(1, local_world))
    # Context groups: each is one host thread that drives its share of the streams, so that one group's serial chain (track
    # kernels -> keyframe bookkeeping -> detect / right LK / triangulate -> BA -> bookkeeping) overlaps the other groups'
    # kernels.  Measured on one B200 (profiles/README.md): 16 groups beat fewer groups at every core count, also with fewer
    # cores than groups as long as the waiting threads sleep (4 cores: 184 k frames/s with 16 sleeping groups, 154 k with 4
    # spinning ones, 125 k with 2 groups x 2 OpenMP threads); with a core per group spinning waits win (16 cores: e2e 164 k
    # vs 153 k).
    gcap = {"cores": my_cores, "half": max(1, my_cores // 2), "none": 1 << 30}[args.group_cap]
    G = max(1, min(args.groups, args.streams, gcap))
    host_threads = max(1, my_cores // G)
    # host wait policy: a group thread spins on its stream while it has a core of its own (lowest wake-up latency); with more
    # group threads than cores they sleep on a blocking-sync event instead
    args.wait_block = (args.wait_mode == "block") or (args.wait_mode == "auto" and my_cores < G)
    spec = CONFIGS[2]
    log("rendering clip ...")
    clip = make_clip(spec["calib"], args.clip_frames)
    cor, L, R, T = clip
    log("clip ready")
    B = args.streams
    peak, peak_src = hbm_peak()
    traffic = ncu_traffic()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ONE set of pipelines, primed once (untimed) until every stream's sliding BA window is full, then timed in several
    # regions that only differ in where the frames live / whether launches are instrumented.  Every region does its own
    # W warm-up steps first; the stream state simply continues from region to region (steady state).
    rig = Rig(svslam, torch, dev, rank, clip, spec, B, G, host_threads, args.variants, args)
    log("rig ready: %d streams, %d context group(s), %d host threads each, %d distinct sequences" % (B, G, host_threads, rig.distinct))
    rig.run_steps(spec["priming"] if args.priming < 0 else args.priming, True, stagger=True)
    log("primed")
    if args.profile_window:   # profiling aid (never a bench number): one device-resident region inside a profiler window
        rig.region(args.steps, args.warmup, not args.profile_e2e, False, barrier, profile_window=True)
        log("profile window done")
        return
    rig.run_steps(args.warmup, False)     # the e2e path's own buffers (pointer tables, staging) exist before anything is timed
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(dev, args.sampler, uuid=gpu_uuid)
    sampler.start()
    dev_pass = rig.region_clean(args.steps, args.warmup, True, False, barrier, dist)
    e2e_pass = rig.region_clean(args.steps, args.warmup, False, False, barrier, dist)
    clocks = sampler.stop()
    # instrumented pass (feeds roofline / kernel_ms only): every launch bracketed by CUDA events on its stream, the context
    # groups stepped ONE AFTER ANOTHER — with the groups overlapping, a bracket would also contain the time the launch waits
    # for SMs held by the other groups' kernels
    if args.no_kernel_pass:
        kern_pass = dev_pass
    else:
        try:
            kern_pass = rig.region(max(4, args.steps // 2), 2, True, True, barrier, dist, serial=True)
            kern_pass["serial_groups"] = True
        except Exception as e:      # never lose the bench line to the instrumented pass
            log("serial instrumented pass failed (%r): concurrent pass instead" % (e,))
            kern_pass = rig.region(args.steps, args.warmup, True, True, barrier, dist)
    distinct = rig.distinct
    V = rig.V
    rig.close()

    ba4 = None
    ctx0 = svslam.Context(dev)
    if not args.no_ba4:
        ba4 = run_ba_config4(ctx0, dist, rank, world, dev)
    ctx0.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # accuracy beside the speed (BASELINE.json: "ATE vs reference"): one stream over the clip's forward pass, ATE against the
    # generator's ground truth (the oracle pipeline's ATE on the same frames is asserted equal within 3 cm in tests/)
    ate = None
    try:
        from svslam import kitti
        c1 = svslam.Context(dev)
        one = c1.slam(1, cor.W, cor.H, cor.K_half(), cor.baseline, half=True, backend_on=1)
        est = [one.add_frames(L[i:i + 1], R[i:i + 1])[0].copy() for i in range(len(L))]
        lost = int(one.status[0] == 3)
        one.close(); c1.close()
        ce, _ = kitti.pose7_to_Twc(np.array(est))
        cg, _ = kitti.pose7_to_Twc(np.asarray(T)[:len(L)])
        ate = {"ate_rmse_m": kitti.ate_rmse(ce, cg), "frames": len(L), "path_m": float(np.linalg.norm(np.diff(cg, axis=0), axis=1).sum()),
               "lost": lost}
    except Exception as e:
        ate = {"error": repr(e)}

    # ---- CPU baseline on this box (bounded sample)
    cpu = None
    clip_path = save_clip(L, R)
    if world > 1:      # the CPU baseline, the other configs and the latency block are N = 1 material (rank 0, one GPU)
        args.no_cpu_baseline, args.configs, args.no_latency = True, "", True
    if not args.no_cpu_baseline:
        log("cpu baseline subprocess ...")
        cpu = cpu_baseline_subprocess(clip_path, 2, args.cpu_frames or 40)

    # ---- the other BASELINE configs + single-stream latency (rank 0, one GPU)
    detail_cfg = {}
    want = [int(x) for x in args.configs.split(",") if x.strip()] if args.configs else []
    clip00 = clip00_path = None
    for cid in want:
        try:
            if cid == 5:
                log("config 5 (StereoBM) ...")
                detail_cfg["config_5"] = stereo_bm_config(svslam, torch, dev, args, peak, peak_src)
                continue
            if cid not in CONFIGS or cid == 2:
                continue
            if clip00 is None:
                log("rendering seq-00-shaped clip ...")
                clip00 = make_clip("kitti00", args.clip_frames, seed=3)
                clip00_path = save_clip(clip00[1], clip00[2], "_00")
            log("config %d ..." % cid)
            streams = {1: 2048, 3: 1024, 4: 96}[cid]
            steps = {1: 20, 3: 20, 4: 10}[cid]
            detail_cfg["config_%d" % cid] = side_config(svslam, torch, dev, rank, cid, clip00, clip00_path, streams, steps, 3, args, my_cores,
                                                        peak, peak_src, traffic)
        except Exception as e:
            detail_cfg["config_%d" % cid] = {"error": repr(e)}
    lat = None
    if not args.no_latency:
        try:
            lat = latency_block(svslam, dev, clip)
        except Exception as e:
            lat = {"error": repr(e)}
    for p_ in (clip_path, clip00_path):
        if p_ and os.path.exists(p_):
            os.remove(p_)

    out = build_report(args, world, B, G, cor, spec, dev_pass, e2e_pass, kern_pass, clocks, peak, peak_src, traffic, cpu, ate, ba4, lat,
                       detail_cfg, distinct, V, cores, my_cores, host_threads, len(L))
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("SVS_BENCH_STREAMS", "4096")))
    ap.add_argument("--groups", type=int, default=int(os.environ.get("SVS_BENCH_GROUPS", "16")),
                    help="independent contexts (one CUDA stream pair + one driver thread each) the streams are split over: one "
                         "group's host keyframe bookkeeping overlaps the other group's kernels")
    ap.add_argument("--group-cap", default=os.environ.get("SVS_BENCH_GROUP_CAP", "none"), choices=["half", "cores", "none"],
                    help="upper bound of the context groups of a rank: half = cores / 2 (a driver thread + one helper each), cores = one per core, none = as many as --groups asks for (more groups than cores: the threads sleep in their waits)")
    ap.add_argument("--variants", type=int, default=24, help="photometric variants of the clip (distinct frame bytes per stream)")
    ap.add_argument("--h2d-mode", type=int, default=2, choices=[0, 2, 3],
                    help="e2e transfer of the pinned host frames: 2 zero-copy kernel reads over PCIe, 0 staged DMA copies")
    ap.add_argument("--clip-frames", type=int, default=48)
    ap.add_argument("--priming", type=int, default=-1, help="untimed steps before warm-up so the BA window is full (-1: per config)")
    ap.add_argument("--eager-right", action="store_true",
                    help="ingest the right image of every frame (default: only for the streams that insert a keyframe in the step)")
    ap.add_argument("--host-tracking", action="store_true", help="device_tracking = 0: every seam of Track() is a host round trip (round-1 path)")
    ap.add_argument("--wait-mode", default="auto", choices=["auto", "spin", "block"],
                    help="host wait policy (svs_set_wait_mode): spin = cudaStreamSynchronize, block = blocking-sync event, auto = block when "
                         "this rank has fewer than 4 cores per context group")
    ap.add_argument("--no-prefetch", action="store_true", help="disable the double-buffered ingest (svs_slam_hint_next)")
    ap.add_argument("--cpu-frames", type=int, default=0,
                    help="frames per stream per step of the CPU arms (0 = automatic)")
    ap.add_argument("--configs", default="1,3,4,5", help="other BASELINE configs measured into detail.config_N (rank 0; '' = none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba4", action="store_true", help="skip the config-4 sharded-BA detail block")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-stream latency block")
    ap.add_argument("--sampler", default="nvml", choices=["nvml", "smi", "none"], help="clock / throttle-reason sampler")
    ap.add_argument("--no-kernel-pass", action="store_true", help="skip the instrumented per-kernel timing pass")
    ap.add_argument("--profile-window", action="store_true",
                    help="profiling aid: run only one pass with cudaProfilerStart/Stop around the timed steps")
    ap.add_argument("--profile-e2e", action="store_true", help="with --profile-window: frames in pinned host memory")
    ap.add_argument("--cpu-baseline-clip", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-config", type=int, default=2, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_baseline_clip:
        cores = os.cpu_count() or 1
        spec = cpu_spec(args.cpu_config)
        n = args.cpu_frames if args.cpu_frames > 0 else 40
        cs = CpuStreams(args.cpu_baseline_clip, cores, spec)
        fps, frames, kfs = cs.step(n)
        # one core alone: the first process repeats the sample while the others idle
        cs.conns[0].send(n)
        r1 = cs.conns[0].recv()
        fps1 = r1[1] / r1[0]
        cs.close()
        print(json.dumps({"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "single_core_value": fps1,
                          "sample": cpu_sample_text(cores, n, 1, spec, int(np.median(cs.window))) +
                                    "; one core alone: %.1f frames/s" % fps1}))
        return
    guard_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
