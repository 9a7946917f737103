"""bench.py pieces that run without a GPU: the reference arm (`--impl reference`) end to end on a tiny sample, and the
helpers the GPU arm relies on."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("svs_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_pingpong_is_a_valid_periodic_camera_path():
    b = _bench()
    n = 7
    seq = [b.pingpong(i, n) for i in range(40)]
    assert seq[:13] == [0, 1, 2, 3, 4, 5, 6, 5, 4, 3, 2, 1, 0]
    assert all(abs(x - y) == 1 for x, y in zip(seq, seq[1:]))            # consecutive frames are neighbours in the clip
    assert seq[:12] == seq[12:24]                                         # period 2n - 2


def test_clock_sampler_degrades_without_a_gpu():
    b = _bench()
    for mode in ("nvml", "smi", "none"):
        s = b.ClockSampler(0, mode, period=0.01)
        s.start()
        out = s.stop()
        assert "reasons" in out and "sm_mhz" in out                       # never raises; reports what it has


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, SVS_CPU_PRIME="12", SVS_BENCH_VERBOSE="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                        "--cpu-frames", "3", "--clip-frames", "16"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "stereo_frames_per_sec" and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["steps"] == 2 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # ranks other than 0 print nothing and exit 0 (torchrun launch of the reference arm)
    env2 = dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"],
                        capture_output=True, text=True, timeout=120, env=env2)
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_gpu_arm_report_block_dry_run():
    """The part of bench.py's GPU arm that turns the measured regions into the JSON line (roofline, traffic, per-kernel view,
    e2e bytes) executed on fabricated measurements: it must produce every key of the bench contract."""
    import textwrap
    import types
    src = open(os.path.join(ROOT, "bench.py")).read()
    a = src.index("    # ---- roofline of the dominant kernel")
    b = src.index("    print(json.dumps(out))\n    if dist is not None:")
    block = textwrap.dedent(src[a:b])
    names = ("push", "track_lk", "pose_lm", "detect", "right_lk", "triangulate", "ba", "host")
    cn = ("frames", "keyframes", "ba_problems", "ba_iterations", "ba_trials", "ba_edges", "ba_lms", "ba_kfs", "lk_points", "pose_edges",
          "h2d_image_bytes", "right_images")
    kernels = ["k_half_nearest", "k_pyr_down", "k_mask_boxes", "k_corner_response", "k_corner_select", "k_corner_greedy", "k_lk_track",
               "k_triangulate", "k_pose_only_lm", "k_ba_window"]
    p = dict(ms=1200.0, wall=1.2, launches=12000, phases={k: 1.0 for k in names}, counts={k: 1000 + i for i, k in enumerate(cn)},
             kern={k: (100.0 + i, 50 + i) for i, k in enumerate(kernels)}, lost=0)
    ns = {}
    exec("import os, sys, json, subprocess, time\nimport numpy as np\n", ns)
    ns.update(ROOT=ROOT, METRIC="stereo_frames_per_sec", UNIT="frames/s", WORKLOAD="w", C=None, log=print, save_clip=None, L=None, R=None,
              args=types.SimpleNamespace(no_cpu_baseline=True, steps=60, warmup=5, eager_right=False, no_prefetch=False, stagger=40,
                                         priming=150, h2d_mode=2, cpu_frames=0),
              kern_pass=p, dev_pass=dict(p, kern={}), e2e_pass=p, cor=types.SimpleNamespace(W=1226, H=370), B=4096, G=16, world=1,
              clocks={"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []}, ba4=None, ate=None, diag=None, nclip=48, cores=16,
              host_threads=1, value=2e5, e2e=1.5e5, img_bytes=1226 * 370)
    exec(block, ns)
    out = json.loads(json.dumps(ns["out"]))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in out, k
    assert out["config"]["workload"] == "w" and "model" not in out["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(out["e2e"])
    r = out["roofline"]
    assert r["kernel"] == "k_ba_window" and r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is not None and r["traffic"] > 0
    assert "k_lk_track" in out["detail"]["kernel_roofline"] and "ncu_standalone" in out["detail"]["kernel_roofline"]["k_lk_track"]
