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
