"""CPU-only tests of the C++ host mirrors (Frontend / Map bookkeeping containers, SE3 algebra, window eviction): built
with g++ straight from stereovision-slam_b200/host (no CUDA), plus a smoke run of the host bookkeeping micro-benchmark."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "stereovision-slam_b200", "host")


def _build(src, out):
    cmd = ["g++", "-O2", "-std=c++17", "-I" + HOST, "-I" + os.path.join(ROOT, "include"), src, os.path.join(HOST, "slam.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_host_unit(tmp_path):
    exe = str(tmp_path / "host_unit")
    _build(os.path.join(ROOT, "tests", "host_unit.cpp"), exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host unit tests ok" in r.stdout


def test_hostbench_smoke(tmp_path):
    exe = str(tmp_path / "hostbench")
    _build(os.path.join(ROOT, "scripts", "hostbench", "hostbench.cpp"), exe)
    r = subprocess.run([exe, "64", "30", "10"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "per stream-frame" in r.stdout and "per keyframe" in r.stdout
