"""include/svslam.h is a plain-C interface: examples/c_abi_demo.c (C99, -pedantic) compiles and links against
libsvslam.so; without a B200 it fails loudly (no CPU fallback), with one it runs the GFTT -> LK -> triangulation seams."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "stereovision-slam_b200")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_demo")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + LIBDIR, "-lsvslam", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_c_demo_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "svs_create failed" in r.stderr


@pytest.mark.gpu
def test_c_demo_runs_the_seams(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "triangulated" in r.stdout
