"""The C-ABI shared library loads and exports every symbol include/svslam.h declares (CPU only;
no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "svslam.h")).read()
    return sorted(set(re.findall(r"SVS_API[^;]*?\b(svs_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_seams():
    syms = declared_symbols()
    for s in ("svs_gftt_detect", "svs_lk_track", "svs_triangulate", "svs_pose_only_lm", "svs_ba_optimize",
              "svs_stereo_bm", "svs_backproject", "svs_half_nearest"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import svslam
    lib = svslam.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_no_device_fails_loudly():
    """Without a B200 the product refuses to run (no CPU fallback)."""
    import svslam
    lib = svslam.load_library()
    n = ctypes.c_int(0)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    try:
        svslam.Context(0)
    except svslam.SvsError as e:
        assert "svs_create failed" in str(e)
    else:
        raise AssertionError("Context() must raise without a GPU")


def test_product_does_not_reference_the_oracle():
    """The shipped path never imports, links or loads anything under oracle/."""
    pkg = os.path.join(ROOT, "stereovision-slam_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|liboracle|oracle/|orc_")
    for dp, _, fs in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                s = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(s), f
    # tools and examples are not allowed to lean on the checker either (only tests/, smoke() and bench.py's CPU legs are)
    for sub in ("scripts", "examples", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, sub)):
            for f in fs:
                if f.endswith((".py", ".c", ".cpp", ".h", ".sh")):
                    s = open(os.path.join(dp, f), errors="ignore").read()
                    assert not bad.search(s), os.path.join(sub, f)
