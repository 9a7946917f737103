import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stereovision-slam_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    import svslam
    c = svslam.Context(0)     # raises (never falls back) when there is no B200
    yield c
    c.close()


@pytest.fixture(scope="session")
def granule():
    """SIMD granule of this host's OpenCV build, calibrated at run time."""
    import cv2
    from oracle import cv_stages
    return cv_stages.calibrate_granule(cv2)
