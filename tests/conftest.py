import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "test", "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_gray(rel):
    import cv2
    img = cv2.imread(os.path.join(DATA, rel), cv2.IMREAD_GRAYSCALE)
    assert img is not None, rel
    return img


@pytest.fixture(scope="session")
def kitti_pair():
    return load_gray("images/0000000000.png"), load_gray("images/0000000001.png")


@pytest.fixture(scope="session")
def tum_pair():
    return load_gray("test_images/0.png"), load_gray("test_images/1.png")


@pytest.fixture(scope="session")
def oracle():
    from oracle import ref_oracle
    ref_oracle.build()
    return ref_oracle


@pytest.fixture(scope="session")
def gpu_ctx():
    import slam_cin0051_b200 as s
    return s.Context.default()


@pytest.fixture(scope="session")
def detector(gpu_ctx):
    import slam_cin0051_b200 as s
    return s.FeatureDetector(os.path.join(DATA, "feature_detector.yml"), gpu_ctx)


@pytest.fixture(scope="session")
def matcher(gpu_ctx):
    import slam_cin0051_b200 as s
    return s.FeatureMatcher(os.path.join(DATA, "feature_matcher.yml"), gpu_ctx)
