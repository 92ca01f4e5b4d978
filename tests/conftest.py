import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

CALIB = ROOT / "tests" / "golden" / "calib"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def calib_dir():
    return CALIB


def _cams(name):
    from svi_mapper_b200 import load_camera
    return load_camera(str(CALIB / f"{name}_left.txt")), load_camera(str(CALIB / f"{name}_right.txt"))


@pytest.fixture(scope="session")
def kitti_cams():
    return _cams("kitti_00")


@pytest.fixture(scope="session")
def kitti1112_cams():
    return _cams("kitti_11_12")


@pytest.fixture(scope="session")
def vi_cams():
    return _cams("vi_sensor")
