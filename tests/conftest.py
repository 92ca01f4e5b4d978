import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

CALIB = ROOT / "tests" / "golden" / "calib"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """A fresh checkout has no binaries (they are git-ignored): build the library, the C++ facade demo and the C checker
    once per session (nvcc cross-compiles without a GPU; nothing is rebuilt when the files are up to date)."""
    from svi_mapper_b200 import build as b
    if b.needs_build():
        b.build_library()
    if not (ROOT / "svi_mapper_b200" / "host" / "facade_demo").exists():
        b.build_host_demo()
    from oracle import c_oracle
    if not (ROOT / "oracle" / "libsvi_oracle.so").exists():
        c_oracle.build(native=False)


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
