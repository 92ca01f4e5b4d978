import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

CALIB = ROOT / "tests" / "golden" / "calib"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _cuda_devices() -> int:
    """Number of CUDA devices the driver reports (0 without a driver); does not need the built library."""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not errored) on a machine without a CUDA device."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_library():
    """libsvi_gpu.so and the C++ facade demo (binaries are git-ignored): built once per session when missing or stale
    (nvcc cross-compiles without a GPU).  Requested by the GPU tests and by the host tests that load the library --
    the pure numpy / cv2 oracle tests run on a machine without nvcc."""
    from svi_mapper_b200 import build as b
    if b.needs_build():
        b.build_library()
    if not (ROOT / "svi_mapper_b200" / "host" / "facade_demo").exists():
        b.build_host_demo()
    return b.LIB


@pytest.fixture(scope="session")
def built_oracle():
    """oracle/libsvi_oracle.so (gcc), the C restatement used as checker."""
    from oracle import c_oracle
    c_oracle.load(native=False)   # builds when missing or older than its source
    return True


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
