"""Build recipe for libsvi_gpu.so (nvcc, sm_100a only) -- in-tree, so the .so travels with the repo
snapshot to the GPU box.  `python -m svi_mapper_b200.build` or __graft_entry__.build()."""
from __future__ import annotations

import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libsvi_gpu.so"

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",          # Harris must round every fp32 op on its own (cv2 parity)
    "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",   # host-side fp64 geometry: same rounding as the kernels, whatever -march is used
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not pathlib.Path(exe).exists():
        raise RuntimeError("nvcc not found: libsvi_gpu.so cannot be built")
    return exe


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "svi_gpu.h"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in sources())


def build_library(force: bool = False, verbose: bool = False) -> pathlib.Path:
    sys.path.insert(0, str(ROOT / "tools"))
    try:
        import gen_pattern_header
        gen_pattern_header.main()
    finally:
        sys.path.pop(0)
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "svi_gpu.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


def build_host_demo(out: pathlib.Path | None = None, flags=("-O2", "-ffp-contract=off")) -> pathlib.Path:
    """g++ build of the C++ host facade's demo driver (links libsvi_gpu.so through an rpath).  The documented build
    line carries -ffp-contract=off; the headers also pin it themselves (pragma), so that the reference's own
    `-O3 -march=native` (CMakeLists.txt:51) gives the same numbers -- tests build it that way too."""
    host = PKG / "host"
    exe = out or host / "facade_demo"
    cmd = ["g++", "-std=c++17", *flags, "-Wall", "-Wextra", "-o", str(exe), str(host / "facade_demo.cpp"),
           "-L" + str(PKG), "-lsvi_gpu", "-Wl,-rpath," + str(PKG)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return exe


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_demo())
