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
    "-Xcompiler", "-pthread",           # svi_multi: one host thread per GPU
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


def build_library(force: bool = False, verbose: bool = False, out: pathlib.Path | None = None, defines=(),
                  pattern_header: pathlib.Path | None = None) -> pathlib.Path:
    """nvcc build of libsvi_gpu.so.  `out` / `defines` / `pattern_header` build a variant next to the product library:
    tuning experiments (-DNAME=value) and libraries with another BRIEF pair table (tools/gen_pattern_header.py --table)."""
    sys.path.insert(0, str(ROOT / "tools"))
    try:
        import gen_pattern_header
        gen_pattern_header.main([])
    finally:
        sys.path.pop(0)
    lib = pathlib.Path(out) if out else LIB
    if not force and not out and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(lib), str(CSRC / "svi_gpu.cu")]
    if pattern_header:
        cmd.insert(1, f"-DSVI_BRIEF_PATTERN_HEADER=\"{pattern_header}\"")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return lib


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


CHECKED_LIB = ROOT / "build" / "libsvi_gpu_checked.so"


def build_checked(force: bool = False) -> pathlib.Path:
    """The library with every run-time index asserted in range (-DSVI_BOUNDS_CHECK, see csrc/common.cuh); the parity tests
    run against it once (tests/test_gpu_parity.py::test_bounds_checked_build) -- compute-sanitizer is closed on the GPU pool."""
    newest = max(p.stat().st_mtime for p in sources())
    if force or not CHECKED_LIB.exists() or CHECKED_LIB.stat().st_mtime < newest:
        CHECKED_LIB.parent.mkdir(parents=True, exist_ok=True)
        build_library(force=True, out=CHECKED_LIB, defines=["SVI_BOUNDS_CHECK"])
    return CHECKED_LIB


ALT_TABLES = {"alt_random": ROOT / "tests" / "golden" / "patterns" / "alt_random.txt",
              "alt_adversarial": ROOT / "tests" / "golden" / "patterns" / "alt_adversarial.txt"}
ALT_DIR = ROOT / "build" / "alt"   # git-ignored, travels to the GPU box with the snapshot


def build_with_table(table: pathlib.Path, out_dir: pathlib.Path) -> dict:
    """libsvi_gpu.so compiled around another BRIEF pair table -- the one-command table swap: header from the text
    table, then the ordinary build line with -DSVI_BRIEF_PATTERN_HEADER pointing at it.  Returns the paths."""
    sys.path.insert(0, str(ROOT / "tools"))
    try:
        import gen_pattern_header
    finally:
        sys.path.pop(0)
    out_dir = pathlib.Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    hdr = out_dir / "brief_pattern_32.h"
    gen_pattern_header.write(table, hdr)
    lib = build_library(force=True, out=out_dir / "libsvi_gpu.so", pattern_header=hdr)
    return {"table": pathlib.Path(table), "header": hdr, "lib": lib}


def build_alt_tables(force: bool = False) -> dict:
    """The two test tables of tests/golden/patterns (see tools/make_alt_patterns.py), built once."""
    out = {}
    newest = max(p.stat().st_mtime for p in sources())
    for name, table in ALT_TABLES.items():
        d = ALT_DIR / name
        lib = d / "libsvi_gpu.so"
        have = lib.exists() and lib.stat().st_mtime >= max(newest, table.stat().st_mtime)
        out[name] = {"table": table, "header": d / "brief_pattern_32.h", "lib": lib} if have and not force else build_with_table(table, d)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="build libsvi_gpu.so (+ the facade demo); --table/--lib builds a variant around another BRIEF pair table")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--table")
    ap.add_argument("--lib", help="output directory of the variant (libsvi_gpu.so, brief_pattern_32.h)")
    ap.add_argument("-D", action="append", default=[], help="extra -D for a tuning variant; needs --out")
    ap.add_argument("--out", help="output path of a tuning variant")
    a = ap.parse_args()
    if a.table:
        print(build_with_table(pathlib.Path(a.table), pathlib.Path(a.lib or (ALT_DIR / pathlib.Path(a.table).stem))))
    elif a.out:
        print(build_library(force=True, verbose=a.v, out=pathlib.Path(a.out), defines=a.D))
    else:
        print(build_library(force=a.force, verbose=a.v))
        print(build_host_demo())
