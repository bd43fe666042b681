"""Build recipe for libssa_ukf.so (in-tree, sm_100a only).

The library is compiled with `-fmad=false`: every fused multiply-add on the path is written
explicitly in csrc/*.h, which is what makes the device arithmetic reproducible bit for bit by the
host twin used in the tests (see csrc/ssa_math.h).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libssa_ukf.so")
SOURCES = ["ssa_ukf.cu"]
HEADERS = ["ssa_math.h", "ssa_orbit.h", "ssa_meas.h", "ssa_ukf_core.h", "ssa_tile.cuh", "ssa_frames.h", "ssa_rng.h", os.path.join("..", "..", "include", "ssa_ukf.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # no implicit contraction: explicit FMAs only (bit-exact host twin)
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "177",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    return None


def is_stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile csrc/ssa_ukf.cu -> csrc/libssa_ukf.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libssa_ukf.so (no CPU fallback exists)")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB
