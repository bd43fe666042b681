"""CPU: the C-ABI shared library loads without a GPU, exports every function include/ssa_ukf.h declares, and the
product refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import helpers as H
from ssa_gym_b200 import _build, _lib


def header_functions():
    src = open(os.path.join(H.ROOT, "include", "ssa_ukf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(?:int|long|size_t|const char\*)\s+\*?\s*(ssa_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_is_built_in_tree_for_sm_100a():
    assert os.path.isfile(_build.LIB), "build with: python -c 'import __graft_entry__ as g; g.build()'"
    assert "compute_100a" in " ".join(_build.NVCC_FLAGS) and "-fmad=false" in _build.NVCC_FLAGS


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_build.LIB)
    names = header_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ssa_ukf.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_struct_layout_matches_header():
    # 8 int32 + (2 + 13 + 13 + 36 + 9 + 3 + 9 + 1) doubles
    assert ctypes.sizeof(_lib.SsaUkfCfg) == 8 * 4 + 86 * 8
    assert _lib.load().ssa_ukf_abi_version() == 1


def test_no_cpu_fallback():
    if H.gpu_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.SsaUkfError):
        _lib.require_gpu()
    from ssa_gym_b200 import dynamics
    with pytest.raises(_lib.SsaUkfError):
        dynamics.fx_xyz_farnocchia(H.X6, 20.0)
    cfg = H.make_cfg(8)
    h = ctypes.c_void_p()
    rc = _lib.load().ssa_ukf_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == _lib.SSA_ENODEV and b"no CPU fallback" in _lib.load().ssa_ukf_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under ssa_gym_b200/ may reference it."""
    pkg = os.path.join(H.ROOT, "ssa_gym_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                assert "libssa_twin" not in txt, f


def test_unknown_operator_is_rejected():
    from ssa_gym_b200 import dynamics
    with pytest.raises(NotImplementedError):
        dynamics.resolve_operator("fx", lambda x, dt: x)
    assert dynamics.resolve_operator("hx", dynamics.hx_aer_erfa) == "hx_aer_erfa"


def test_plain_c_program_links_against_the_header_and_library(tmp_path):
    """The boundary is a C ABI: include/ssa_ukf.h is valid C99 and a plain-C caller links against libssa_ukf.so, reads the
    ABI version, asks for the device count and produces the path's per-step host input (the trans_matrix table)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text('#include <stdio.h>\n#include "ssa_ukf.h"\n'
                   'int main(void) {\n  double out[18];\n'
                   '  int rc = ssa_trans_matrix_table(2020, 5, 4, 0.0, 20.0, 2, NULL, 0, out);\n'
                   '  printf("%d %d %.17g %.17g\\n", rc, ssa_ukf_abi_version(), out[0], out[9]);\n'
                   '  return ssa_ukf_device_count() < 0;\n}\n')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_build.LIB)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(H.ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lssa_ukf", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == "0" and out[1] == "1"
    from datetime import datetime
    from ssa_gym_b200 import transformations as T
    ref = T.gcrs2irts_matrix_native(datetime(2020, 5, 4), 20.0, 2)
    assert float(out[2]) == ref[0, 0, 0] and float(out[3]) == ref[1, 0, 0]
