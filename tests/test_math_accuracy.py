"""CPU: accuracy of the product's own fp64 elementary functions (ssa_gym_b200/csrc/ssa_math.h, host build)
against mpmath, in ulps, on the argument ranges the path uses.  The same source runs on the GPU bit for bit
(tests/test_gpu_bitexact.py), so these bounds hold for the kernels."""
import numpy as np
import pytest

import helpers as H

mp = pytest.importorskip("mpmath")
mp.mp.prec = 200


def ulp_err(y, ref):
    out = []
    for yi, ri in zip(y, ref):
        rf = float(ri)
        if rf == 0:
            out.append(0.0 if yi == 0 else np.inf)
            continue
        out.append(float(abs(mp.mpf(float(yi)) - ri) / mp.mpf(float(np.spacing(abs(rf))))))
    return np.array(out)


RNG = np.random.default_rng(0)
N = 4000
CASES = [
    ("sin", mp.sin, np.concatenate([RNG.uniform(-7, 7, N), RNG.uniform(-1e3, 1e3, N), np.pi * np.arange(-8, 9) / 2]), 1.6),
    ("cos", mp.cos, np.concatenate([RNG.uniform(-7, 7, N), RNG.uniform(-1e3, 1e3, N), np.pi * np.arange(-8, 9) / 2]), 1.6),
    ("tan", mp.tan, RNG.uniform(-1.57, 1.57, N), 3.0),
    ("atan", mp.atan, np.concatenate([RNG.uniform(-5, 5, N), RNG.standard_cauchy(N) * 10]), 1.1),
    ("asin", mp.asin, np.concatenate([RNG.uniform(-1, 1, N), 1 - 10.0 ** RNG.uniform(-16, 0, 500), [1.0, -1.0, 0.5]]), 1.3),
    ("acos", mp.acos, np.concatenate([RNG.uniform(-1, 1, N), 1 - 10.0 ** RNG.uniform(-16, 0, 500), [1.0, -1.0, 0.5]]), 1.3),
    ("exp", mp.exp, RNG.uniform(-20, 20, N), 1.1),
    ("log", mp.log, 10.0 ** RNG.uniform(-10, 10, N), 1.1),
    ("sinh", mp.sinh, RNG.uniform(-5, 5, N), 3.0),
    ("cosh", mp.cosh, RNG.uniform(-5, 5, N), 2.0),
    ("tanh", mp.tanh, RNG.uniform(-5, 5, N), 3.5),
    ("atanh", mp.atanh, RNG.uniform(-0.999, 0.999, N), 3.5),
    ("asinh", mp.asinh, RNG.uniform(-5, 5, N), 3.5),
    ("acosh", mp.acosh, RNG.uniform(1, 10, N), 3.5),
]


@pytest.mark.parametrize("op,fn,x,bound", CASES, ids=[c[0] for c in CASES])
def test_unary_ulp(op, fn, x, bound):
    y = H.lib_math("twin", op, x)
    e = ulp_err(y, [fn(mp.mpf(float(v))) for v in x])
    assert e.max() <= bound, (op, e.max())


def test_atan2_ulp_and_special_values():
    a = RNG.standard_normal(N) * 10.0 ** RNG.uniform(-3, 8, N)
    b = RNG.standard_normal(N) * 10.0 ** RNG.uniform(-3, 8, N)
    y = H.lib_math("twin", "atan2", a, b)
    e = ulp_err(y, [mp.atan2(mp.mpf(float(u)), mp.mpf(float(v))) for u, v in zip(a, b)])
    assert e.max() <= 1.5
    sa = np.array([0.0, 0.0, -0.0, 1.0, -1.0, 0.0, -0.0, 0.0])
    sb = np.array([1.0, -1.0, -1.0, 0.0, 0.0, 0.0, 0.0, -0.0])
    assert np.array_equal(H.lib_math("twin", "atan2", sa, sb), np.arctan2(sa, sb))


def test_python_mod_is_exact():
    for scale in (1, 1e3, 1e6, 1e12, 1e17, 1e300):
        a = np.concatenate([RNG.standard_normal(20000) * scale, np.arange(-50, 50) * 2 * np.pi,
                            np.nextafter(np.arange(-50, 50) * 2 * np.pi, 0)])
        m = np.full(a.size, 2 * np.pi)
        assert np.array_equal(H.lib_math("twin", "pymod", a, m), a % (2 * np.pi))
