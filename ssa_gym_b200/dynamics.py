"""The operator callables of the reference's env_config, evaluated on the GPU.

The reference hands five Python callables to filterpy through `env_config` (envs/__init__.py:27-28):
fx, hx, mean_z, residual_z, msqrt.  Here the same NAMES exist and are accepted by the environment BY
IDENTITY (or by `__name__`, so the reference's own function objects work too) and mapped to the
built-in device implementations inside the fused kernel; any other callable is rejected with an
explicit error — there is deliberately no path that calls back into Python per sigma point.

Called directly, each operator evaluates the DEVICE build of the function through the C ABI
(ssa_unit_* in include/ssa_ukf.h), so `fx(x, dt)` returns what the kernel computes.

Reference: envs/farnocchia.py:1053-1062 (fx), envs/dynamics.py:207-231 (hx_xyz, hx_aer_erfa),
:260-278 (residual_z_aer, residual_xyz, mean_xyz), :342-354 (mean_z_uvw), :402-417 (robust_cholesky).
"""
import ctypes

import numpy as np

from . import _lib
from .transformations import trans_uvw_ecef

_vp = ctypes.c_void_p


def _p(a):
    return a.ctypes.data_as(_vp)


def fx_xyz_farnocchia(x, dt, device=0):
    """Two-body propagation of state(s) x [.., 6] by dt seconds (mu = 398600441800000.0)."""
    lib = _lib.require_gpu()
    x = np.ascontiguousarray(x, dtype=np.float64)
    shape = x.shape
    xs = x.reshape(-1, 6)
    out = np.empty_like(xs)
    exc = np.zeros(len(xs), dtype=np.int32)
    _lib.check(lib.ssa_unit_fx(_p(xs), float(dt), _p(out), _p(exc), len(xs), device), "ssa_unit_fx")
    if exc.any():
        raise ArithmeticError("fx_xyz_farnocchia: the reference would raise inside numba for %d state(s)" % int(exc.astype(bool).sum()))
    return out.reshape(shape)


def hx_xyz(x_gcrs, trans_matrix=None, observer_lla=None, observer_itrs=None, time=None):
    return np.asarray(x_gcrs)[..., :3]


def hx_aer_erfa(x_gcrs, trans_matrix, observer_lla, observer_itrs, time=None, device=0):
    """[az, el, range] of GCRS position(s) seen from the observer, via trans_matrix (GCRS->ITRS)."""
    lib = _lib.require_gpu()
    x = np.ascontiguousarray(x_gcrs, dtype=np.float64)
    lead = x.shape[:-1]
    stride = x.shape[-1]
    xs = x.reshape(-1, stride)
    M = np.ascontiguousarray(trans_matrix, dtype=np.float64).reshape(9)
    oi = np.ascontiguousarray(observer_itrs, dtype=np.float64).reshape(3)
    T = np.ascontiguousarray(trans_uvw_ecef(observer_lla[0], observer_lla[1]), dtype=np.float64).reshape(9)
    out = np.empty((len(xs), 3))
    _lib.check(lib.ssa_unit_hx_aer(_p(xs), stride, _p(M), _p(oi), _p(T), _p(out), len(xs), device), "ssa_unit_hx_aer")
    return out.reshape(lead + (3,))


def _aer(op, a, b=None, device=0):
    lib = _lib.require_gpu()
    a = np.ascontiguousarray(a, dtype=np.float64)
    shape = a.shape
    a2 = a.reshape(-1, 3)
    b2 = None if b is None else np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), shape)).reshape(-1, 3)
    out = np.empty_like(a2)
    _lib.check(lib.ssa_unit_aer(op, _p(a2), _p(b2) if b2 is not None else None, _p(out), len(a2), device), "ssa_unit_aer")
    return out.reshape(shape)


def aer2uvw(aer):
    return _aer(0, aer)


def uvw2aer(uvw):
    return _aer(1, uvw)


def residual_z_aer(a, b):
    return _aer(2, a, b)


def residual_xyz(a, b):
    return np.subtract(a, b)


def mean_xyz(a, w):
    return np.dot(w, a)


def mean_z_uvw(sigmas, Wm):
    """uvw2aer(Wm . aer2uvw(sigmas)) with the kernel's fixed summation order (k = 0..12, FMA)."""
    uvw = aer2uvw(np.asarray(sigmas, dtype=np.float64))
    Wm = np.asarray(Wm, dtype=np.float64)
    acc = Wm[0] * uvw[0]
    for k in range(1, len(Wm)):
        # fma is not available in numpy; this direct call is a convenience, the fused kernel is the
        # authority (tests compare the kernel, not this helper, against the oracle)
        acc = acc + Wm[k] * uvw[k]
    return uvw2aer(acc)


def robust_cholesky(a, lam=1.0, device=0):
    """Upper Cholesky factor with the reference's diagonal-inflation fallback (10**i, i = -6..9)."""
    lib = _lib.require_gpu()
    a = np.asarray(a, dtype=np.float64).reshape(6, 6)
    iu = np.triu_indices(6)
    Pp = np.ascontiguousarray(a[iu]).reshape(1, 21)
    Up = np.empty_like(Pp)
    ret = np.zeros(1, dtype=np.int32)
    _lib.check(lib.ssa_unit_robust_chol(_p(Pp), float(lam), _p(Up), _p(ret), 1, device), "ssa_unit_robust_chol")
    if ret[0] < 0:
        raise np.linalg.LinAlgError
    U = np.zeros((6, 6))
    U[iu] = Up[0]
    return U


# ---- identity / name mapping used by the environment ---------------------------------------------
DEVICE_OPERATORS = {
    "fx": {"fx_xyz_farnocchia"},
    "hx": {"hx_aer_erfa", "hx_xyz"},
    "mean_z": {"mean_z_uvw", "mean_xyz", None},
    "residual_z": {"residual_z_aer", "residual_xyz", "subtract", None},
    "msqrt": {"robust_cholesky"},
}


def resolve_operator(role, fn):
    """Return the canonical device-operator name for a config callable or raise."""
    name = None if fn is None else getattr(fn, "__name__", None)
    if name in DEVICE_OPERATORS[role]:
        return name
    raise NotImplementedError(
        f"env_config['{role}'] = {fn!r} has no device implementation; supported: "
        f"{sorted(n for n in DEVICE_OPERATORS[role] if n)} (the GPU path never calls back into Python, "
        f"there is no CPU fallback)")


# the operator combination the kernels implement for each observation type (anything else would silently be a
# different filter: with mean_z = None filterpy averages the angles with np.dot, with residual_z = np.subtract it does
# not wrap the azimuth)
DEVICE_COMBINATIONS = {
    "aer": {"hx": {"hx_aer_erfa"}, "mean_z": {"mean_z_uvw"}, "residual_z": {"residual_z_aer"}},
    "xyz": {"hx": {"hx_xyz"}, "mean_z": {"mean_xyz", None}, "residual_z": {"residual_xyz", "subtract", None}},
}


def validate_operators(config, obs_type):
    """Resolve env_config's operator callables (envs/__init__.py:27-28; SS2:112-116, 211-214) to the device operators
    and check that the (hx, mean_z, residual_z) triple is the one the kernels run for `obs_type`.  Missing keys take
    the device combination's own operators.  Returns the canonical names; raises otherwise (no CPU fallback)."""
    if obs_type not in DEVICE_COMBINATIONS:
        raise ValueError('Invalid Observation Type: ' + str(obs_type))
    names = {"fx": resolve_operator("fx", config.get("fx", fx_xyz_farnocchia)),
             "msqrt": resolve_operator("msqrt", config.get("msqrt", robust_cholesky))}
    defaults = {"aer": {"hx": hx_aer_erfa, "mean_z": mean_z_uvw, "residual_z": residual_z_aer},
                "xyz": {"hx": hx_xyz, "mean_z": None, "residual_z": None}}[obs_type]
    for role in ("hx", "mean_z", "residual_z"):
        fn = config[role] if role in config and not (role == "hx" and config[role] is None) else defaults[role]
        name = resolve_operator(role, fn)
        if name not in DEVICE_COMBINATIONS[obs_type][role]:
            raise ValueError(f"env_config['{role}'] = {name} does not belong to obs_type '{obs_type}': the device filter for it "
                             f"uses {sorted(str(n) for n in DEVICE_COMBINATIONS[obs_type][role])}")
        names[role] = name
    return names
