"""Shared test infrastructure: loads the oracle (oracle/liboracle.so), the host twin
(tests/twin/libssa_twin.so) and the product (ssa_gym_b200) and runs the same batch step through each."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ssa_gym_b200 import _lib  # noqa: E402
from ssa_gym_b200.transformations import arcsec2rad, lla2ecef, trans_uvw_ecef  # noqa: E402
from ssa_gym_b200.ukf import Q_discrete_white_noise_block, merwe_weights  # noqa: E402

VP = ctypes.c_void_p
TWIN_DIR = os.path.join(ROOT, "tests", "twin")
TWIN_LIB = os.path.join(TWIN_DIR, "libssa_twin.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# SOFA cookbook matrix quoted by the reference (tests.py:107-109), used as a fixed trans_matrix
CEL2TER06AXY = np.array([[+0.973104317697536, +0.230363826239128, -0.000703163481769],
                         [-0.230363800456036, +0.973104570632801, +0.000118545368117],
                         [+0.000711560162594, +0.000046626402444, +0.999999745754024]])
OBSERVER_DEG = (38.828198, -77.305352, 20.0)  # envs/__init__.py:24
X6 = np.array([34090858.3, 23944774.4, 6503066.82, -1983.785080, 2150.41744, 913.881611])  # tests.py:130


def _newer(src_list, target):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list if os.path.isfile(s))


def build_twin():
    csrc = os.path.join(ROOT, "ssa_gym_b200", "csrc")
    deps = [os.path.join(TWIN_DIR, "twin.cpp")] + [os.path.join(csrc, h) for h in
                                                    ("ssa_math.h", "ssa_orbit.h", "ssa_meas.h", "ssa_ukf_core.h")]
    if _newer(deps, TWIN_LIB):
        gxx = "/usr/bin/g++" if os.path.isfile("/usr/bin/g++") else "g++"
        cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-mavx2", "-fopenmp", "-shared", "-fPIC",
               "-Wno-unknown-pragmas", "-o", TWIN_LIB, os.path.join(TWIN_DIR, "twin.cpp")]
        subprocess.run(cmd, check=True, cwd=TWIN_DIR)
    return TWIN_LIB


def build_oracle():
    deps = [os.path.join(ORACLE_DIR, "ukf_oracle.c"), os.path.join(ROOT, "include", "ssa_ukf.h")]
    if _newer(deps, ORACLE_LIB):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return ORACLE_LIB


_cache = {}


def twin():
    if "twin" not in _cache:
        _cache["twin"] = ctypes.CDLL(build_twin())
    return _cache["twin"]


def oracle():
    if "oracle" not in _cache:
        _cache["oracle"] = ctypes.CDLL(build_oracle())
    return _cache["oracle"]


def p(a):
    return None if a is None else a.ctypes.data_as(VP)


def make_cfg(N, E=None, m=None, dt=20.0, alpha=1e-4, beta=2.0, kappa=-3.0, q_sigma=0.000025, R=None,
             observer_deg=OBSERVER_DEG, obs_limit_deg=-90.0, obs_type="aer", resample=True, reward_type="jones",
             n_steps=480):
    """ssa_ukf_cfg with the reference's default env_config (envs/__init__.py:23-28)."""
    if E is None:
        E, m = 1, N
    Wm, Wc, lam = merwe_weights(6, alpha, beta, kappa)
    c = _lib.SsaUkfCfg()
    c.abi_version = 1
    c.n_objects, c.n_envs, c.m = N, E, m
    c.obs_type = {"aer": 0, "xyz": 1}[obs_type]
    c.resample_after_predict = 1 if resample else 0
    c.reward_type = {"jones": 0, "trinary": 1, "shaped": 2}[reward_type]
    c.n_steps = n_steps
    c.dt, c.lam_plus_n = dt, lam
    for i in range(13):
        c.Wm[i], c.Wc[i] = Wm[i], Wc[i]
    Q = Q_discrete_white_noise_block(dt, q_sigma ** 2)
    for i, v in enumerate(Q.ravel()):
        c.Q[i] = v
    if R is None:
        R = np.diag([arcsec2rad ** 2] * 2 + [1e3 ** 2])
    R = np.asarray(R, dtype=float)
    if R.ndim == 1:
        R = np.tile(R, (3, 1))
    for i, v in enumerate(R.ravel()):
        c.R[i] = v
    lla = np.array([np.radians(observer_deg[0]), np.radians(observer_deg[1]), observer_deg[2]])
    oi = lla2ecef(lla)
    T = trans_uvw_ecef(lla[0], lla[1])
    for i in range(3):
        c.obs_itrs[i] = oi[i]
    for i, v in enumerate(np.asarray(T, dtype=float).ravel()):
        c.T[i] = v
    c.obs_limit = np.radians(obs_limit_deg)
    return c


IU = np.triu_indices(6)


def pack_P(Pfull):
    return np.ascontiguousarray(np.asarray(Pfull).reshape(-1, 6, 6)[:, IU[0], IU[1]])


def unpack_P(Ppacked):
    Pp = np.asarray(Ppacked).reshape(-1, 21)
    out = np.zeros((len(Pp), 6, 6))
    out[:, IU[0], IU[1]] = Pp
    out[:, IU[1], IU[0]] = Pp
    return out


class HostState:
    """Arrays of one batch in the reference's host layout (P full [N,6,6])."""

    def __init__(self, x_true, x, P):
        N = len(x_true)
        self.N = N
        self.x_true = np.ascontiguousarray(x_true, dtype=np.float64).copy()
        self.x = np.ascontiguousarray(x, dtype=np.float64).copy()
        P = np.asarray(P, dtype=np.float64)
        self.P = np.ascontiguousarray(np.broadcast_to(P, (N, 6, 6))).copy()
        self.status = np.zeros(N, np.int32)
        self.infl = np.zeros(N, np.int32)
        self.obs = np.zeros((N, 12))
        self.dpos, self.dvel, self.spos, self.svel, self.trace = (np.zeros(N) for _ in range(5))
        self.z_true = np.full((N, 3), np.nan)
        self.y = np.full((N, 3), np.nan)
        self.S = np.full((N, 3, 3), np.nan)
        self.sigmas_h = np.zeros((N, 13, 3))
        self.visible = np.zeros(N, np.uint8)
        self.updated = np.zeros(N, np.uint8)


def cpu_step(which, cfg, st, M, flags, actions=None, z_noise=None):
    """One batch step through the oracle ('oracle', P full) or the host twin ('twin', P packed)."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    act = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
    zn = None if z_noise is None else np.ascontiguousarray(z_noise, dtype=np.float64)
    if which == "oracle":
        fn, P = oracle().oracle_step, st.P
    else:
        fn, P = twin().twin_step, pack_P(st.P)
    fn(ctypes.byref(cfg), p(M), ctypes.c_int(flags), p(st.x_true), p(st.x), p(P), p(st.status), p(st.infl), p(act), p(zn),
       p(st.obs), p(st.dpos), p(st.dvel), p(st.spos), p(st.svel), p(st.trace), p(st.z_true), p(st.y), p(st.S),
       p(st.sigmas_h), p(st.visible), p(st.updated))
    if which != "oracle":
        st.P = unpack_P(P)
    return st


def lib_fx(which, x, dt):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 6)
    out = np.empty_like(x)
    exc = np.zeros(len(x), np.int32)
    if which == "gpu":
        _lib.check(_lib.require_gpu().ssa_unit_fx(p(x), float(dt), p(out), p(exc), len(x), 0), "ssa_unit_fx")
    else:
        L, name = (oracle(), "oracle_fx") if which == "oracle" else (twin(), "twin_fx")
        getattr(L, name)(p(x), ctypes.c_double(dt), p(out), p(exc), ctypes.c_int(len(x)))
    return out, exc


def lib_hx(which, x, M, obs_itrs, T):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 6)
    M = np.ascontiguousarray(M, dtype=np.float64)
    oi = np.ascontiguousarray(obs_itrs, dtype=np.float64)
    T = np.ascontiguousarray(T, dtype=np.float64)
    out = np.empty((len(x), 3))
    if which == "gpu":
        _lib.check(_lib.require_gpu().ssa_unit_hx_aer(p(x), 6, p(M), p(oi), p(T), p(out), len(x), 0), "ssa_unit_hx_aer")
    else:
        L, name = (oracle(), "oracle_hx_aer") if which == "oracle" else (twin(), "twin_hx_aer")
        getattr(L, name)(p(x), p(M), p(oi), p(T), p(out), ctypes.c_int(len(x)))
    return out


MATH_OPS = ["sin", "cos", "tan", "atan", "asin", "acos", "exp", "log", "sinh", "cosh", "tanh", "atanh", "asinh",
            "acosh", "pow23", "atan2", "pymod"]


def lib_math(which, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    out = np.empty_like(a)
    if which == "gpu":
        _lib.check(_lib.require_gpu().ssa_unit_math(MATH_OPS.index(op), p(a), p(b), p(out), a.size, 0), "ssa_unit_math")
    else:
        fn = getattr(twin(), "twin_" + op)
        if b is None:
            fn(p(a), p(out), ctypes.c_int(a.size))
        else:
            fn(p(a), p(b), p(out), ctypes.c_int(a.size))
    return out


def c2_inputs(N=20000, steps=1, catalog=None):
    """SURVEY 8(d) C2 inputs: x_filter = x_true + RandomState(0).normal * x_sigma, P0, z_noise from RandomState(1)."""
    from ssa_gym_b200.catalog import synthetic_catalog
    cat = synthetic_catalog(N, 0) if catalog is None else np.asarray(catalog)[:N]
    x = cat + np.random.RandomState(0).normal(size=(N, 6)) * np.array([1e5] * 3 + [1e2] * 3)
    P0 = np.diag([1e10] * 3 + [1e4] * 3)
    zn = np.random.RandomState(1).normal(size=(steps, N, 3)) * np.array([arcsec2rad, arcsec2rad, 1e3])
    return cat, x, P0, zn


def bits_equal(a, b):
    """Bit-for-bit equality of float arrays.  NaNs must sit in the same places; their sign/payload is not
    compared (an x86 SSE operation generates the 'indefinite' NaN 0xFFF8..., an sm_100a one 0x7FF8...)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        if not np.array_equal(na, nb):
            return False
        return np.array_equal(np.where(na, 0.0, a).view(np.uint64), np.where(nb, 0.0, b).view(np.uint64))
    return np.array_equal(a, b)


def gpu_available():
    try:
        return _lib.load().ssa_ukf_device_count() > 0
    except Exception:
        return False


class TwinBackedUKF:
    """Same interface as ssa_gym_b200.ukf.BatchedUKF, executed by the host twin (tests only).  Lets the tests run
    the SAME environment code once on the GPU and once on the twin and demand equality of whole episodes."""

    def __init__(self, n_envs, m, dt, Q, R, obs_lla, obs_limit_rad, alpha=1e-4, beta=2.0, kappa=-3.0, obs_type="aer",
                 reward_type="jones", n_steps=480, resample_after_predict=True, device=0):
        from ssa_gym_b200 import _lib as F
        self.F = F
        self.N, self.m, self.n_envs = n_envs * m, m, n_envs
        Wm, Wc, lam = merwe_weights(6, alpha, beta, kappa)
        c = _lib.SsaUkfCfg()
        c.abi_version, c.n_objects, c.n_envs, c.m = 1, self.N, n_envs, m
        c.obs_type = {"aer": 0, "xyz": 1}[obs_type]
        c.resample_after_predict = 1 if resample_after_predict else 0
        c.reward_type, c.n_steps, c.dt, c.lam_plus_n = 0, n_steps, dt, lam
        for i in range(13):
            c.Wm[i], c.Wc[i] = Wm[i], Wc[i]
        for i, v in enumerate(np.asarray(Q, dtype=float).ravel()):
            c.Q[i] = v
        R = np.asarray(R, dtype=float)
        if R.ndim == 1:
            R = np.tile(R, (3, 1))
        for i, v in enumerate(R.ravel()):
            c.R[i] = v
        lla = np.asarray(obs_lla, dtype=float)
        for i, v in enumerate(lla2ecef(lla)):
            c.obs_itrs[i] = v
        for i, v in enumerate(np.asarray(trans_uvw_ecef(lla[0], lla[1]), dtype=float).ravel()):
            c.T[i] = v
        c.obs_limit = obs_limit_rad
        self.cfg = c
        self.st = None
        self.actions = np.zeros(n_envs, np.int32)
        self.z_noise = np.zeros((self.N, 3))

    def reset(self, x_true, x_filter, P0, stream=None):
        self.st = HostState(np.reshape(x_true, (self.N, 6)), np.reshape(x_filter, (self.N, 6)), P0)

    def upload(self, field, arr, stream=None):
        F = self.F
        if field == F.F_ACTIONS:
            self.actions = np.ascontiguousarray(arr, dtype=np.int32).reshape(self.n_envs)
        elif field == F.F_Z_NOISE:
            self.z_noise = np.ascontiguousarray(arr, dtype=np.float64).reshape(self.N, 3)
        else:
            raise NotImplementedError

    def step(self, M, flags, stream=None):
        cpu_step("twin", self.cfg, self.st, np.asarray(M, dtype=float).reshape(3, 3), flags, actions=self.actions,
                 z_noise=self.z_noise)

    def download(self, field, out=None, stream=None):
        F, st = self.F, self.st
        m = {F.F_X_TRUE: st.x_true, F.F_X_FILTER: st.x, F.F_P_FILTER: st.P, F.F_OBS: st.obs, F.F_DELTA_POS: st.dpos,
             F.F_DELTA_VEL: st.dvel, F.F_SIGMA_POS: st.spos, F.F_SIGMA_VEL: st.svel, F.F_TRACE: st.trace,
             F.F_Z_TRUE: st.z_true, F.F_Y: st.y, F.F_S: st.S, F.F_SIGMAS_H: st.sigmas_h, F.F_VISIBLE: st.visible,
             F.F_UPDATED: st.updated, F.F_STATUS: st.status, F.F_INFLATIONS: st.infl}
        return m[field].copy()

    def snapshot(self, stream=None):
        st = self.st
        return {"x_true": st.x_true.copy(), "x_filter": st.x.copy(), "P_filter": st.P.copy(), "obs": st.obs.copy(),
                "delta_pos": st.dpos.copy(), "delta_vel": st.dvel.copy(), "sigma_pos": st.spos.copy(),
                "sigma_vel": st.svel.copy(), "z_true": st.z_true.copy(), "y": st.y.copy(), "S": st.S.copy(),
                "sigmas_h": st.sigmas_h.copy(), "status": st.status.copy(), "visible": st.visible.copy(),
                "updated": st.updated.copy()}

    def sync(self, stream=None):
        pass

    def close(self):
        pass
