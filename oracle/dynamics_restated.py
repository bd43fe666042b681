"""ORACLE (test infrastructure): the five operator callables of envs/dynamics.py and the njit geometry of
envs/transformations.py, restated in numpy.  envs/dynamics.py itself cannot be imported here (it pulls
astropy, poliastro and pymap3d at module level), envs/transformations.py can (oracle/ref_loader.py) and is
used to validate these restatements and to generate the golden vectors.

Each function cites the reference file:line it follows.
"""
import ctypes
import os

import numpy as np
from numpy import arcsin as asin, arctan2 as atan2, cos, pi, sin, sqrt
import scipy.linalg

tau = 2 * pi
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_E = sqrt(WGS84_F * (2 - WGS84_F))
arcsec2rad = pi / 648000


def lla2ecef(obs_lla, a=WGS84_A, f=WGS84_F, e=WGS84_E):  # transformations.py:216-235
    lat, lon, alt = obs_lla[0], obs_lla[1], obs_lla[2]
    N = a / np.sqrt(1 - e ** 2 * sin(lat) ** 2)
    return np.array([(N + alt) * cos(lat) * cos(lon), (N + alt) * cos(lat) * sin(lon), (N * (1 - e ** 2) + alt) * sin(lat)])


def ecef2aer(obs_lla, ecef_sat, ecef_obs):  # transformations.py:329-352
    lat, lon = obs_lla[0], obs_lla[1]
    trans_uvw_ecef = np.array([[-sin(lat) * cos(lon), -sin(lon), cos(lat) * cos(lon)],
                               [-sin(lat) * sin(lon), cos(lon), cos(lat) * sin(lon)],
                               [cos(lat), 0, sin(lat)]])
    delta_ecef = ecef_sat - ecef_obs
    R_enz = trans_uvw_ecef.T @ delta_ecef
    r = sqrt(np.sum(delta_ecef ** 2))
    az = atan2(R_enz[1], R_enz[0])
    if az < 0:
        az = az + 2 * pi
    el = asin(R_enz[2] / r)
    return np.array([az, el, r])


def aer2uvw(aer):  # transformations.py:283-297
    az, el, r = aer
    return np.array([r * cos(el) * cos(az), r * cos(el) * sin(az), r * sin(el)])


def uvw2aer(uvw):  # transformations.py:300-316
    u, v, w = uvw
    r = sqrt(np.sum(np.asarray(uvw) ** 2))
    az = atan2(v, u)
    if az < 0:
        az = az + tau
    el = asin(w / r)
    return np.array([az, el, r])


def make_operators(tr=None):
    """Return (hx_aer_erfa, residual_z_aer, mean_z_uvw, hx_xyz) bound to the geometry functions of `tr`
    (the reference's transformations module) or to the restatements above when tr is None."""
    _ecef2aer = tr.ecef2aer if tr is not None else ecef2aer
    _aer2uvw = tr.aer2uvw if tr is not None else aer2uvw
    _uvw2aer = tr.uvw2aer if tr is not None else uvw2aer

    def hx_aer_erfa(x_gcrs, trans_matrix, observer_lla, observer_itrs, time=None):  # dynamics.py:219-231
        x_itrs = trans_matrix @ x_gcrs[:3]
        return _ecef2aer(observer_lla, x_itrs, observer_itrs)

    def residual_z_aer(a, b):  # dynamics.py:260-267
        c = np.empty(a.shape)
        c[0] = np.arctan2(np.sin(a[0] - b[0]), np.cos(a[0] - b[0]))
        c[1] = a[1] - b[1]
        c[2] = a[2] - b[2]
        return c

    def mean_z_uvw(sigmas, Wm):  # dynamics.py:342-354
        aers = np.empty(shape=sigmas.shape)
        for i in range(len(aers)):
            aers[i] = _aer2uvw(sigmas[i])
        uvw_mean = np.dot(Wm, aers)
        return _uvw2aer(uvw_mean)

    def hx_xyz(x_gcrs, trans_matrix=None, observer_lla=None, observer_itrs=None, time=None):  # dynamics.py:207-217
        return x_gcrs[:3]

    return hx_aer_erfa, residual_z_aer, mean_z_uvw, hx_xyz


def robust_cholesky(a):  # dynamics.py:402-417
    try:
        return scipy.linalg.cholesky(a)
    except Exception:
        i = -6
        done = False
        while not done:
            e = np.eye(len(a)) * 10 ** i
            try:
                return scipy.linalg.cholesky(a + e)
            except Exception:
                i += 1
            if i == 10:
                done = True
    raise np.linalg.LinAlgError


# ---- fx through the C oracle (used where the reference tree is absent, e.g. on the GPU box) ------------
def oracle_fx_callable():
    here = os.path.dirname(os.path.abspath(__file__))
    L = ctypes.CDLL(os.path.join(here, "liboracle.so"))
    vp = ctypes.c_void_p

    def fx(x, dt):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(6)
        exc = np.zeros(1, np.int32)
        L.oracle_fx(x.ctypes.data_as(vp), ctypes.c_double(dt), out.ctypes.data_as(vp), exc.ctypes.data_as(vp), 1)
        if exc[0]:
            raise AssertionError("fx raised inside numba in the reference")
        return out
    return fx


# ---- envs/results.py:36-84, 431-433 and envs/reward.py:6-50 ----------------------------------------------
def observations(filters_x, filters_P):  # results.py:60-72
    n = len(filters_x)
    observation = np.zeros((n, 12))
    for i in range(n):
        observation[i, :6] = filters_x[i]
        observation[i, 6:] = np.diag(filters_P[i])
    return observation


def dist3d(u, v):  # results.py:75-78
    return np.sqrt(np.sum((u - v) ** 2, axis=1))


def var3d(u):  # results.py:81-84
    return np.sqrt(np.sum(u, axis=1))


def error(states, obs):  # results.py:36-47
    return dist3d(obs[:, :3], states[:, :3]), dist3d(obs[:, 3:6], states[:, 3:]), var3d(obs[:, 6:9]), var3d(obs[:, 9:])


def reward_proportional_trinary_true(delta_pos):  # results.py:431-433
    return np.mean(((delta_pos < 1e4) * 1 + (delta_pos < 1e7) * 1)) / 2


def score_scaled_trace_P(P, dt=None):  # reward.py:6-14
    diag = np.diag(P)
    return np.sqrt(np.sum(diag[:3])) + np.sqrt(np.sum(diag[3:])) * 30


def score_trace_P(P):  # reward.py:17-19
    return np.trace(P)


def score_scaled_det_P(P, dt=30.0):  # reward.py:30-32
    return np.power(np.multiply(np.linalg.det(P), dt ** 6), 1 / 12)


def score_det_P(P, dt=30.0):  # reward.py:35-37
    return np.linalg.det(P)


def score_det_pos_P(P):  # reward.py:40-42
    return np.linalg.det(P[:3, :3])
