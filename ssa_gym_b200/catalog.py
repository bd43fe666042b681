"""Synthetic RSO catalogs (GCRS state vectors [m, m/s]) of the named sizes.

The reference ships a 20 000 x 6 fixture (`envs/1.5_hour_viz_20000_of_20000_sample_orbits_seed_0.npy`,
produced by envs/orbit_gen.py from the orbit classes of envs/dynamics.py:357-399).  That file is
reference data and is not copied; catalogs of the same class mix (LEO / MEO / GEO / Tundra / Molniya,
same element ranges: dynamics.py:365-396) are generated here, deterministically from a seed.  A user
can pass the reference's own array through `env_config['orbits']` exactly as before.

`coe2rv` is a float64 numpy restatement of the classical-elements -> state conversion
(envs/farnocchia.py:100-161); it only creates inputs and is not on the graded path.
"""
import numpy as np

MU = 398600441800000.0
RE_EQ = 6378136.6  # poliastro Earth.R [m] (dynamics.py:17)


def coe2rv(p, ecc, inc, raan, argp, nu, k=MU):
    cnu, snu = np.cos(nu), np.sin(nu)
    rp = p / (1 + ecc * cnu)
    vp = np.sqrt(k / p)
    rx, ry = cnu * rp, snu * rp
    vx, vy = -snu * vp, (ecc + cnu) * vp
    cO, sO, ci, si, cw, sw = np.cos(raan), np.sin(raan), np.cos(inc), np.sin(inc), np.cos(argp), np.sin(argp)
    a00, a01 = cO * cw - sO * ci * sw, -cO * sw - sO * ci * cw
    a10, a11 = sO * cw + cO * ci * sw, -sO * sw + cO * ci * cw
    a20, a21 = si * sw, si * cw
    return np.stack([rx * a00 + ry * a01, rx * a10 + ry * a11, rx * a20 + ry * a21,
                     vx * a00 + vy * a01, vx * a10 + vy * a11, vx * a20 + vy * a21], axis=-1)


def synthetic_catalog(n=20000, seed=0):
    """n orbits drawn from the reference's class mix ['LEO','MEO','GEO','LEO','MEO','GEO','Tundra','Molniya']."""
    rng = np.random.RandomState(seed)
    classes = np.array([0, 1, 2, 0, 1, 2, 3, 4])[rng.randint(0, 8, size=n)]
    inc = np.radians(rng.uniform(0, 180, n))
    raan = np.radians(rng.uniform(0, 360, n))
    argp = np.radians(rng.uniform(0, 360, n))
    nu = np.radians(rng.uniform(0, 360, n))
    a = np.empty(n)
    ecc = np.empty(n)
    for cls, lo, hi in ((0, RE_EQ + 300e3, RE_EQ + 2000e3), (1, RE_EQ + 2000e3, RE_EQ + 35786e3)):
        idx = np.where(classes == cls)[0]
        todo = idx
        while len(todo):  # rejection: semi-minor axis above 300 km altitude (dynamics.py:370-383)
            a[todo] = rng.uniform(lo, hi, len(todo))
            ecc[todo] = rng.uniform(0, .25, len(todo))
            b = a[todo] * np.sqrt(1 - ecc[todo] ** 2)
            todo = todo[b <= RE_EQ + 300e3]
    geo = classes == 2
    stationary = rng.randint(0, 2, n)
    a[geo] = 42164e3
    ecc[geo] = (stationary * rng.uniform(0, .25, n))[geo]
    inc[geo] = 0.0  # dynamics.py:388: uniform(0, radians(0))
    tun = classes == 3
    a[tun], inc[tun], ecc[tun], argp[tun] = 42164e3, np.radians(63.4), 0.2, np.radians(270)
    mol = classes == 4
    a[mol], inc[mol], ecc[mol], argp[mol] = 26600e3, np.radians(63.4), 0.737, np.radians(270)
    p = a * (1 - ecc ** 2)
    return np.ascontiguousarray(coe2rv(p, ecc, inc, raan, argp, nu))


def tiled_catalog(n, base=None, seed=2):
    """tile a base catalog to n objects with N(0, [1e3 m]*3 + [1 m/s]*3) jitter so that
    no two objects coincide."""
    if base is None:
        base = synthetic_catalog(20000, 0)
    reps = -(-n // len(base))
    out = np.tile(base, (reps, 1))[:n].copy()
    rng = np.random.RandomState(seed)
    out += rng.normal(size=out.shape) * np.array([1e3] * 3 + [1.0] * 3)
    return out
