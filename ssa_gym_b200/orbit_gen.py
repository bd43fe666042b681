"""Catalog generator: envs/orbit_gen.py as a batched device job.

The reference draws one candidate orbit at a time (`init_state_vec`, dynamics.py:357-399), propagates it through a
4-hour window at 150-second steps and keeps it if it stays above 300 km and the observer never loses it for 1.5 hours
(orbit_gen.py:47-75): 20 000 accepted orbits take hours of Python.  Here candidates are drawn in batches, and ONE
call of `ssa_orbit_gen_eval` (two kernels: propagation + altitude + elevation for every (candidate, time), then the
gap rule per candidate) decides the whole batch.

The acceptance rule is the reference's, decision for decision (tests: against the reference's own functions).  The
DISTRIBUTION of the accepted catalog is the reference's: the regime of every output slot is drawn first (1/3 LEO, 1/3
MEO, 1/9 GEO, 1/9 Tundra, 1/9 Molniya) and candidates are retried inside that regime until one is accepted
(orbit_gen.py:51-54), with the element ranges of dynamics.py:362-397 including the exo-atmospheric rejection loop.
The candidate STREAM is not the reference's: a sequential accept/reject loop on one RandomState cannot be batched, so
candidates are drawn vectorised.
"""
import ctypes

import numpy as np

from . import _lib
from .catalog import RE_EQ, coe2rv
from .transformations import default_eops, deg2rad, gcrs2irts_matrix_b, lla2ecef, load_eop_c04, time_table, trans_uvw_ecef

REGIMES = ('LEO', 'MEO', 'GEO', 'Tundra', 'Molniya')
REGIME_P = (1 / 3, 1 / 3, 1 / 9, 1 / 9, 1 / 9)          # orbit_gen.py:52


def sample_candidates(k, rng, p=REGIME_P, reg=None):
    """k candidate states [k, 6] (GCRS, m and m/s) with the distributions of dynamics.py:357-399; `reg` fixes the regime
    of every candidate (indices into REGIMES), else it is drawn with probabilities p."""
    reg = rng.choice(len(REGIMES), size=k, p=p) if reg is None else np.asarray(reg, dtype=int)
    inc = np.radians(rng.uniform(0, 180, k))
    raan = np.radians(rng.uniform(0, 360, k))
    argp = np.radians(rng.uniform(0, 360, k))
    nu = np.radians(rng.uniform(0, 360, k))
    a = np.empty(k)
    ecc = np.empty(k)
    for cls, lo, hi in ((0, RE_EQ + 300e3, RE_EQ + 2000e3), (1, RE_EQ + 2000e3, RE_EQ + 35786e3)):
        todo = np.where(reg == cls)[0]
        while len(todo):  # exo-atmospheric rejection: semi-minor axis above 300 km (dynamics.py:370-383)
            a[todo] = rng.uniform(lo, hi, len(todo))
            ecc[todo] = rng.uniform(0, .25, len(todo))
            todo = todo[a[todo] * np.sqrt(1 - ecc[todo] ** 2) <= RE_EQ + 300e3]
    geo = reg == 2
    stationary = rng.randint(0, 2, k)
    a[geo] = 42164e3
    ecc[geo] = (stationary * rng.uniform(0, .25, k))[geo]
    inc[geo] = 0.0                                         # dynamics.py:388: stationary * uniform(0, radians(0))
    tun = reg == 3
    a[tun], inc[tun], ecc[tun], argp[tun] = 42164e3, np.radians(63.4), 0.2, np.radians(270)
    mol = reg == 4
    a[mol], inc[mol], ecc[mol], argp[mol] = 26600e3, np.radians(63.4), 0.737, np.radians(270)
    return np.ascontiguousarray(coe2rv(a * (1 - ecc ** 2), ecc, inc, raan, argp, nu))


def evaluate(candidates, trans_table, step_s, obs_lla, obs_limit_rad, min_alt=300e3, first_window=18, max_gap=36,
             details=False, device=0):
    """Acceptance flags (bool [K]) of the candidates; with details=True also elevation and altitude [K, n].
    `max_gap` (in samples) may be fractional: the reference compares the integer gap lengths with the float limit
    (orbit_gen.py:66), which for integers is the comparison with its ceiling."""
    cand = np.ascontiguousarray(candidates, dtype=np.float64).reshape(-1, 6)
    table = np.ascontiguousarray(trans_table, dtype=np.float64).reshape(-1, 9)
    K, n = len(cand), len(table)
    obs_lla = np.asarray(obs_lla, dtype=np.float64)
    obs_itrs = np.ascontiguousarray(lla2ecef(obs_lla))
    T = np.ascontiguousarray(trans_uvw_ecef(obs_lla[0], obs_lla[1]), dtype=np.float64).reshape(9)
    acc = np.zeros(K, dtype=np.uint8)
    el = np.zeros((K, n)) if details else None
    alt = np.zeros((K, n)) if details else None
    vp = lambda x_: None if x_ is None else x_.ctypes.data_as(ctypes.c_void_p)
    lib = _lib.require_gpu()
    _lib.check(lib.ssa_orbit_gen_eval(vp(cand), K, vp(table), n, float(step_s), vp(obs_itrs), vp(T), float(obs_limit_rad),
                                      float(min_alt), int(first_window), int(np.ceil(max_gap)), vp(acc), vp(el), vp(alt), int(device)),
               "ssa_orbit_gen_eval")
    return (acc.astype(bool), el, alt) if details else acc.astype(bool)


def generate_catalog(samples=20000, seed=0, step_size=60 * 2.5, max_gap_hours=1.5, duration_hours=4, obs_limit_deg=15,
                     first_window_min=45, observer=(38.828198, -77.305352, 20.0), t_0=None, trans_matrix=None, eop_file=None,
                     batch=16384, device=0):
    """`samples` orbits visible from `observer` at least every `max_gap_hours` (the parameters and their defaults are
    the module constants of envs/orbit_gen.py:28-43).  Returns (catalog [samples, 6], acceptance rate)."""
    from datetime import datetime
    n = int(np.ceil(duration_hours * 60 * 60 / step_size))
    if trans_matrix is None:
        eops = load_eop_c04(eop_file) if eop_file else default_eops()
        trans_matrix = gcrs2irts_matrix_b(time_table(t_0 or datetime(2020, 5, 4, 0, 0, 0), step_size, n), eops)
    obs_lla = np.array(observer) * [deg2rad, deg2rad, 1]
    rng = np.random.RandomState(seed)
    # The reference fixes the regime of every OUTPUT slot first and retries inside that regime until a candidate is
    # accepted (orbit_gen.py:51-54), so the accepted catalog has exactly the drawn mix (1/3, 1/3, 1/9, 1/9, 1/9) whatever
    # the regimes' acceptance rates are.  Same here: the slots' regimes are drawn once, every batch proposes candidates
    # for the still-empty slots (cyclically, so that a batch is always full) and a slot takes its first accepted one.
    slot_regime = rng.choice(len(REGIMES), size=samples, p=REGIME_P)
    catalog = np.zeros((samples, 6))
    filled = np.zeros(samples, dtype=bool)
    drawn = accepted = 0
    while not filled.all():
        todo = np.where(~filled)[0]
        slots = todo[np.arange(batch) % len(todo)]
        cand = sample_candidates(batch, rng, reg=slot_regime[slots])
        ok = evaluate(cand, trans_matrix, step_size, obs_lla, np.radians(obs_limit_deg), first_window=int(first_window_min * 60 / step_size),
                      max_gap=max_gap_hours * 60 * 60 / step_size, device=device)
        drawn += len(cand)
        accepted += int(ok.sum())
        first = {}
        for k_ in np.where(ok)[0]:
            first.setdefault(int(slots[k_]), int(k_))
        idx = np.array(sorted(first), dtype=int)
        if len(idx):
            catalog[idx] = cand[[first[i] for i in idx]]
            filled[idx] = True
    generate_catalog.last_slot_regime = slot_regime
    return catalog, accepted / drawn
