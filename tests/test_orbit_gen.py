"""Catalog generator (envs/orbit_gen.py).  CPU: the oracle restatement and the product's arithmetic
(host twin) against the golden fixture built from the reference's own fx / ecef2lla / ecef2aer
(tests/golden/make_golden_orbit_gen.py); -m gpu: the kernels against the twin bit for bit, and an end-to-end
generate_catalog run whose every orbit passes the reference rule."""
import ctypes
import os

import numpy as np
import pytest

import helpers as H
from oracle import orbit_gen_oracle as OG

G = np.load(os.path.join(H.GOLDEN, "golden_orbit_gen.npz"))
LIMIT = np.radians(15.0)


def _T():
    lla = G["lla"]
    from ssa_gym_b200.transformations import trans_uvw_ecef
    return np.ascontiguousarray(trans_uvw_ecef(lla[0], lla[1]), dtype=np.float64).reshape(9)


def _twin_eval(cand, table, step):
    K, n = len(cand), len(table)
    acc = np.zeros(K, np.uint8); el = np.zeros((K, n)); alt = np.zeros((K, n))
    cand = np.ascontiguousarray(cand); tab = np.ascontiguousarray(table).reshape(n, 9)
    H.twin().twin_orbit_gen_eval(H.p(cand), K, H.p(tab), n, ctypes.c_double(step), H.p(np.ascontiguousarray(G["obs_itrs"])),
                                 H.p(_T()), ctypes.c_double(LIMIT), ctypes.c_double(300e3), 18, 36, H.p(acc), H.p(el), H.p(alt))
    return acc.astype(bool), el, alt


def _decisive(el, alt):
    """candidates whose decision does not hang on a sample within 1e-9 of a threshold"""
    return (np.abs(el - LIMIT).min(1) > 1e-9) & (np.abs(alt - 300e3).min(1) > 1e-3)


def test_oracle_matches_reference_built_golden():
    fx = lambda s, dt: H.lib_fx("oracle", s, dt)[0]
    hx = lambda s, M: H.lib_hx("oracle", s, M, G["obs_itrs"], _T())
    acc, el, alt = OG.evaluate(G["candidates"], G["table"], float(G["step"]), G["obs_itrs"], _T(), LIMIT, fx, hx)
    assert np.max(np.abs(el - G["elevation"])) < 1e-10
    assert np.max(np.abs(alt - G["altitude"])) < 1e-5          # metres, on 1e6..4e7 m
    assert np.array_equal(acc, G["accept"]) and 20 < acc.sum() < 200


def test_product_arithmetic_matches_reference_built_golden():
    acc, el, alt = _twin_eval(G["candidates"], G["table"], float(G["step"]))
    assert np.max(np.abs(el - G["elevation"])) < 1e-10
    assert np.max(np.abs(alt - G["altitude"])) < 1e-5
    d = _decisive(G["elevation"], G["altitude"])
    assert d.mean() > 0.95 and np.array_equal(acc[d], G["accept"][d])   # integer work: exact wherever it is decidable
    assert np.array_equal(acc, G["accept"])


def test_gap_rule_edge_cases():
    """orbit_gen.py:60-73: always visible; seen late (first window empty); gap of exactly max_gap; below 300 km once."""
    n, step = 96, 150.0
    hi = np.full(n, 5e5)
    vis = lambda mask: np.where(mask, 0.5, 0.0)   # elevation above / below the 15 degree limit
    always = np.ones(n, bool)
    late = always.copy(); late[:18] = False                      # nothing in the first 45 minutes
    gap35 = always.copy(); gap35[30:65] = False                  # 35 samples < 36
    gap36 = always.copy(); gap36[30:66] = False                  # 36 samples: not < max_gap
    assert OG.accept_rule(hi, vis(always), LIMIT, step)
    assert not OG.accept_rule(hi, vis(late), LIMIT, step)
    assert OG.accept_rule(hi, vis(gap35), LIMIT, step)
    assert not OG.accept_rule(hi, vis(gap36), LIMIT, step)
    low = hi.copy(); low[40] = 299e3
    assert not OG.accept_rule(low, vis(always), LIMIT, step)


def test_sampler_distribution():
    from ssa_gym_b200.orbit_gen import sample_candidates
    from ssa_gym_b200.catalog import RE_EQ
    c = sample_candidates(60000, np.random.RandomState(3))
    mu = 398600441800000.0
    r = np.linalg.norm(c[:, :3], axis=1); v2 = (c[:, 3:] ** 2).sum(1)
    a = 1.0 / (2.0 / r - v2 / mu)
    h = np.cross(c[:, :3], c[:, 3:])
    e = np.linalg.norm(np.cross(c[:, 3:], h) / mu - c[:, :3] / r[:, None], axis=1)
    geo_like = np.abs(a - 42164e3) < 1.0
    mol = np.abs(a - 26600e3) < 1.0
    assert abs(geo_like.mean() - 2 / 9) < 0.01 and abs(mol.mean() - 1 / 9) < 0.01      # GEO + Tundra share a
    leo = (a < RE_EQ + 2000e3 + 1) & ~geo_like & ~mol
    assert abs(leo.mean() - 1 / 3) < 0.01
    assert (a[leo] * np.sqrt(1 - e[leo] ** 2) > RE_EQ + 300e3 - 1).all() and e[leo].max() < 0.25 + 1e-9
    assert np.allclose(e[mol], 0.737, atol=1e-9)


@pytest.mark.gpu
def test_gpu_kernels_equal_twin_and_golden():
    from ssa_gym_b200 import orbit_gen
    acc_g, el_g, alt_g = orbit_gen.evaluate(G["candidates"], G["table"], float(G["step"]), G["lla"], LIMIT, details=True)
    acc_t, el_t, alt_t = _twin_eval(G["candidates"], G["table"], float(G["step"]))
    assert H.bits_equal(el_g, el_t) and H.bits_equal(alt_g, alt_t) and np.array_equal(acc_g, acc_t)
    assert np.array_equal(acc_g, G["accept"])


@pytest.mark.gpu
def test_generate_catalog_end_to_end():
    from datetime import datetime
    from ssa_gym_b200 import orbit_gen
    from ssa_gym_b200.transformations import gcrs2irts_matrix_approx, time_table
    table = gcrs2irts_matrix_approx(time_table(datetime(2020, 5, 4), 150.0, 96))
    cat, rate = orbit_gen.generate_catalog(samples=3000, seed=1, trans_matrix=table, batch=8192)
    assert cat.shape == (3000, 6) and 0.0 < rate < 0.5      # (proposals concentrate on the regimes that are hard to accept)
    acc, el, alt = _twin_eval(cat, table, 150.0)
    assert acc.all()                                              # every orbit of the catalog passes the rule
    assert (alt > 300e3).all() and ((el >= LIMIT).sum(1) > 0).all()
    # the ACCEPTED catalog has the reference's regime mix (orbit_gen.py:51-54 fixes the regime per output slot and
    # retries inside it): classify the orbits back from their elements and compare with the slots' drawn regimes
    mu = 398600441800000.0
    r, v = np.linalg.norm(cat[:, :3], axis=1), np.linalg.norm(cat[:, 3:], axis=1)
    a = 1.0 / (2.0 / r - v * v / mu)
    h = np.cross(cat[:, :3], cat[:, 3:])
    ecc = np.sqrt(np.maximum(0.0, 1.0 - np.sum(h * h, 1) / (mu * a)))
    inc = np.degrees(np.arccos(h[:, 2] / np.linalg.norm(h, axis=1)))
    geo_a = np.abs(a - 42164e3) < 1.0
    cls = np.where(np.abs(ecc - 0.737) < 1e-6, 4, np.where(geo_a & (np.abs(inc - 63.4) < 1e-6) & (np.abs(ecc - 0.2) < 1e-6), 3,
                   np.where(geo_a & (inc < 1e-9), 2, np.where(a < 6378136.6 + 2000e3, 0, 1))))
    assert np.array_equal(cls, orbit_gen.generate_catalog.last_slot_regime)
    frac = np.bincount(cls, minlength=5) / len(cls)
    assert np.all(np.abs(frac - np.array(orbit_gen.REGIME_P)) < 0.03), frac
    # fractional gap limit: integer gap lengths against a float limit == against its ceiling (orbit_gen.py:66)
    ok_a = orbit_gen.evaluate(G["candidates"], G["table"], float(G["step"]), G["lla"], LIMIT, max_gap=35.2)
    ok_b = orbit_gen.evaluate(G["candidates"], G["table"], float(G["step"]), G["lla"], LIMIT, max_gap=36)
    assert np.array_equal(ok_a, ok_b)
