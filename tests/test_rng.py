"""CPU: the counter-based generator of the device-resident episodic mode (csrc/ssa_rng.h, through the host twin).
Philox4x32-10 is pinned to the published known-answer vectors of Random123 (Salmon et al., SC'11, kat_vectors);
the normal / uniform transforms are checked statistically."""
import ctypes

import numpy as np
from scipy import stats

import helpers as H


def _philox(ctr, key):
    c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, np.uint32)
    H.twin().twin_philox(H.p(c), H.p(k), H.p(o))
    return [int(v) for v in o]


def test_philox4x32_10_known_answers():
    assert _philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normals_are_standard_normal():
    n = 600000
    for seed in (0, 1, 0xDEADBEEFCAFEF00D):
        out = np.zeros(n)
        H.twin().twin_normals(ctypes.c_uint64(seed), H.p(out), ctypes.c_int(n))
        assert np.isfinite(out).all()
        assert abs(out.mean()) < 5 / np.sqrt(n) and abs(out.var() - 1) < 5 * np.sqrt(2 / n)
        assert abs(stats.skew(out)) < 0.02 and abs(stats.kurtosis(out)) < 0.04
        assert stats.kstest(out[:100000], "norm").pvalue > 1e-3
        # the three components drawn for one (object, step) are uncorrelated
        c = np.corrcoef(out.reshape(-1, 3).T)
        assert np.abs(c - np.eye(3)).max() < 0.01
    a = np.zeros(3000); b = np.zeros(3000)
    H.twin().twin_normals(ctypes.c_uint64(7), H.p(a), ctypes.c_int(3000))
    H.twin().twin_normals(ctypes.c_uint64(8), H.p(b), ctypes.c_int(3000))
    assert not np.array_equal(a, b) and abs(np.corrcoef(a, b)[0, 1]) < 0.08   # different seeds: different streams


def test_reset_draws_distribution_and_determinism():
    """k_env_reset's draws (through the twin): catalog rows uniform, x_filter - x_true ~ N(0, x_sigma), P = P0,
    reproducible from (seed, episode), different across episodes and environments."""
    E, m, n_orb = 4000, 10, 37
    rng = np.random.RandomState(0)
    orbits = rng.normal(size=(n_orb, 6)) * 1e7
    x_sigma = np.array([1e3] * 3 + [10.0] * 3)
    P0 = np.diag(x_sigma ** 2)
    sig = np.concatenate([x_sigma, [1e-5, 2e-5, 50.0], P0[np.triu_indices(6)]])
    seeds = np.arange(E, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(12345)

    def draw(episode0, done=None):
        ep = np.full(E, episode0, np.uint32); si = np.full(E, 5, np.int32)
        xt = np.zeros((E * m, 6)); xf = np.zeros((E * m, 6)); P = np.zeros((E * m, 21))
        st = np.ones(E * m, np.int32); infl = np.ones(E * m, np.int32)
        H.twin().twin_env_reset(E, m, H.p(seeds), H.p(ep), H.p(si), H.p(done) if done is not None else None, H.p(orbits),
                                n_orb, H.p(sig), H.p(xt), H.p(xf), H.p(P), H.p(st), H.p(infl))
        return xt, xf, P, st, infl, ep, si

    xt, xf, P, st, infl, ep, si = draw(0)
    assert (ep == 1).all() and (si == 0).all() and (st == 0).all() and (infl == 0).all()
    assert np.array_equal(P, np.tile(P0[np.triu_indices(6)], (E * m, 1)))
    rows = np.array([np.where((orbits == r).all(1))[0][0] for r in xt])
    counts = np.bincount(rows, minlength=n_orb)
    assert stats.chisquare(counts).pvalue > 1e-3
    d = (xf - xt) / x_sigma
    assert abs(d.mean()) < 5 / np.sqrt(d.size) and abs(d.var() - 1) < 5 * np.sqrt(2 / d.size)
    assert stats.kstest(d.ravel()[:100000], "norm").pvalue > 1e-3
    xt2, xf2 = draw(0)[:2]
    assert np.array_equal(xt, xt2) and np.array_equal(xf, xf2)            # same (seed, episode): same draw
    xt3, xf3 = draw(1)[:2]
    assert not np.array_equal(xf, xf3)                                    # next episode: new draw
    done = np.zeros(E, np.uint8); done[::3] = 1
    _, xf4, _, st4, _, ep4, _ = draw(0, done)
    assert np.array_equal(ep4, np.where(done, 1, 0)) and (xf4.reshape(E, m, 6)[done == 0] == 0).all()


def test_step_noise_scaled_and_addressable():
    E, m = 500, 10
    seeds = np.arange(E, dtype=np.uint64) + np.uint64(99)
    sig = np.zeros(30); sig[6:9] = [4.8e-6, 4.8e-6, 1e3]
    ep = np.ones(E, np.uint32)
    out = {}
    for step in (0, 1, 7):
        si = np.full(E, step, np.int32); z = np.zeros((E * m, 3))
        H.twin().twin_env_noise(E, m, H.p(seeds), H.p(ep), H.p(si), H.p(sig), H.p(z))
        out[step] = z
        zn = z / sig[6:9]
        assert abs(zn.mean()) < 5 / np.sqrt(zn.size) and abs(zn.var() - 1) < 0.05
    assert not np.array_equal(out[0], out[1]) and not np.array_equal(out[1], out[7])
    si = np.full(E, 1, np.int32); z = np.zeros((E * m, 3))
    H.twin().twin_env_noise(E, m, H.p(seeds), H.p(ep), H.p(si), H.p(sig), H.p(z))
    assert np.array_equal(z, out[1])
