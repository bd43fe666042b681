"""Innovation whiteness statistics (SS2:644-698 autocorrelation, SS2:782-832 Durbin-Watson).

CPU: the numpy restatement (oracle/diagnostics.py) against closed forms and hand-computed values of the published
statsmodels algorithms (statsmodels is not in the image).  -m gpu: ssa_innovation_stats through the C ABI against that
restatement, and the env's innovation_dw_test / autocorrelation over a played episode."""
import numpy as np
import pytest

import helpers as H
from oracle import diagnostics as OD


def test_durbin_watson_restatement_known_values():
    e = np.array([[1.0, 2.0], [2.0, 0.0], [4.0, -2.0], [3.0, 1.0]])
    # column 0: diffs 1, 2, -1 -> 6 / (1 + 4 + 16 + 9) = 0.2; column 1: diffs -2, -2, 3 -> 17 / 9
    assert np.allclose(OD.durbin_watson(e), [6 / 30, 17 / 9], rtol=1e-15)
    rng = np.random.RandomState(0)
    w = rng.normal(size=(20000, 3))
    assert np.all(np.abs(OD.durbin_watson(w) - 2.0) < 0.05)                 # white noise: 2 (1 - r), r = 0
    assert np.all(OD.durbin_watson(np.cumsum(w, axis=0)) < 0.05)            # a random walk is strongly correlated


def test_acf_conservative_restatement():
    x = np.array([1.0, 3.0, 2.0, 5.0, 4.0])
    xo = x - x.mean()
    want = np.array([np.sum(xo[:5 - k] * xo[k:]) for k in range(4)]) / np.sum(xo * xo)
    assert np.allclose(OD.acf_conservative(x, 3), want, rtol=1e-15)
    # missing='conservative': demean with the mean of the valid entries, zeros at the NaNs, normalisation by acov[0]
    xm = np.array([1.0, np.nan, 2.0, 5.0, np.nan, 4.0])
    ok = ~np.isnan(xm)
    z = np.where(ok, xm - np.nanmean(xm), 0.0)
    want = np.array([np.sum(z[:6 - k] * z[k:]) for k in range(4)]) / np.sum(z * z)
    assert np.allclose(OD.acf_conservative(xm, 3), want, rtol=1e-15)
    assert OD.acf_conservative(np.random.RandomState(1).normal(size=4000), 40)[0] == 1.0


@pytest.mark.gpu
def test_device_innovation_stats_against_restatement():
    from ssa_gym_b200.ukf import innovation_stats
    rng = np.random.RandomState(3)
    B, n, nlags = 37, 479, 40
    y = rng.normal(size=(B, n, 3)) * np.array([1e-5, 1e-5, 1e3])
    y[5] = np.cumsum(y[5], axis=0)                                           # a correlated series
    valid = rng.uniform(size=(B, n)) < 0.7
    valid[0] = True
    valid[1, 3:] = False                                                     # three observations only
    yn = np.where(valid[:, :, None], y, np.nan)
    dw, acf = innovation_stats(yn, valid, nlags)
    for b in range(B):
        e = y[b][valid[b]]
        assert np.allclose(dw[b], OD.durbin_watson(e), rtol=1e-12), b
        for c in range(3):
            assert np.allclose(acf[b, c], OD.acf_conservative(yn[b, :, c], nlags), rtol=1e-10, atol=1e-13), (b, c)
    assert np.all(np.abs(dw[0] - 2.0) < 0.3) and np.all(dw[5] < 0.2)


@pytest.mark.gpu
def test_env_durbin_watson_and_autocorrelation():
    import ssa_gym_b200
    from ssa_gym_b200 import agents, env as envmod
    cfg = dict(ssa_gym_b200.env_config, steps=120, rso_count=6, reward_type="trinary", obs_limit=-90)
    env = envmod.SSA_Tasker_Env(cfg)
    env.seed(3)
    obs, done = env.reset(), False
    while not done:
        obs, _, done, _ = env.step(int(agents.agent_visible_greedy(obs, env)))
    innovation, innovations = env.innovation()
    assert innovation.shape == (119, 3) and len(innovations) == 6
    table = np.asarray(env.innovation_dw_test())
    ok = ~np.isnan(innovation).any(axis=1)
    assert ok.sum() > 30
    assert np.allclose(table[0], np.round(OD.durbin_watson(innovation[ok]), 3), atol=1.01e-3)
    per = [OD.durbin_watson(inn[~np.isnan(inn).any(axis=1)]) for inn in innovations if (~np.isnan(inn).any(axis=1)).sum() >= 2]
    assert np.allclose(table[1], np.round(np.min(per, axis=0), 3), atol=1.01e-3) and np.allclose(table[2], np.round(np.max(per, axis=0), 3), atol=1.01e-3)
    ac, acs = env.autocorrelation(nlags=20)
    assert ac.shape == (3, 21) and len(acs) == 6 and np.all(ac[:, 0] == 1.0)
    for c in range(3):
        assert np.allclose(ac[c], OD.acf_conservative(innovation[:, c], 20), rtol=1e-10, atol=1e-13)
    env.close()
