"""TEST INFRASTRUCTURE — restatement of the acceptance rule of the reference's catalog generator
(envs/orbit_gen.py:47-75) on top of the C oracle's fx / hx, the checker of `ssa_orbit_gen_eval`.

    x_gcrs[i]  = fx(candidate, step*i)                         orbit_gen.py:56 (fx_xyz_markley upstream: the same two-body
                                                                flow; poliastro is absent, the oracle's Farnocchia fx is used)
    x_itrs[i]  = x_gcrs[i][:3] @ trans_matrix[i]               orbit_gen.py:57
    alt[i]     = ecef2lla(x_itrs[i])[2]                        orbit_gen.py:58, transformations.py:239-279
    el[i]      = hx_aer_erfa(x_gcrs[i], trans_matrix[i], ...)  orbit_gen.py:59
    rule                                                        orbit_gen.py:60-73
"""
import itertools

import numpy as np

A = 6378137.0
F = 1.0 / 298.257223563
B = (1 - F) * A


def ecef_altitude(ecef):
    """transformations.py:239-279 (You 2000), altitude only."""
    x, y, z = ecef
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    E = np.sqrt(A ** 2 - B ** 2)
    u = np.sqrt(0.5 * (r ** 2 - E ** 2) + 0.5 * np.sqrt((r ** 2 - E ** 2) ** 2 + 4 * E ** 2 * z ** 2))
    Q = np.hypot(x, y)
    huE = np.hypot(u, E)
    if not (Q == 0 or u == 0):
        beta = np.arctan(huE / u * z / Q)
    else:
        beta = np.pi / 2 if z >= 0 else -np.pi / 2
    eps = ((B * u - A * huE + E ** 2) * np.sin(beta)) / (A * huE * 1 / np.cos(beta) - E ** 2 * np.cos(beta))
    beta += eps
    alt = np.hypot(z - B * np.sin(beta), Q - A * np.cos(beta))
    if x ** 2 / A ** 2 + y ** 2 / A ** 2 + z ** 2 / B ** 2 < 1:
        alt = -alt
    return alt


def accept_rule(altitude, elevation, obs_limit, step_size, max_gap_hours=1.5, first_window_min=45):
    """orbit_gen.py:60-73 for one candidate (arrays over the sample times)."""
    visibility = elevation >= obs_limit
    gaps = [sum(1 for _ in group) for key, group in itertools.groupby((visibility - 1) * -1) if key]
    if np.all(altitude > 300 * 1000):
        if not gaps == []:
            if sum(visibility[0:int(first_window_min * 60 / step_size)]) > 0:
                if np.max(gaps) < max_gap_hours * 60 * 60 / step_size:
                    return True
        else:
            if np.all(visibility):
                return True
    return False


def evaluate(candidates, trans_table, step_size, obs_itrs, T, obs_limit, fx, hx):
    """fx(states [K,6], dt) -> [K,6]; hx(states [K,6], M) -> [K,3].  Returns accept [K], el [K,n], alt [K,n]."""
    K, n = len(candidates), len(trans_table)
    el = np.zeros((K, n)); alt = np.zeros((K, n))
    for i in range(n):
        x = fx(candidates, step_size * i)
        M = trans_table[i].reshape(3, 3)
        el[:, i] = hx(x, M)[:, 1]
        xi = x[:, :3] @ M
        alt[:, i] = [ecef_altitude(v) for v in xi]
    acc = np.array([accept_rule(alt[c], el[c], obs_limit, step_size) for c in range(K)])
    return acc, el, alt
