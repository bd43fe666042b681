#!/usr/bin/env python
"""High-precision ground truth for the UKF predict + update of the 24-object golden case (tests/golden/
golden_ukf_aer_resample.npz): the SAME algorithm (filterpy's predict / update as the reference drives it, SURVEY Appendix
B; sigma weights, Q, R, observer constants as the IEEE doubles the reference computes) evaluated with mpmath at 60
significant digits, so that rounding plays no role.  The only liberty: fx is the exact two-body flow (universal-variable
Kepler solution converged to working precision) — the map the reference's farnocchia() approximates to 1e-16.

    python tests/golden/make_exact.py        (pure mpmath + numpy; does not need /root/reference)

Writes tests/golden/golden_ukf_exact.npz: x_pred, P_pred, x, P, y, S per step (exact values rounded to double).
tests/test_exact_truth.py compares |implementation - exact| with |reference-built golden - exact|.
"""
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 60
MU = mp.mpf(398600441800000.0)


def stumpff(z):
    if abs(z) < mp.mpf(10) ** -20:
        return mp.mpf(1) / 2 - z / 24, mp.mpf(1) / 6 - z / 120
    if z > 0:
        s = mp.sqrt(z)
        return (1 - mp.cos(s)) / z, (s - mp.sin(s)) / (s * z)
    s = mp.sqrt(-z)
    return (mp.cosh(s) - 1) / (-z), (mp.sinh(s) - s) / (s * (-z))


def fx_exact(x, dt):
    r0v, v0v = x[:3], x[3:]
    r0 = mp.sqrt(sum(c * c for c in r0v))
    v2 = sum(c * c for c in v0v)
    rv = sum(a * b for a, b in zip(r0v, v0v))
    alpha = 2 / r0 - v2 / MU
    sm = mp.sqrt(MU)
    chi = sm * abs(alpha) * dt
    for _ in range(200):
        z = alpha * chi * chi
        C, S = stumpff(z)
        F = rv / sm * chi * chi * C + (1 - alpha * r0) * chi ** 3 * S + r0 * chi - sm * dt
        dF = rv / sm * chi * (1 - z * S) + (1 - alpha * r0) * chi * chi * C + r0
        step = F / dF
        chi -= step
        if abs(step) < mp.mpf(10) ** -50 * max(1, abs(chi)):
            break
    z = alpha * chi * chi
    C, S = stumpff(z)
    f = 1 - chi * chi / r0 * C
    g = dt - chi ** 3 / sm * S
    r = [f * a + g * b for a, b in zip(r0v, v0v)]
    rn = mp.sqrt(sum(c * c for c in r))
    fd = sm / (rn * r0) * chi * (z * S - 1)
    gd = 1 - chi * chi / rn * C
    v = [fd * a + gd * b for a, b in zip(r0v, v0v)]
    return r + v


def chol_upper(A):
    n = len(A)
    U = [[mp.mpf(0)] * n for _ in range(n)]
    for j in range(n):
        d = A[j][j] - sum(U[k][j] ** 2 for k in range(j))
        U[j][j] = mp.sqrt(d)
        for c in range(j + 1, n):
            U[j][c] = (A[j][c] - sum(U[k][j] * U[k][c] for k in range(j))) / U[j][j]
    return U


def sigma_points(x, P, lam):
    U = chol_upper([[lam * P[i][j] for j in range(6)] for i in range(6)])
    pts = [list(x)]
    pts += [[x[j] + U[k][j] for j in range(6)] for k in range(6)]
    pts += [[x[j] - U[k][j] for j in range(6)] for k in range(6)]
    return pts


def hx_exact(x, M, obs_itrs, T):
    xi = [sum(M[i][j] * x[j] for j in range(3)) for i in range(3)]
    d = [xi[i] - obs_itrs[i] for i in range(3)]
    e = [sum(T[j][i] * d[j] for j in range(3)) for i in range(3)]  # T^T d
    r = mp.sqrt(sum(c * c for c in d))
    az = mp.atan2(e[1], e[0])
    if az < 0:
        az += 2 * mp.pi
    return [az, mp.asin(e[2] / r), r]


def aer2uvw(a):
    return [a[2] * mp.cos(a[1]) * mp.cos(a[0]), a[2] * mp.cos(a[1]) * mp.sin(a[0]), a[2] * mp.sin(a[1])]


def uvw2aer(u):
    r = mp.sqrt(sum(c * c for c in u))
    az = mp.atan2(u[1], u[0])
    if az < 0:
        az += 2 * mp.pi
    return [az, mp.asin(u[2] / r), r]


def residual(a, b):
    d = a[0] - b[0]
    return [mp.atan2(mp.sin(d), mp.cos(d)), a[1] - b[1], a[2] - b[2]]


def inv3(S):
    return (mp.matrix(S) ** -1).tolist()


def to_mp(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return [mp.mpf(float(v)) for v in a]
    return [[mp.mpf(float(v)) for v in row] for row in a]


def main():
    import helpers as H
    g = np.load(os.path.join(OUT, "golden_ukf_aer_resample.npz"))
    geo = np.load(os.path.join(OUT, "golden_geometry.npz"))
    cfg = H.make_cfg(8)
    from ssa_gym_b200.ukf import Q_discrete_white_noise_block, merwe_weights
    Wm, Wc, lam = merwe_weights(6, float(g["alpha"]), 2.0, -3.0)
    Wm, Wc, lam = to_mp(Wm), to_mp(Wc), mp.mpf(float(lam))
    dt = mp.mpf(float(g["dt"]))
    Q = to_mp(Q_discrete_white_noise_block(float(g["dt"]), float(g["q_sigma"]) ** 2))
    R = to_mp(g["R"])
    M = to_mp(geo["M"])
    obs_itrs = to_mp(np.array(cfg.obs_itrs))
    T = to_mp(np.array(cfg.T).reshape(3, 3))
    assert np.array_equal(np.array(cfg.obs_itrs), geo["obs_itrs"])
    n_obj, steps = g["x0"].shape[0], g["z_noise"].shape[0]
    rec = {k: np.zeros((steps, n_obj) + s) for k, s in (("x_true", (6,)), ("x_pred", (6,)), ("P_pred", (6, 6)), ("x", (6,)),
                                                         ("P", (6, 6)), ("y", (3,)), ("S", (3, 3)))}
    f64 = lambda v: np.array([[float(c) for c in row] for row in v]) if isinstance(v[0], list) else np.array([float(c) for c in v])
    for j in range(n_obj):
        xt, x, P = to_mp(g["x_true0"][j]), to_mp(g["x0"][j]), to_mp(g["P0"])
        for s in range(steps):
            xt = fx_exact(xt, dt)
            # predict
            F = [fx_exact(sp, dt) for sp in sigma_points(x, P, lam)]
            xb = [sum(Wm[k] * F[k][i] for k in range(13)) for i in range(6)]
            Pb = [[sum(Wc[k] * (F[k][a] - xb[a]) * (F[k][b] - xb[b]) for k in range(13)) + Q[a][b] for b in range(6)] for a in range(6)]
            # re-draw (filterpy >= 1.4.5) and update
            sig = sigma_points(xb, Pb, lam)
            Z = [hx_exact(sp, M, obs_itrs, T) for sp in sig]
            um = [sum(Wm[k] * aer2uvw(Z[k])[i] for k in range(13)) for i in range(3)]
            zp = uvw2aer(um)
            rz = [residual(Z[k], zp) for k in range(13)]
            S = [[sum(Wc[k] * rz[k][a] * rz[k][b] for k in range(13)) + R[a][b] for b in range(3)] for a in range(3)]
            Pxz = [[sum(Wc[k] * (sig[k][i] - xb[i]) * rz[k][a] for k in range(13)) for a in range(3)] for i in range(6)]
            SI = inv3(S)
            K = [[sum(Pxz[i][c] * SI[c][a] for c in range(3)) for a in range(3)] for i in range(6)]
            zt = hx_exact(xt, M, obs_itrs, T)
            z = [zt[a] + mp.mpf(float(g["z_noise"][s, j, a])) for a in range(3)]
            y = residual(z, zp)
            xn = [xb[i] + sum(K[i][a] * y[a] for a in range(3)) for i in range(6)]
            SKt = [[sum(S[a][c] * K[i][c] for c in range(3)) for i in range(6)] for a in range(3)]
            Pn = [[Pb[i][jj] - sum(K[i][a] * SKt[a][jj] for a in range(3)) for jj in range(6)] for i in range(6)]
            for k_, v in (("x_true", xt), ("x_pred", xb), ("P_pred", Pb), ("x", xn), ("P", Pn), ("y", y), ("S", S)):
                rec[k_][s, j] = f64(v)
            x, P = xn, Pn
        print("object", j, "done", flush=True)
    np.savez_compressed(os.path.join(OUT, "golden_ukf_exact.npz"), dps=mp.mp.dps, **rec)
    for k_ in ("x_true", "x_pred", "P_pred", "x", "P", "y", "S"):
        d = np.abs(g[k_] - rec[k_]) / (np.abs(rec[k_]) + 1e-300)
        print(k_, "golden vs exact: median rel", np.median(d), "max", d.max())


if __name__ == "__main__":
    main()
