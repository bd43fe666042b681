#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/*.npz from the REFERENCE's own functions.

Run here in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_golden.py

What is real reference code and what is restated:
  * fx                      = envs/farnocchia.py::fx_xyz_farnocchia, imported unmodified (numba)
  * rv2coe / coe2rv         = envs/farnocchia.py, imported unmodified
  * lla2ecef, ecef2aer, aer2uvw, uvw2aer, ecef2lla = envs/transformations.py, imported with the 6-line
                              astropy._erfa stub of oracle/ref_loader.py
  * hx_aer_erfa, residual_z_aer, mean_z_uvw, robust_cholesky = 5-line bodies of envs/dynamics.py restated in
                              oracle/dynamics_restated.py around the imported geometry (dynamics.py itself
                              needs astropy/poliastro/pymap3d at import time)
  * filterpy                = absent from the image; numpy restatement oracle/filterpy_restated.py
  * the environment         = oracle/env_oracle.py (restated ssa_tasker_simple_2.py hot path)
The fixtures are inputs + outputs; tests/test_oracle_golden.py checks the portable oracle (C) against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402
from oracle import dynamics_restated as D  # noqa: E402
from oracle import env_oracle as EO  # noqa: E402
from oracle.filterpy_restated import MerweScaledSigmaPoints, Q_discrete_white_noise, UnscentedKalmanFilter  # noqa: E402
from ssa_gym_b200.transformations import gcrs2irts_matrix_approx, time_table  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MU = 398600441800000.0
CEL2TER06AXY = np.array([[+0.973104317697536, +0.230363826239128, -0.000703163481769],
                         [-0.230363800456036, +0.973104570632801, +0.000118545368117],
                         [+0.000711560162594, +0.000046626402444, +0.999999745754024]])
X6 = np.array([34090858.3, 23944774.4, 6503066.82, -1983.785080, 2150.41744, 913.881611])
OBS_DEG = (38.828198, -77.305352, 20.0)


def main():
    assert rl.available(), "reference tree not found"
    far, tr, cat = rl.farnocchia(), rl.transformations(), rl.catalog()
    fx = far.fx_xyz_farnocchia
    rng = np.random.RandomState(12345)

    # ---- 1. fx / rv2coe -----------------------------------------------------------------------------
    ecc = np.array([far.rv2coe(MU, c[:3], c[3:])[1] for c in cat[:4000]])
    inc = np.array([far.rv2coe(MU, c[:3], c[3:])[2] for c in cat[:4000]])
    special = np.where((ecc < 1e-8) | (np.abs(inc) < 1e-8))[0][:40]
    molniya = np.where(ecc > 0.7)[0][:40]
    idx = np.unique(np.concatenate([np.arange(120), special, molniya, [18703 % 4000, 14746 % 4000]]))
    states = np.concatenate([cat[idx], cat[idx[:60]] + rng.normal(size=(60, 6)) * np.array([1e5] * 3 + [1e2] * 3), X6[None]])
    dts = np.array([20.0, 30.0, 600.0, 86400.0])
    fx_out = np.array([[fx(s, dt) for s in states] for dt in dts])
    coe = np.array([far.rv2coe(MU, s[:3], s[3:]) for s in states])
    np.savez_compressed(os.path.join(OUT, "golden_fx.npz"), states=states, dts=dts, fx_out=fx_out, coe=coe,
                        catalog_sample=cat[:512])

    # ---- 2. geometry ----------------------------------------------------------------------------------
    lla = np.array([np.radians(OBS_DEG[0]), np.radians(OBS_DEG[1]), OBS_DEG[2]])
    obs_itrs = tr.lla2ecef(lla)
    hx, residual, mean_uvw, _ = D.make_operators(tr)
    aer = np.array([hx(s, CEL2TER06AXY, lla, obs_itrs) for s in states])
    uvw = np.array([tr.aer2uvw(a) for a in aer])
    aer_back = np.array([tr.uvw2aer(u) for u in uvw])
    az = np.radians([0, 0.001, 90.0, 180, 270.0, 359.99, 360])
    el = np.radians([-90.00, -89.99, -0.999, 0, 0.999, 89.99, 90.00])
    sr = np.array([-1000.0001, -1, -0.0001, 0, 0.0001, 1, 1000.0001])
    from itertools import permutations
    ra = np.array([[a[0], e[0], s[0]] for a, e, s in zip(permutations(az, 2), permutations(el, 2), permutations(sr, 2))])
    rb = np.array([[a[1], e[1], s[1]] for a, e, s in zip(permutations(az, 2), permutations(el, 2), permutations(sr, 2))])
    rres = np.array([residual(a, b) for a, b in zip(ra, rb)])
    test3 = tr.ecef2aer(tr.ecef2lla(np.array([1285410., -4797210., 3994830.])), np.array([1202990., -4824940., 3999870.]),
                        np.array([1285410., -4797210., 3994830.]))
    mz = mean_uvw(np.array([[6.2, .1, 1e7], [.1, .12, 1.0001e7], [.05, .08, .9999e7]]), np.array([.5, .25, .25]))
    np.savez_compressed(os.path.join(OUT, "golden_geometry.npz"), lla=lla, obs_itrs=obs_itrs, M=CEL2TER06AXY, states=states,
                        aer=aer, uvw=uvw, aer_back=aer_back, res_a=ra, res_b=rb, res_out=rres, test3=test3, mean_z_known=mz)

    # ---- 3. UKF predict/update, catalog mode, default (AER) configuration ------------------------------
    def run_ukf(resample, obs_type, n_obj=24, steps=3, alpha=1e-4, dt=20.0, q_sigma=0.000025, R=None, P0=None):
        hx_aer, res_aer, mean_uvw2, hx_xyz = D.make_operators(tr)
        Q = Q_discrete_white_noise(dim=2, dt=dt, var=q_sigma ** 2, block_size=3, order_by_dim=False)
        if R is None:
            R = np.diag([D.arcsec2rad ** 2] * 2 + [1e3 ** 2])
        if P0 is None:
            P0 = np.diag([1e10] * 3 + [1e4] * 3)
        xt = cat[idx[:n_obj]].copy()
        xf = xt + np.random.RandomState(0).normal(size=(n_obj, 6)) * np.array([1e5] * 3 + [1e2] * 3)
        zn = np.random.RandomState(1).normal(size=(steps, n_obj, 3)) * (np.array([D.arcsec2rad, D.arcsec2rad, 1e3]) if obs_type == "aer" else 10.0)
        filters = []
        for j in range(n_obj):
            pts = MerweScaledSigmaPoints(6, alpha, 2.0, -3.0, sqrt_method=D.robust_cholesky)
            if obs_type == "aer":
                f = UnscentedKalmanFilter(6, 3, dt, hx_aer, fx, pts, sqrt_fn=D.robust_cholesky, z_mean_fn=mean_uvw2,
                                          residual_z=res_aer, resample_after_predict=resample)
            else:
                f = UnscentedKalmanFilter(6, 3, dt, hx_xyz, fx, pts, resample_after_predict=resample)
            f.x, f.P, f.Q, f.R = xf[j].copy(), P0.copy(), Q.copy(), np.array(R, dtype=float).copy()
            filters.append(f)
        rec = {k: [] for k in ("x_true", "x_pred", "P_pred", "x", "P", "y", "S", "sigmas_h")}
        kw = dict(trans_matrix=CEL2TER06AXY, observer_lla=lla, observer_itrs=obs_itrs)
        for s in range(steps):
            xt = np.array([fx(v, dt) for v in xt])
            xs, Ps, xp, Pp, ys, Ss, sh = [], [], [], [], [], [], []
            for j, f in enumerate(filters):
                f.predict()
                xp.append(f.x.copy()); Pp.append(f.P.copy())
                z = (hx_aer(xt[j], **kw) if obs_type == "aer" else xt[j][:3]) + zn[s, j]
                f.update(z, **kw)
                xs.append(f.x.copy()); Ps.append(f.P.copy()); ys.append(f.y.copy()); Ss.append(f.S.copy()); sh.append(f.sigmas_h.copy())
            for k, v in zip(("x_true", "x_pred", "P_pred", "x", "P", "y", "S", "sigmas_h"), (xt, xp, Pp, xs, Ps, ys, Ss, sh)):
                rec[k].append(np.array(v))
        out = {k: np.array(v) for k, v in rec.items()}
        out.update(x_true0=cat[idx[:n_obj]], x0=xf, P0=P0, z_noise=zn, R=np.array(R, dtype=float), alpha=alpha, dt=dt, q_sigma=q_sigma)
        return out

    for name, kw in (("aer_resample", dict(resample=True, obs_type="aer")), ("aer_noresample", dict(resample=False, obs_type="aer")),
                     ("xyz_resample", dict(resample=True, obs_type="xyz", R=np.diag([125.0] * 3)))):
        np.savez_compressed(os.path.join(OUT, f"golden_ukf_{name}.npz"), **run_ukf(**kw))

    # tests.py Test 6 / Test 7 scenario with the env's fx (envs/farnocchia.py): 50 predicts, update, 50 predicts, update
    # (sqrt_method = robust_cholesky instead of tests.py's plain cholesky: with alpha=1e-3 the posterior P of
    #  this scenario is indefinite at the 1e-3 level and plain cholesky raises or not depending on the last
    #  ulp of fx — measured here with three ulp-different fx builds; see DESIGN.md 'filterpy fork')
    pts = MerweScaledSigmaPoints(6, 0.001, 2.0, -3.0, sqrt_method=D.robust_cholesky)
    f = UnscentedKalmanFilter(6, 3, 30.0, D.make_operators(tr)[3], fx, pts, sqrt_fn=D.robust_cholesky)
    f.x, f.P = X6.copy(), np.eye(6) * np.array([1000, 1000, 1000, 1, 1, 1.0])
    f.Q = Q_discrete_white_noise(dim=2, dt=30.0, var=0.000001 ** 2, block_size=3, order_by_dim=False)
    xt = X6.copy()
    hist = []
    for rep in range(2):
        for i in range(50):
            f.predict(30.0)
            xt = fx(xt, 30.0)
        hist.append((f.x.copy(), f.P.copy(), xt.copy()))
        f.update(z=xt[:3], R=np.array([125, 125, 125]))
        hist.append((f.x.copy(), f.P.copy(), xt.copy()))
    np.savez_compressed(os.path.join(OUT, "golden_test6_7.npz"), x=np.array([h[0] for h in hist]), P=np.array([h[1] for h in hist]),
                        x_true=np.array([h[2] for h in hist]))

    # ---- 4. environment episodes (C1), heuristic agent ------------------------------------------------
    from datetime import datetime
    n_steps = 40
    base = {"steps": n_steps, "rso_count": 10, "time_step": 20., "t_0": datetime(2020, 5, 4, 0, 0, 0), "obs_limit": -90,
            "observer": OBS_DEG, "update_interval": 1, "obs_type": "aer", "z_sigma": (1, 1, 1e3),
            "x_sigma": tuple([1e5] * 3 + [1e2] * 3), "q_sigma": 0.000025, "P_0": np.diag(([1e5 ** 2] * 3 + [1e2 ** 2] * 3)),
            "R": np.diag(([D.arcsec2rad ** 2] * 2 + [1e3 ** 2])), "alpha": 0.0001, "beta": 2., "kappa": 3 - 6,
            "orbits": cat[:512], "obs_returned": "flatten", "reward_type": "jones"}
    tm = gcrs2irts_matrix_approx(time_table(base["t_0"], base["time_step"], n_steps))
    for name, over, agent in (("default", {}, EO.agent_visible_greedy), ("mask15_trinary", {"obs_limit": 15, "reward_type": "trinary"}, EO.agent_visible_greedy),
                              ("naive_greedy", {"rso_count": 7}, EO.agent_naive_greedy)):
        cfg = dict(base); cfg.update(over)
        env = EO.OracleEnv(cfg, fx, tm, tr=tr)
        env.seed(0)
        env.action_space.seed(0)
        obs = env.reset()
        acts, rews, dones, obss, vis = [], [], [], [obs.copy()], [env.visible_objects().copy()]
        done = False
        while not done and env.i + 1 < n_steps:
            a = agent(obs, env)
            obs, r, done, _ = env.step(int(a))
            acts.append(int(a)); rews.append(float(r)); dones.append(bool(done)); obss.append(obs.copy())
            v = np.zeros(cfg["rso_count"], bool); v[env.visible_objects()] = True; vis.append(v)
        v0 = np.zeros(cfg["rso_count"], bool); v0[vis[0]] = True; vis[0] = v0
        k = env.i + 1
        np.savez_compressed(os.path.join(OUT, f"golden_env_{name}.npz"), trans_matrix=tm, actions=np.array(acts), rewards=np.array(rews),
                            dones=np.array(dones), obs=np.array(obss), visible=np.array(vis), x_true=env.x_true[:k], x_filter=env.x_filter[:k],
                            P_filter=env.P_filter[:k], delta_pos=env.delta_pos[:k], z_noise=env.z_noise, orbits=cat[:512],
                            rso_count=cfg["rso_count"], obs_limit=cfg["obs_limit"], reward_type=cfg["reward_type"], n_steps=n_steps)
        print(name, "steps", len(acts), "actions", acts[:12], "failed", env.failed_filters_id)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
