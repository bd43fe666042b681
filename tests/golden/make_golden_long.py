#!/usr/bin/env python
"""Generate the 480-step golden EPISODES (tests/golden/golden_env_long_*.npz) from the REFERENCE's own functions:
the restated environment (oracle/env_oracle.py) around envs/farnocchia.py::fx_xyz_farnocchia (numba, unmodified)
and the njit geometry of envs/transformations.py, driven by agent_visible_greedy for the full default episode
length (480 steps, envs/__init__.py:23) with reward_type 'trinary' (no early termination) and a 15 degree
elevation mask, for m = 10 and m = 40 objects.

Run here in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_golden_long.py
Stored per episode: actions, rewards, dones, visibility masks, delta_pos and trace(P) per step and object, the state
every 60 steps, and a probe of the noise table (the replay re-draws it from the same seed: SS2:206-221).
"""
import os
import sys
from datetime import datetime

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402
from oracle import dynamics_restated as D  # noqa: E402
from oracle import env_oracle as EO  # noqa: E402
from ssa_gym_b200.transformations import gcrs2irts_matrix_approx, time_table  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
OBS_DEG = (38.828198, -77.305352, 20.0)


def main():
    assert rl.available(), "reference tree not found"
    far, tr, cat = rl.farnocchia(), rl.transformations(), rl.catalog()
    fx = far.fx_xyz_farnocchia
    n_steps = 480
    base = {"steps": n_steps, "rso_count": 10, "time_step": 20., "t_0": datetime(2020, 5, 4, 0, 0, 0), "obs_limit": 15,
            "observer": OBS_DEG, "update_interval": 1, "obs_type": "aer", "z_sigma": (1, 1, 1e3),
            "x_sigma": tuple([1e5] * 3 + [1e2] * 3), "q_sigma": 0.000025, "P_0": np.diag(([1e5 ** 2] * 3 + [1e2 ** 2] * 3)),
            "R": np.diag(([D.arcsec2rad ** 2] * 2 + [1e3 ** 2])), "alpha": 0.0001, "beta": 2., "kappa": 3 - 6,
            "orbits": cat[:512], "obs_returned": "flatten", "reward_type": "trinary"}
    tm = gcrs2irts_matrix_approx(time_table(base["t_0"], base["time_step"], n_steps))
    for name, m in (("long_m10", 10), ("long_m40", 40)):
        cfg = dict(base)
        cfg["rso_count"] = m
        env = EO.OracleEnv(cfg, fx, tm, tr=tr)
        env.seed(0)
        env.action_space.seed(0)
        obs = env.reset()
        acts, rews, dones, vis, trace, margins = [], [], [], [], [], []

        def mask():
            v = np.zeros(m, bool)
            v[env.visible_objects()] = True
            return v
        vis.append(mask())
        trace.append(np.array([np.trace(P) for P in env.P_filter[env.i]]))
        done = False
        while not done and env.i + 1 < n_steps:
            a = EO.agent_visible_greedy(obs, env)
            v = np.where(vis[-1])[0]
            if len(v) > 1:  # relative margin between the two largest candidate traces (how close the decision was)
                t = np.sort(trace[-1][v])
                margins.append((t[-1] - t[-2]) / t[-1])
            else:
                margins.append(np.inf)
            obs, r, done, _ = env.step(int(a))
            acts.append(int(a)); rews.append(float(r)); dones.append(bool(done))
            vis.append(mask())
            trace.append(np.array([np.trace(P) for P in env.P_filter[env.i]]))
        # the reference's OWN reproducibility: the same episode with every fx output moved by one ulp in a random
        # direction (what a different libm / compiler does to it), teacher-forced to the golden actions.  The number of
        # decisions this perturbed reference would have taken differently is the floor for any other implementation.
        prng = np.random.RandomState(99)

        def fx_ulp(x, dt):
            out = fx(x, dt)
            return np.nextafter(out, out + np.where(prng.randint(0, 2, 6) == 1, 1.0, -1.0) * np.abs(out) - (out == 0))
        env2 = EO.OracleEnv(cfg, fx_ulp, tm, tr=tr)
        env2.seed(0)
        env2.action_space.seed(0)
        obs2 = env2.reset()
        self_flip_steps, self_noise = [], []
        for kk, a_gold in enumerate(acts):
            a2 = int(EO.agent_visible_greedy(obs2, env2))
            tr2 = np.array([np.trace(P) for P in env2.P_filter[env2.i]])
            self_noise.append(np.max(np.abs(tr2 - trace[kk]) / trace[kk]))
            if a2 != a_gold:
                self_flip_steps.append(kk)
            obs2, _, _, _ = env2.step(int(a_gold))
        k = env.i + 1
        sub = np.arange(0, k, 60)
        np.savez_compressed(os.path.join(OUT, f"golden_env_{name}.npz"), trans_matrix=tm, actions=np.array(acts), rewards=np.array(rews),
                            dones=np.array(dones), visible=np.array(vis), delta_pos=env.delta_pos[:k], trace=np.array(trace),
                            margins=np.array(margins), sub_steps=sub, x_true_sub=env.x_true[sub], x_filter_sub=env.x_filter[sub],
                            z_noise_probe=env.z_noise[::37].copy(), z_noise_sum=np.sum(env.z_noise), orbits=cat[:512],
                            rso_count=m, obs_limit=cfg["obs_limit"], reward_type=cfg["reward_type"], n_steps=n_steps,
                            failed=np.array(env.failed_filters_id, dtype=int), self_flip_steps=np.array(self_flip_steps, dtype=int),
                            self_noise=np.array(self_noise))
        mg = np.array(margins)
        print(name, "steps", len(acts), "failed", env.failed_filters_id, "decisions with margin < 1e-6:", int(np.sum(mg < 1e-6)),
              "< 1e-3:", int(np.sum(mg < 1e-3)), "random (no visible object):", int(np.sum(~np.isfinite(mg))),
              "| reference vs its own 1-ulp-perturbed fx: flips", len(self_flip_steps), "first at", self_flip_steps[:3],
              "median trace discrepancy", float(np.median(self_noise)))


if __name__ == "__main__":
    main()
