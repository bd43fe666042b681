"""-m gpu: BASELINE.json config 3 — E parallel environments advanced by one device step must reproduce E
independent single environments (same seeds): observations, rewards, dones, auto-resets and the device-side
greedy tasker decisions, all EXACT (the single envs run on the host twin of the same arithmetic)."""
import numpy as np
import pytest

import helpers as H
import ssa_gym_b200
from ssa_gym_b200 import _lib as F
from ssa_gym_b200 import agents
from ssa_gym_b200.transformations import gcrs2irts_matrix_approx, time_table
from ssa_gym_b200.vec_env import VecSSATaskerEnv
from test_gpu_env import make_env

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("over,tasker,agent", [
    ({"steps": 40, "reward_type": "jones"}, F.TASKER_VISIBLE_GREEDY, agents.agent_visible_greedy),
    ({"steps": 30, "reward_type": "trinary", "obs_limit": 15}, F.TASKER_POS_ERROR_GREEDY, agents.agent_pos_error_greedy),
    ({"steps": 25, "reward_type": "shaped", "rso_count": 6}, F.TASKER_NAIVE_GREEDY, agents.agent_naive_greedy),
    # the other two observation layouts of SS2:164-177 / 355-361
    ({"steps": 20, "reward_type": "jones", "rso_count": 7, "obs_returned": "aer"}, F.TASKER_VISIBLE_GREEDY, agents.agent_visible_greedy),
    ({"steps": 20, "reward_type": "trinary", "rso_count": 5, "obs_returned": "2d"}, F.TASKER_NAIVE_GREEDY, agents.agent_naive_greedy),
])
def test_vec_env_equals_independent_envs(over, tasker, agent):
    E, total_steps = 24, 55
    cfg = dict(ssa_gym_b200.env_config)
    cfg.update(over)
    cfg["trans_matrix"] = gcrs2irts_matrix_approx(time_table(cfg["t_0"], cfg["time_step"], cfg["steps"]))
    seeds = list(range(100, 100 + E))
    vec = VecSSATaskerEnv(cfg, E, seeds=seeds)
    singles = []
    for e in range(E):
        env = make_env("twin", **{k: v for k, v in cfg.items()})
        env.seed(seeds[e]); env.action_space.seed(seeds[e]); vec.action_spaces[e].seed(seeds[e])
        singles.append(env)
    vec.seed(seeds)
    obs_v = vec.vector_reset()
    obs_s = [env.reset() for env in singles]
    assert all(H.bits_equal(obs_v[e], obs_s[e]) for e in range(E))
    n_resets = 0
    for t in range(total_steps):
        a_v = vec.greedy_actions(tasker)
        a_s = [int(agent(obs_s[e], singles[e])) for e in range(E)]
        assert list(a_v) == a_s, (t, list(a_v), a_s)                                  # tasking: bit-exact
        obs_v, r_v, d_v, _ = vec.vector_step(a_v)
        for e, env in enumerate(singles):
            o, r, d, _ = env.step(a_s[e])
            assert H.bits_equal(np.float64(r), np.float64(r_v[e])) and bool(d) == bool(d_v[e]), (t, e, r, r_v[e], d, d_v[e])
            if d:
                o = env.reset()
                n_resets += 1
            obs_s[e] = o
            assert H.bits_equal(obs_v[e], np.asarray(o)), (t, e)
    assert n_resets >= E  # every env went through at least one auto-reset
    vec.close()


# ---- device-resident episodic mode (rng='device'): vectorised reset + counter-based noise --------------------------
def _emu_env_reduce(st, E, m, step_idx, reward_type, n_steps):
    """numpy restatement of ssa_env_reduce_kernel (SS2:324-354, agents.py:7-81) on the twin's outputs."""
    dpos = st.dpos.reshape(E, m); dvel = st.dvel.reshape(E, m); tr = st.trace.reshape(E, m)
    vis = st.visible.reshape(E, m).astype(bool)
    max_dpos = dpos.max(1)
    tri = ((dpos < 1e4).astype(int) + (dpos < 1e7).astype(int)).sum(1)
    trinary = (tri.astype(float) / float(m)) / 2.0
    reward = np.zeros(E); done = np.zeros(E, bool)
    if reward_type == "jones":
        hi, lo = max_dpos > 5e6, max_dpos < 3e4
        done |= hi | lo
        reward[lo & ~hi] = 1.0
    else:
        reward = trinary
    done |= (step_idx + 1 >= n_steps)
    greedy = np.full((E, F.N_TASKERS), -1, np.int32)
    greedy[:, F.TASKER_NAIVE_GREEDY] = tr.argmax(1)
    for e in range(E):
        idx = np.where(vis[e])[0]
        if np.any(idx):  # agents.py:37 tests the INDEX array: false-y when only object 0 is visible
            greedy[e, F.TASKER_VISIBLE_GREEDY] = idx[tr[e, idx].argmax()]
            greedy[e, F.TASKER_POS_ERROR_GREEDY] = idx[dpos[e, idx].argmax()]
            greedy[e, F.TASKER_VEL_ERROR_GREEDY] = idx[dvel[e, idx].argmax()]
            greedy[e, F.TASKER_VISIBLE_GREEDY_AER] = idx[np.nan_to_num(tr[e, idx], nan=0.001, posinf=0.001, neginf=0.001).argmax()]
    return reward, done, greedy


@pytest.mark.parametrize("over", [{"steps": 14, "reward_type": "jones", "update_interval": 1},
                                  {"steps": 9, "reward_type": "trinary", "update_interval": 2, "obs_limit": 10},
                                  # m > 64: the per-environment reductions (step and refresh) run CTA-wide instead of per warp
                                  {"steps": 7, "reward_type": "trinary", "update_interval": 1, "rso_count": 70}])
def test_device_episodic_mode_equals_twin_emulation(over):
    """ssa_ukf_rollout_reset / ssa_ukf_rollout_step (one graph launch per step: noise, UKF kernels, reward / done,
    auto-reset, fresh obs, greedy taskers) against a step-by-step emulation built from the host twin: the same
    counter-based draws (twin_env_reset / twin_env_noise), twin_step with the per-env trans_matrix, numpy
    reward / done / taskers.  Observations, rewards, dones and tasker decisions must be EXACT over many auto-resets."""
    E, total_steps = 48, 40
    cfg = dict(ssa_gym_b200.env_config)
    cfg.update(over)
    n, m = cfg["steps"], cfg["rso_count"]
    cfg["trans_matrix"] = gcrs2irts_matrix_approx(time_table(cfg["t_0"], cfg["time_step"], n))
    seeds = list(range(500, 500 + E))
    vec = VecSSATaskerEnv(cfg, E, seeds=seeds, rng="device")
    tw = H.twin()
    keys = np.array(seeds, dtype=np.uint64)
    table = np.ascontiguousarray(vec.trans_matrix).reshape(n, 9)
    sig = np.concatenate([vec.x_sigma, vec.z_sigma, vec.P_0[np.triu_indices(6)]])
    orbits = np.ascontiguousarray(vec.orbits, dtype=np.float64)
    N = E * m
    episode = np.zeros(E, np.uint32); step_idx = np.zeros(E, np.int32)
    xt = np.zeros((N, 6)); xf = np.zeros((N, 6)); Pp = np.zeros((N, 21)); status = np.zeros(N, np.int32); infl = np.zeros(N, np.int32)

    def emu_reset(st, done):
        tw.twin_env_reset(E, m, H.p(keys), H.p(episode), H.p(step_idx), H.p(done) if done is not None else None, H.p(orbits),
                          len(orbits), H.p(sig), H.p(xt), H.p(xf), H.p(Pp), H.p(status), H.p(infl))
        sel = np.repeat(done.astype(bool), m) if done is not None else np.ones(N, bool)
        if st is None:
            st = H.HostState(xt, xf, H.unpack_P(Pp))
        else:
            st.x_true[sel] = xt[sel]; st.x[sel] = xf[sel]; st.P[sel] = H.unpack_P(Pp)[sel]
            st.status[sel] = 0; st.infl[sel] = 0
        return st

    tcfg = H.make_cfg(N, E=E, m=m, dt=cfg["time_step"], alpha=cfg["alpha"], beta=cfg["beta"], kappa=cfg["kappa"],
                      q_sigma=cfg["q_sigma"], R=vec.R, observer_deg=cfg["observer"], obs_limit_deg=cfg["obs_limit"], n_steps=n)
    st = emu_reset(None, None)
    H.cpu_step("twin", tcfg, st, table[step_idx].reshape(E, 9), F.STEP_EPILOGUE | F.STEP_M_PER_ENV)
    _, _, g_e = _emu_env_reduce(st, E, m, step_idx, cfg["reward_type"], n)
    assert H.bits_equal(vec.obs.reshape(N, 12), st.obs)
    assert np.array_equal(vec._io["greedy"][:, :5], g_e[:, :5])     # (agent_shannon: test_gpu_vs_oracle.py)
    rng = np.random.RandomState(11)
    n_resets = 0
    for t in range(total_steps):
        a = vec.greedy_actions(F.TASKER_VISIBLE_GREEDY) if t % 2 else rng.randint(0, m, size=E)
        a = np.asarray(a, dtype=np.int32)
        obs_v, r_v, d_v, _ = vec.vector_step(a)
        # emulation
        zn = np.zeros((N, 3))
        tw.twin_env_noise(E, m, H.p(keys), H.p(episode), H.p(step_idx), H.p(sig), H.p(zn))
        nxt = step_idx + 1
        a_eff = np.where(nxt % cfg["update_interval"] == 0, a, -1).astype(np.int32)
        Ms = table[np.minimum(nxt, n - 1)].reshape(E, 9)
        H.cpu_step("twin", tcfg, st, Ms, F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ACT | F.STEP_EPILOGUE | F.STEP_M_PER_ENV,
                   actions=a_eff, z_noise=zn)
        step_idx[:] = nxt
        r_e, d_e, _ = _emu_env_reduce(st, E, m, step_idx, cfg["reward_type"], n)
        assert np.array_equal(d_v, d_e), (t, d_v, d_e)
        assert H.bits_equal(r_v, r_e), (t, r_v, r_e)
        if d_e.any():
            n_resets += int(d_e.sum())
            st = emu_reset(st, d_e.astype(np.uint8))
            H.cpu_step("twin", tcfg, st, table[step_idx].reshape(E, 9), F.STEP_EPILOGUE | F.STEP_M_PER_ENV)
        _, _, g_e = _emu_env_reduce(st, E, m, step_idx, cfg["reward_type"], n)
        assert H.bits_equal(obs_v.reshape(N, 12), st.obs), t
        assert np.array_equal(vec._io["greedy"][:, :5], g_e[:, :5]), t
        sh = vec._io["greedy"][:, F.TASKER_SHANNON]
        assert np.array_equal(sh >= 0, g_e[:, F.TASKER_VISIBLE_GREEDY] >= 0)
        assert all(st.visible.reshape(E, m)[e, sh[e]] for e in range(E) if sh[e] >= 0)
        assert np.array_equal(vec.i, step_idx), t
    assert n_resets >= E
    assert H.bits_equal(vec.ukf.download(F.F_X_FILTER), st.x) and H.bits_equal(vec.ukf.download(F.F_X_TRUE), st.x_true)
    vec.close()


def test_device_resident_step_equals_host_io_step():
    """vector_step_device (actions read from the device buffer, obs / reward / done / greedy left on the device, no copies,
    no synchronisation) is the same step as vector_step: every output bit-equal over auto-resets."""
    import torch
    E = 64
    cfg = dict(ssa_gym_b200.env_config, steps=12)
    cfg["trans_matrix"] = gcrs2irts_matrix_approx(time_table(cfg["t_0"], cfg["time_step"], cfg["steps"]))
    seeds = list(range(900, 900 + E))
    a_env = VecSSATaskerEnv(cfg, E, seeds=seeds, rng="device")
    b_env = VecSSATaskerEnv(cfg, E, seeds=seeds, rng="device")
    v = b_env.device_views()
    assert np.array_equal(v["obs"].cpu().numpy(), a_env.obs)
    rng = np.random.RandomState(4)
    n_done = 0
    for t in range(30):
        act = rng.randint(0, cfg["rso_count"], size=E).astype(np.int32)
        obs, rew, done, _ = a_env.vector_step(act)
        v["actions"].copy_(torch.from_numpy(act).to(v["actions"].device))
        b_env.vector_step_device()
        torch.cuda.synchronize()
        assert H.bits_equal(v["obs"].cpu().numpy(), obs) and H.bits_equal(v["reward"].cpu().numpy(), rew)
        assert np.array_equal(v["done"].cpu().numpy().astype(bool), done)
        assert np.array_equal(v["greedy"].cpu().numpy(), a_env._io["greedy"])
        n_done += int(done.sum())
    assert n_done > E
    a_env.close(); b_env.close()


@pytest.mark.parametrize("layout", ["flatten", "2d"])
def test_float32_observations_are_the_rounded_float64_ones(layout):
    """obs_dtype='float32' (SSA_ROLLOUT_OBS_F32): the step's observations leave the device as floats — the float64 rows
    rounded to nearest, bit for bit numpy's astype(float32) — and reward / done / greedy are untouched, over auto-resets."""
    E = 48
    cfg = dict(ssa_gym_b200.env_config, steps=10, obs_returned=layout)
    cfg["trans_matrix"] = gcrs2irts_matrix_approx(time_table(cfg["t_0"], cfg["time_step"], cfg["steps"]))
    seeds = list(range(300, 300 + E))
    a_env = VecSSATaskerEnv(cfg, E, seeds=seeds, rng="device")
    b_env = VecSSATaskerEnv(dict(cfg, obs_dtype="float32"), E, seeds=seeds, rng="device")
    assert b_env.observation_space.dtype == np.float32 and a_env.observation_space.dtype == np.float64
    o64, o32 = a_env.vector_reset(), b_env.vector_reset()
    assert o32.dtype == np.float32 and np.array_equal(o32, o64.astype(np.float32))
    rng = np.random.RandomState(5)
    n_done = 0
    for t in range(25):
        act = rng.randint(0, cfg["rso_count"], size=E).astype(np.int32)
        o64, r64, d64, _ = a_env.vector_step(act)
        o32, r32, d32, _ = b_env.vector_step(act)
        assert o32.dtype == np.float32 and o32.shape == o64.shape
        assert np.array_equal(o32.view(np.uint32), o64.astype(np.float32).view(np.uint32))
        assert H.bits_equal(r32, r64) and np.array_equal(d32, d64)
        assert np.array_equal(a_env._io["greedy"], b_env._io["greedy"])
        n_done += int(d64.sum())
    assert n_done > E
    a_env.close(); b_env.close()
    with pytest.raises(ValueError):
        VecSSATaskerEnv(dict(cfg, obs_dtype="float32", obs_returned="aer"), 4, rng="device")
    with pytest.raises(ValueError):
        VecSSATaskerEnv(dict(cfg, obs_dtype="float32"), 4, rng="host")
