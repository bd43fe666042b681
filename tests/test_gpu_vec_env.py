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
