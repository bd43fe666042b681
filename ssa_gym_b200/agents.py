"""Heuristic tasking agents with the reference's names and call signature `agent(obs, env)`.

Reference: agents.py:7-81.  For the single-environment drop-in (`SSA_Tasker_Env`) they read the same env
attributes the reference agents read (`env.P_filter[env.i]`, `env.visible_objects()`, `env.delta_pos`, ...),
whose arrays were produced by the GPU step.  The integer decisions reproduce numpy's conventions exactly:
`np.trace` sums the diagonal left to right, `np.argmax` returns the FIRST maximum, and `if not
np.any(visible)` tests the index array — it is also "empty" when the only visible object is index 0, in
which case the reference samples a random action (agents.py:37).

For vectorised environments the same four argmax rules are evaluated on the device by
`ssa_ukf_env_reduce` (include/ssa_ukf.h, SSA_TASKER_*), see ssa_gym_b200/vec_env.py.
"""
import numpy as np


def agent_naive_greedy(obs, env=None):
    trace = [np.trace(P) for P in env.P_filter[env.i]]
    return np.argmax(trace)


def agent_naive_random(obs=None, env=None):
    return env.action_space.sample()


def agent_shannon(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    with np.errstate(divide='ignore', invalid='ignore'):
        calculate = [(np.log(np.linalg.det(P) / np.linalg.det(P_i)))
                     for P, P_i in zip(env.P_filter[env.i][visible], env.P_filter[env.i - 1][visible])]
    visible_id = np.argmax(calculate)
    return visible[visible_id]


def agent_visible_random(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    return np.random.choice(visible)


def agent_visible_greedy(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    visible_trace = [np.trace(P) for P in env.P_filter[env.i][visible]]
    visible_id = np.argmax(visible_trace)
    return visible[visible_id]


def agent_visible_greedy_spoiled(obs, env, p=0.25):
    visible = env.visible_objects()
    random = env.action_space.sample()
    if not np.any(visible):
        return random
    visible_trace = [np.trace(P) for P in env.P_filter[env.i][visible]]
    visible_id = np.argmax(visible_trace)
    greedy = visible[visible_id]
    return np.random.choice(a=[greedy, random], p=[1 - p, p])


def agent_visible_greedy_aer(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    visible_trace = obs.reshape(int(len(obs) / 4), 4)[visible, 3]
    visible_id = np.argmax(visible_trace)
    return visible[visible_id]


def agent_pos_error_greedy(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    visible_positional_error = env.delta_pos[env.i, visible]
    visible_id = np.argmax(visible_positional_error)
    return visible[visible_id]


def agent_vel_error_greedy(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    visible_velocity_error = env.delta_vel[env.i, visible]
    visible_id = np.argmax(visible_velocity_error)
    return visible[visible_id]
