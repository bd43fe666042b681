"""Host-side episode logic shared by the single drop-in environment (env.py) and the vectorised one (vec_env.py,
rng='host'): the random draws of a reset and the reward / done rule of a step.

Stream parity with the reference is the point of the host-RNG mode, so the draws consume the generator in exactly
the reference's order (SS2:206-221): per object one `randint` for the catalog row and one `normal(size=6)` for the
initial filter error, then the measurement noise of the whole episode object by object, step by step.  One
`normal(size=(n, m, 3))` call produces the same numbers as the reference's n*m successive `normal(size=3)` calls
(numpy's legacy generator fills arrays in C order from the same stream) at a fraction of the interpreter overhead.
"""
import numpy as np


def draw_episode(rng, orbits, m, n, x_sigma, z_sigma):
    """Returns (x_true0 [m,6], x_noise [m,6], z_noise [n,m,3]) for one environment."""
    x_true0 = np.empty((m, 6))
    x_noise = np.empty((m, 6))
    for j in range(m):
        x_true0[j] = orbits[rng.randint(low=0, high=orbits.shape[0]), :]
        x_noise[j] = rng.normal(size=6) * x_sigma
    z_noise = rng.normal(size=(n, m, 3)) * z_sigma
    return x_true0, x_noise, z_noise


def step_reward(reward_type, i, n, action, delta_pos_i, sigma_pos_prev, rewards_so_far):
    """(reward, done) of step i (SS2:324-354).  rewards_so_far = rewards[:i] (only the 'shaped' rule reads it)."""
    worst = np.max(delta_pos_i)
    done, reward = False, 0
    if reward_type == 'trinary':                                     # results.py:431-433
        reward = np.mean(((delta_pos_i < 1e4) * 1 + (delta_pos_i < 1e7) * 1)) / 2
    elif worst > 5e6:                                                # lost an object: episode over, no reward
        done, reward = True, 0
    elif worst < 3e4:                                                # every object within 30 km: success
        done, reward = True, (1 if reward_type == 'jones' else 1 - np.sum(rewards_so_far))
    elif reward_type == 'shaped':
        reward = 1 / n if action == np.argmax(sigma_pos_prev) else -1 / n
    if i + 1 >= n:
        done = True
    return reward, done
