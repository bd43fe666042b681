"""ORACLE (test infrastructure): a CPU restatement of the reference environment's hot path.

Follows envs/ssa_tasker_simple_2.py:72-241 (init/reset), :243-367 (step), :369-382 (filter_error),
:410-434 (visibility) and :834-840 (aer_obs) line by line, on top of the numpy restatement of filterpy
(oracle/filterpy_restated.py) and the operator callables of oracle/dynamics_restated.py.  `fx` is the
reference's own numba function when /root/reference is present (golden-vector generation) and the C oracle
otherwise.  Plotting/diagnostic methods of the reference class are out of scope.

The heuristic agents of agents.py:7-81 and gym's `seeding.np_random` (gym <= 0.21; absent from this image)
are restated at the bottom.
"""
import hashlib
import struct

import numpy as np

from . import dynamics_restated as D
from .filterpy_restated import MerweScaledSigmaPoints, Q_discrete_white_noise, UnscentedKalmanFilter

deg2rad = np.pi / 180


# ---- gym.utils.seeding (gym 0.17-0.21) -----------------------------------------------------------------
def _bigint_from_bytes(b):
    sizeof_int = 4
    padding = sizeof_int - len(b) % sizeof_int
    b += b"\0" * padding
    int_count = int(len(b) / sizeof_int)
    unpacked = struct.unpack("{}I".format(int_count), b)
    accum = 0
    for i, val in enumerate(unpacked):
        accum += 2 ** (sizeof_int * 8 * i) * val
    return accum


def _int_list_from_bigint(bigint):
    if bigint == 0:
        return [0]
    ints = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        ints.append(mod)
    return ints


def np_random(seed=None):
    if seed is None:
        import os
        seed = _bigint_from_bytes(os.urandom(8))
    seed = int(seed) % 2 ** (8 * 8)
    h = hashlib.sha512(str(seed).encode("utf8")).digest()
    rng = np.random.RandomState()
    rng.seed(_int_list_from_bigint(_bigint_from_bytes(h[:8])))
    return rng, seed


class Discrete:
    def __init__(self, n):
        self.n = n
        self.np_random, _ = np_random(None)

    def seed(self, seed=None):
        self.np_random, seed = np_random(seed)
        return [seed]

    def sample(self):
        return self.np_random.randint(self.n)

    def contains(self, x):
        return int(x) == x and 0 <= int(x) < self.n


class OracleEnv:
    """SSA_Tasker_Env restated (hot path only)."""

    def __init__(self, config, fx, trans_matrix, tr=None, resample_after_predict=True):
        self.t_0 = config.get("t_0")
        self.dt = config["time_step"]
        self.n = config["steps"]
        self.m = config["rso_count"]
        self.obs_limit = np.radians(config["obs_limit"])
        self.obs_returned = config["obs_returned"]
        self.reward_type = config["reward_type"]
        self.orbits = config["orbits"]
        self.obs_lla = np.array(config["observer"]) * [deg2rad, deg2rad, 1]
        lla2ecef = tr.lla2ecef if tr is not None else D.lla2ecef
        self._ecef2aer = tr.ecef2aer if tr is not None else D.ecef2aer
        self.obs_itrs = lla2ecef(self.obs_lla)
        self.update_interval = config["update_interval"]
        self.i = 0
        self.obs_type = config["obs_type"]
        hx_aer, residual_aer, mean_uvw, hx_xyz = D.make_operators(tr)
        if self.obs_type == "aer":
            self.z_sigma = config["z_sigma"] * np.array([D.arcsec2rad, D.arcsec2rad, 1])
            self.hx, self.mean_z, self.residual_z = hx_aer, mean_uvw, residual_aer
        else:
            self.z_sigma = np.array(config["z_sigma"])
            self.hx, self.mean_z, self.residual_z = hx_xyz, None, None
        self._hx_aer = hx_aer
        self.x_sigma = np.array(config["x_sigma"])
        self.Q = Q_discrete_white_noise(dim=2, dt=self.dt, var=config["q_sigma"] ** 2, block_size=3, order_by_dim=False)
        self.fx = fx
        self.msqrt = D.robust_cholesky
        self.alpha, self.beta, self.kappa = config["alpha"], config["beta"], config["kappa"]
        self.resample_after_predict = resample_after_predict
        x_dim, z_dim = 6, 3
        self.P_0 = np.copy(np.diag(self.x_sigma ** 2)) if config["P_0"] is None else np.copy(config["P_0"])
        self.R = np.diag(self.z_sigma ** 2) if config["R"] is None else np.copy(config["R"])
        n, m = self.n, self.m
        self.x_true = np.empty((n, m, x_dim))
        self.x_filter = np.empty((n, m, x_dim))
        self.P_filter = np.empty((n, m, x_dim, x_dim))
        self.obs = np.empty((n, m, x_dim * 2))
        self.trans_matrix = np.asarray(trans_matrix)
        self.z_noise = np.empty((n, m, z_dim))
        self.z_true = np.empty((n, m, z_dim))
        self.y = np.empty((n, m, z_dim))
        self.S = np.empty((n, m, z_dim, z_dim))
        self.x_noise = np.empty((m, x_dim))
        self.filters = []
        self.delta_pos = np.empty((n, m))
        self.delta_vel = np.empty((n, m))
        self.sigma_pos = np.empty((n, m))
        self.sigma_vel = np.empty((n, m))
        self.rewards = np.empty(n)
        self.failed_filters_id = []
        self.failed_filters_msg = ["None"] * m
        self.actions = np.empty(n, dtype=int)
        self.obs_taken = np.empty(n, dtype=bool)
        self.x_failed = np.array([1e20, 1e20, 1e20, 1e12, 1e12, 1e12])
        self.P_failed = np.diag([1e20, 1e20, 1e20, 1e12, 1e12, 1e12])
        self.sigmas_h = np.empty((n, x_dim * 2 + 1, z_dim))
        self.action_space = Discrete(m)
        self.observation = np.zeros(m * 4)
        self.np_random = None
        self.init_seed = self.seed()
        self.reset()

    def seed(self, seed=None):  # SS2:188-191
        self.np_random, seed = np_random(seed)
        self.init_seed = seed
        return [seed]

    def reset(self):  # SS2:193-241
        self.x_true[:], self.x_filter[:], self.P_filter[:], self.obs[:], self.sigmas_h[:] = [0] * 5
        self.z_true[:], self.y[:], self.S[:] = np.nan, np.nan, np.nan
        self.filters = []
        for j in range(self.m):
            self.x_true[0][j] = self.orbits[self.np_random.randint(low=0, high=self.orbits.shape[0]), :]
            self.x_noise[j] = self.np_random.normal(size=6) * self.x_sigma
            self.x_filter[0][j] = np.copy(self.x_true[0][j] + self.x_noise[j])
            self.P_filter[0][j] = np.copy(self.P_0)
            ukf = UnscentedKalmanFilter(dim_x=6, dim_z=3, dt=self.dt, fx=self.fx, hx=self.hx,
                                        points=MerweScaledSigmaPoints(n=6, alpha=self.alpha, beta=self.beta, kappa=self.kappa,
                                                                      sqrt_method=self.msqrt),
                                        z_mean_fn=self.mean_z, residual_z=self.residual_z, sqrt_fn=self.msqrt,
                                        resample_after_predict=self.resample_after_predict)
            ukf.x = np.copy(self.x_filter[0][j])
            ukf.P = np.copy(self.P_filter[0][j])
            ukf.R = np.copy(self.R)
            ukf.Q = np.copy(self.Q)
            self.filters.append(ukf)
        for i in range(self.n):
            for j in range(self.m):
                self.z_noise[i, j] = self.np_random.normal(size=3) * self.z_sigma
        self.delta_pos[:], self.delta_vel[:], self.sigma_pos[:], self.sigma_vel[:] = [np.nan] * 4
        self.actions[:], self.obs_taken[:], self.failed_filters_id = 0, False, []
        self.failed_filters_msg = ["None"] * self.m
        self.obs[0] = D.observations(self.x_filter[0], self.P_filter[0])
        self.delta_pos[0], self.delta_vel[0], self.sigma_pos[0], self.sigma_vel[0] = D.error(self.x_true[0], self.obs[0])
        self.rewards[:] = 0
        self.i = 0
        if self.obs_returned == "flatten":
            return self.obs[0].flatten()
        elif self.obs_returned == "aer":
            self.observation = self.aer_obs(np.zeros(self.m * 4))
            return self.observation
        return self.obs[0]

    def filter_error(self, object_id, code):  # SS2:369-382
        self.filters[object_id].x = np.copy(self.x_failed)
        self.filters[object_id].P = np.copy(self.P_failed)
        self.failed_filters_msg[object_id] = code
        self.failed_filters_id.append(object_id)

    def step(self, a):  # SS2:243-367
        assert self.action_space.contains(a), "%r (%s) invalid" % (a, type(a))
        self.i += 1
        i = self.i
        self.actions[i] = np.copy(a)
        for j in range(self.m):
            self.x_true[i][j] = self.fx(self.x_true[i - 1][j], self.dt)
        for j in range(self.m):
            if not (j in self.failed_filters_id):
                try:
                    self.filters[j].predict()
                    if np.any(np.isnan(self.filters[j].x)):
                        self.filter_error(j, "predict nan")
                except ValueError:
                    self.filter_error(j, "predict ValueError")
                except np.linalg.LinAlgError:
                    self.filter_error(j, "predict LinAlgError")
                except Exception:
                    self.filter_error(j, "predict Unknown")
            self.x_filter[i, j] = np.copy(self.filters[j].x)
            self.P_filter[i, j] = np.copy(self.filters[j].P)
        if (i % self.update_interval) == 0:
            if not (a in self.failed_filters_id):
                hx_kwargs = {"trans_matrix": self.trans_matrix[i], "observer_itrs": self.obs_itrs,
                             "observer_lla": self.obs_lla, "time": None}
                self.z_true[i, a] = self.hx(self.x_true[i][a], **hx_kwargs)
                if self.object_visible([a])[0]:
                    try:
                        self.filters[a].update(self.z_true[i, a] + self.z_noise[i, a], **hx_kwargs)
                        self.y[i, a] = np.copy(self.filters[a].y)
                        self.S[i, a] = np.copy(self.filters[a].S)
                        self.sigmas_h[i] = np.copy(self.filters[a].sigmas_h)
                        self.obs_taken[i] = True
                        if np.any(np.isnan(self.filters[a].x)):
                            self.filter_error(a, "update nan")
                    except ValueError:
                        self.filter_error(a, "update ValueError")
                    except np.linalg.LinAlgError:
                        self.filter_error(a, "update LinAlgError")
                    except Exception:
                        self.filter_error(a, "update Unknown")
                    self.x_filter[i, a] = np.copy(self.filters[a].x)
                    self.P_filter[i, a] = np.copy(self.filters[a].P)
        self.obs[i] = D.observations(self.x_filter[i], self.P_filter[i])
        self.delta_pos[i], self.delta_vel[i], self.sigma_pos[i], self.sigma_vel[i] = D.error(self.x_true[i], self.obs[i])
        done = False
        if self.reward_type == "jones":
            if np.max(self.delta_pos[i]) > 5e6:
                done, self.rewards[i] = True, 0
            elif np.max(self.delta_pos[i]) < 3e4:
                done, self.rewards[i] = True, 1
            elif i + 1 >= self.n:
                done, self.rewards[i] = True, 0
            else:
                done, self.rewards[i] = False, 0
        elif self.reward_type == "trinary":
            self.rewards[i] = D.reward_proportional_trinary_true(self.delta_pos[i])
        elif self.reward_type == "shaped":
            if np.max(self.delta_pos[i]) > 5e6:
                done, self.rewards[i] = True, 0
            elif np.max(self.delta_pos[i]) < 3e4:
                done = True
                self.rewards[i] = 1 - np.sum(self.rewards[:i])
            elif a == np.argmax(self.sigma_pos[i - 1]):
                self.rewards[i] = 1 / self.n
            else:
                self.rewards[i] = -1 / self.n
        if i + 1 >= self.n:
            done = True
        if self.obs_returned == "flatten":
            return self.obs[i].flatten(), self.rewards[i], done, {}
        elif self.obs_returned == "aer":
            self.observation = self.aer_obs(self.observation)
            return self.observation, np.nan_to_num(self.rewards[i], nan=0.5, posinf=0.5, neginf=0.5), done, {}
        return self.obs[i], np.nan_to_num(self.rewards[i], nan=0.5, posinf=0.5, neginf=0.5), done, {}

    def visible_objects(self):  # SS2:410-416
        return np.where(self.object_visible([j for j in range(self.m)]))[0]

    def object_visible(self, RSO_ID):  # SS2:418-425
        x_itrs = np.array([self.trans_matrix[self.i] @ self.x_true[self.i, j, :3] for j in RSO_ID])
        el = np.array([self._ecef2aer(self.obs_lla, x, self.obs_itrs)[1] for x in x_itrs])
        return el >= self.obs_limit

    def aer_obs(self, obs):  # SS2:834-840
        i = self.i
        for j in range(self.m):
            obs[4 * j: 4 * j + 3] = self._hx_aer(self.x_filter[i, j, :3], self.trans_matrix[i], self.obs_lla, self.obs_itrs)
            obs[4 * j + 3] = np.trace(self.P_filter[i, j])
        return np.nan_to_num(obs, copy=False, nan=0.001, posinf=0.001, neginf=0.001)


# ---- agents.py:7-81 --------------------------------------------------------------------------------------
def agent_naive_greedy(obs, env=None):
    return np.argmax([np.trace(P) for P in env.P_filter[env.i]])


def agent_visible_greedy(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    visible_trace = [np.trace(P) for P in env.P_filter[env.i][visible]]
    return visible[np.argmax(visible_trace)]


def agent_pos_error_greedy(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    return visible[np.argmax(env.delta_pos[env.i, visible])]


def agent_vel_error_greedy(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    return visible[np.argmax(env.delta_vel[env.i, visible])]


def agent_visible_greedy_aer(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    visible_trace = obs.reshape(int(len(obs) / 4), 4)[visible, 3]
    return visible[np.argmax(visible_trace)]


def agent_shannon(obs, env):
    visible = env.visible_objects()
    if not np.any(visible):
        return env.action_space.sample()
    with np.errstate(divide="ignore", invalid="ignore"):
        calc = [np.log(np.linalg.det(P) / np.linalg.det(P_i))
                for P, P_i in zip(env.P_filter[env.i][visible], env.P_filter[env.i - 1][visible])]
    return visible[np.argmax(calc)]
