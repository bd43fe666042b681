"""Drop-in `SSA_Tasker_Env`: the reference's gym environment with the estimation hot path on the GPU.

Same constructor (`SSA_Tasker_Env(config)`), same `seed / reset / step` API (old 4-tuple gym step), same
observation / reward / done semantics, same public attributes the agents and scripts read
(`i, n, m, dt, P_filter, x_filter, x_true, delta_pos, delta_vel, sigma_pos, sigma_vel, rewards, actions,
failed_filters_id, visible_objects(), action_space, init_seed, ...`) — reference: envs/
ssa_tasker_simple_2.py:72-434, 834-840.  What changed is WHERE the work runs: the per-RSO Python loop over
filterpy objects (SS2:265-315) is one fused launch sequence of `libssa_ukf.so` over all m objects
(ssa_gym_b200/csrc/ssa_ukf.cu); RNG, configuration, failure bookkeeping, reward logic and the history arrays
stay on the host exactly as in the reference.

The plug points `fx, hx, mean_z, residual_z, msqrt` of env_config are accepted by identity/name and mapped
to the built-in device operators (ssa_gym_b200/dynamics.py); anything else raises — there is no CPU
fallback and no per-sigma-point Python callback.
"""
import time
from copy import copy
from datetime import datetime

import numpy as np

from . import _lib, dynamics
from .episode import draw_episode, step_reward
from .gym_shim import Env, seeding, spaces
from .transformations import arcsec2rad, default_eops, deg2rad, gcrs2irts_matrix_b, lla2ecef, load_eop_c04, time_table
from .ukf import BatchedUKF, Q_discrete_white_noise_block

F = _lib


class SSA_Tasker_Env(Env):
    metadata = {"render.modes": ["live", "none"]}
    visualization = None

    # timers of the reference's `runtime` table that still mean something here (the plot_* entries have no counterpart)
    RUNTIME_KEYS = ('__init__', 'reset', 'step', 'perform predictions', 'Observations and Reward', 'filter_error',
                    'visible_objects')
    # per-step history arrays of the reference (SS2:132-161): name -> trailing shape
    HISTORIES = {'x_true': (6,), 'x_filter': (6,), 'P_filter': (6, 6), 'obs': (12,), 'z_noise': (3,), 'z_true': (3,), 'y': (3,),
                 'S': (3, 3), 'delta_pos': (), 'delta_vel': (), 'sigma_pos': (), 'sigma_vel': (), 'scores': (), 'nees': ()}

    def __init__(self, config=None):
        t_start = time.time()
        if config is None:
            from . import env_config
            config = env_config
        self.runtime = dict.fromkeys(self.RUNTIME_KEYS, 0)
        self._read_config(config)
        self._allocate_histories()
        self._make_spaces()
        # the device-resident filters: one handle for the m RSOs of this environment
        self._device = config.get('device', 0)
        self.ukf = BatchedUKF(n_envs=1, m=self.m, dt=self.dt, Q=self.Q, R=self.R, obs_lla=self.obs_lla,
                              obs_limit_rad=self.obs_limit, alpha=self.alpha, beta=self.beta, kappa=self.kappa,
                              obs_type=self.obs_type, reward_type=self.reward_type, n_steps=self.n,
                              resample_after_predict=config.get('resample_after_predict', True), device=self._device)
        self.np_random = None
        self.init_seed = self.seed()
        self.reset()
        self.runtime['__init__'] += time.time() - t_start

    def _read_config(self, config):
        """env_config (envs/__init__.py:23-28) -> scalars, noise models, operators, trans_matrix table (SS2:81-137)."""
        g = config.get
        self.t_0, self.dt, self.n, self.m = g('t_0', datetime(2020, 5, 4, 0, 0, 0)), config['time_step'], config['steps'], config['rso_count']
        self.obs_limit = np.radians(config['obs_limit'])
        self.obs_returned, self.reward_type = config['obs_returned'], config['reward_type']
        self.obs_type, self.update_interval = config['obs_type'], config['update_interval']
        self.alpha, self.beta, self.kappa = config['alpha'], config['beta'], config['kappa']
        self.orbits = np.asarray(g('orbits') if g('orbits') is not None else _default_orbits())
        self.obs_lla = np.array(config['observer']) * [deg2rad, deg2rad, 1]
        self.obs_itrs = lla2ecef(self.obs_lla)
        # operator plug points: resolved to the device implementations, never called from the hot path
        dynamics.validate_operators(config, self.obs_type)
        self.fx, self.msqrt = g('fx', dynamics.fx_xyz_farnocchia), g('msqrt', dynamics.robust_cholesky)
        self.hx = g('hx') or (dynamics.hx_aer_erfa if self.obs_type == 'aer' else dynamics.hx_xyz)
        self.mean_z, self.residual_z = g('mean_z'), g('residual_z')
        # noise models: measurement sigmas in (arcsec, arcsec, m) for 'aer' (SS2:100-104)
        z_unit = np.array([arcsec2rad, arcsec2rad, 1]) if self.obs_type == 'aer' else 1.0
        self.z_sigma = np.asarray(config['z_sigma']) * z_unit
        self.x_sigma = np.array(config['x_sigma'])
        self.Q = Q_discrete_white_noise_block(self.dt, config['q_sigma'] ** 2, 3)
        self.P_0 = np.copy(np.diag(self.x_sigma ** 2)) if g('P_0') is None else np.copy(config['P_0'])
        self.R = np.diag(self.z_sigma ** 2) if g('R') is None else np.copy(config['R'])
        # GCRS -> ITRS rotation of every step: an INPUT of the path (SS2:136-137)
        self.time = time_table(self.t_0, self.dt, self.n)
        if g('trans_matrix') is not None:
            self.trans_matrix = np.ascontiguousarray(config['trans_matrix'], dtype=np.float64)
            assert self.trans_matrix.shape == (self.n, 3, 3), "trans_matrix must be [steps, 3, 3]"
        else:
            self.eops = load_eop_c04(config['eop_file']) if g('eop_file') else default_eops()
            self.trans_matrix = np.ascontiguousarray(gcrs2irts_matrix_b(self.time, self.eops))

    def _allocate_histories(self):
        n, m = self.n, self.m
        for name, tail in self.HISTORIES.items():
            setattr(self, name, np.empty((n, m) + tail))
        self.sigmas_h = np.empty((n, 13, 3))          # of the tasked object only (SS2:304)
        self.x_noise = np.empty((m, 6))
        self.rewards = np.empty(n)
        self.actions = np.empty(n, dtype=int)
        self.obs_taken = np.empty(n, dtype=bool)
        self.i = 0
        self.failed_filters_id, self.failed_filters_msg, self.visibility = [], ["None"] * m, []
        self.x_failed = np.array([1e20] * 3 + [1e12] * 3)   # sentinels of a failed filter (SS2:157-158)
        self.P_failed = np.diag(self.x_failed)
        self._visible_now = np.zeros(m, dtype=bool)

    def _make_spaces(self):
        """SS2:164-177: Discrete(m) actions; observations 'flatten' [m*12], 'aer' [m*4] or [m, 12]."""
        m = self.m
        self.action_space = spaces.Discrete(m)
        shape = {'flatten': (m * 12,), 'aer': (m * 4,)}.get(self.obs_returned, (m, 12))
        self.observation_space = spaces.Box(low=np.full(shape, -np.inf), high=np.full(shape, np.inf), dtype=np.float64)
        if self.obs_returned == 'aer':
            self.observation = np.zeros(m * 4)

    # ------------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        self.np_random, seed = seeding.np_random(seed)
        self.init_seed = seed
        return [seed]

    def reset(self):
        s = time.time()
        self.x_true[:], self.x_filter[:], self.P_filter[:], self.obs[:], self.sigmas_h[:] = [0] * 5
        self.z_true[:], self.y[:], self.S[:] = np.nan, np.nan, np.nan
        # the random draws of the episode, in the reference's order (SS2:206-221)
        self.x_true[0], self.x_noise[:], self.z_noise[:] = draw_episode(self.np_random, self.orbits, self.m, self.n,
                                                                          self.x_sigma, self.z_sigma)
        self.x_filter[0] = self.x_true[0] + self.x_noise
        self.P_filter[0] = self.P_0
        self.scores[:], self.delta_pos[:], self.delta_vel[:], self.sigma_pos[:], self.sigma_vel[:] = [np.nan] * 5
        self.actions[:], self.obs_taken[:], self.failed_filters_id, self.visibility = 0, False, [], []
        self.failed_filters_msg = ["None"] * self.m
        self.ukf.reset(self.x_true[0], self.x_filter[0], self.P_0)
        # obs[0], error(x_true[0], obs[0]) and the visibility at step 0: one epilogue-only launch
        self.ukf.step(self.trans_matrix[0], F.STEP_EPILOGUE)
        self._pull(0)
        self.rewards[:] = 0
        self.i = 0
        self.runtime['reset'] += time.time() - s
        if self.obs_returned == 'flatten':
            return self.obs[0].flatten()
        elif self.obs_returned == 'aer':
            self.observation = self.aer_obs(np.zeros(self.m * 4))
            return self.observation
        return self.obs[0]

    def _pull(self, i):
        """All history rows of step i in one device-to-host copy (ssa_ukf_snapshot)."""
        v = self.ukf.snapshot()
        self.x_true[i] = v["x_true"]
        self.x_filter[i] = v["x_filter"]
        self.P_filter[i] = v["P_filter"]
        self.obs[i] = v["obs"]
        self.delta_pos[i], self.delta_vel[i] = v["delta_pos"], v["delta_vel"]
        self.sigma_pos[i], self.sigma_vel[i] = v["sigma_pos"], v["sigma_vel"]
        self._visible_now = v["visible"].astype(bool)
        return v

    def step(self, a):
        step_s = time.time()
        assert self.action_space.contains(a), "%r (%s) invalid" % (a, type(a))
        self.i += 1
        i = self.i
        self.actions[i] = np.copy(a)
        flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_EPILOGUE | F.STEP_RECORD
        do_update = (i % self.update_interval) == 0
        if do_update:
            flags |= F.STEP_UPDATE_ACT
        s = time.time()
        self.ukf.upload(F.F_ACTIONS, np.array([a], dtype=np.int32))
        self.ukf.upload(F.F_Z_NOISE, self.z_noise[i])
        self.ukf.step(self.trans_matrix[i], flags)
        snap = self._pull(i)
        status = snap["status"].copy()
        self.runtime['perform predictions'] += time.time() - s
        if np.any(status & F.ST_TRUTHEXC):
            # the reference propagates an uncaught numba exception out of step() here (SS2:266)
            raise ArithmeticError("fx raised while propagating a true state")
        if do_update and not (a in self.failed_filters_id):
            if not (status[a] & F.ST_FAILED) or (status[a] & F.ST_IN_UPDATE):
                self.z_true[i, a] = snap["z_true"][a]
            if snap["updated"][a]:
                self.y[i, a] = snap["y"][a]
                self.S[i, a] = snap["S"][a]
                self.sigmas_h[i] = snap["sigmas_h"][a]
                self.obs_taken[i] = True
        for j in np.where(status & F.ST_FAILED)[0]:
            if j not in self.failed_filters_id:
                self.filter_error(int(j), int(status[j]))
        s = time.time()
        self.rewards[i], done = step_reward(self.reward_type, i, self.n, a, self.delta_pos[i], self.sigma_pos[i - 1],
                                            self.rewards[:i])
        self.runtime['Observations and Reward'] += time.time() - s
        self.runtime['step'] += time.time() - step_s
        if self.obs_returned == 'flatten':
            return self.obs[i].flatten(), self.rewards[i], done, {}
        elif self.obs_returned == 'aer':
            self.observation = self.aer_obs(self.observation)
            return self.observation, np.nan_to_num(self.rewards[i], nan=0.5, posinf=0.5, neginf=0.5), done, {}
        return self.obs[i], np.nan_to_num(self.rewards[i], nan=0.5, posinf=0.5, neginf=0.5), done, {}

    def filter_error(self, object_id, status):
        s = time.time()
        activity = 'update' if status & F.ST_IN_UPDATE else 'predict'
        kind = (', LinAlgError. ' if status & F.ST_LINALG else ', %s returned nan. ' % activity if status & F.ST_NAN
                else ', Unknown. ')
        prev = max(self.i - 1, 0)
        err = np.array([np.sqrt(np.sum((self.x_filter[prev, object_id, :3] - self.x_true[prev, object_id, :3]) ** 2)),
                        np.sqrt(np.sum((self.x_filter[prev, object_id, 3:] - self.x_true[prev, object_id, 3:]) ** 2)),
                        np.sqrt(np.sum(np.diag(self.P_filter[prev, object_id])[:3])),
                        np.sqrt(np.sum(np.diag(self.P_filter[prev, object_id])[3:]))])
        msg = ["".join(['Object ', str(object_id), ' failed on ', activity, ' step ', str(self.i), kind, str(np.round(err, 2))])]
        self.failed_filters_msg[object_id] = copy(msg)
        self.failed_filters_id.append(object_id)
        self.runtime['filter_error'] += time.time() - s

    def render(self, mode='live'):
        return None  # plotting is out of scope

    def visible_objects(self):
        s = time.time()
        viz = np.where(self._visible_now)[0]
        self.runtime['visible_objects'] += time.time() - s
        return viz

    def object_visible(self, RSO_ID=[]):
        if not RSO_ID:
            print('RSO ID expected, but not supplied')
            return RSO_ID
        return self._visible_now[np.asarray(RSO_ID)]

    def object_visibility(self):
        return self._visible_now.copy()

    def failed_filters(self):
        if not self.failed_filters_id:
            print("No failed Objects")
        else:
            for rso_id in self.failed_filters_id:
                print(self.failed_filters_msg[rso_id])

    def aer_obs(self, obs):
        i = self.i
        aer = dynamics.hx_aer_erfa(self.x_filter[i], self.trans_matrix[i], self.obs_lla, self.obs_itrs, device=self._device)
        for j in range(self.m):
            obs[4 * j: 4 * j + 3] = aer[j]
            obs[4 * j + 3] = np.trace(self.P_filter[i, j])
        return np.nan_to_num(obs, copy=False, nan=0.001, posinf=0.001, neginf=0.001)

    # -- consistency diagnostics (SS2:436-446, 564-624), evaluated on the device over the episode's histories --------
    def _history_diagnostics(self):
        """NEES of every (step, object) up to the current step and NIS / innovation-bound flags of every tasked
        observation, in one launch of ssa_ukf_diagnostics over the host histories (n*m pseudo-objects)."""
        steps = min(self.i + 1, self.n)
        N = steps * self.m
        tmp = BatchedUKF(n_envs=1, m=N, dt=self.dt, Q=self.Q, R=self.R, obs_lla=self.obs_lla, obs_limit_rad=self.obs_limit,
                         alpha=self.alpha, beta=self.beta, kappa=self.kappa, obs_type=self.obs_type,
                         device=self._device)
        try:
            tmp.reset(self.x_true[:steps].reshape(N, 6), self.x_filter[:steps].reshape(N, 6),
                      self.P_filter[:steps].reshape(N, 36))
            upd = np.zeros((steps, self.m), dtype=np.uint8)
            for i in range(1, steps):
                a = int(self.actions[i])
                if not np.isnan(self.y[i, a]).any():
                    upd[i, a] = 1
            tmp.upload(F.F_Y, np.nan_to_num(self.y[:steps].reshape(N, -1)))
            tmp.upload(F.F_S, np.nan_to_num(self.S[:steps].reshape(N, 3, 3)))
            tmp.upload(F.F_UPDATED, upd.reshape(N))
            nees, nis, flags = tmp.diagnostics()
        finally:
            tmp.close()
        return nees.reshape(steps, self.m), nis.reshape(steps, self.m), flags.reshape(steps, self.m)

    def anees(self):
        """SS2:436-446: average normalised estimation error squared over the episode (steps taken so far)."""
        nees, _, _ = self._history_diagnostics()
        self.nees[:] = np.nan
        self.nees[:len(nees)] = nees
        # NaN marks a covariance that is not positive definite at that step (the reference's explicit inverse
        # returns an arbitrary number there); such entries are left out of the average
        return np.nanmean(nees)

    def nis(self):
        """The series plotted by SS2:564-569: NIS of the tasked object's innovation at every step that had an update."""
        _, nis, flags = self._history_diagnostics()
        return np.array([nis[i, int(self.actions[i])] for i in range(1, len(nis)) if flags[i, int(self.actions[i])] & 0x80])

    def innovation_bounds(self):
        """SS2:598-604: percentage of the tasked innovations inside one / two innovation standard deviations,
        rows = (sigma, two sigmas), columns = measurement components (the reference renders this as a table)."""
        _, _, flags = self._history_diagnostics()
        f = np.array([flags[i, int(self.actions[i])] for i in range(1, len(flags)) if flags[i, int(self.actions[i])] & 0x80])
        if len(f) == 0:
            return np.full((2, 3), np.nan)
        one = np.stack([(f >> a) & 1 for a in range(3)], axis=1)
        two = np.stack([(f >> (3 + a)) & 1 for a in range(3)], axis=1)
        return np.round(np.stack((np.mean(one, axis=0), np.mean(two, axis=0))) * 100, 2)

    # -- innovation whiteness (SS2:644-698, 782-832): Durbin-Watson statistic and autocorrelation, on the device ------
    def innovation(self):
        """SS2:644-653: the tasked object's innovation at every step (NaN where no observation was taken) and, per object,
        the same series with the steps at which ANOTHER object was observed set to NaN."""
        steps = min(self.i + 1, self.n)
        innovation = np.array([self.y[i, int(self.actions[i])] for i in range(1, steps)]).reshape(-1, 3)
        taken = np.asarray(self.obs_taken[1:steps], dtype=bool)
        innovation[~taken] = np.nan
        innovations = []
        for j in range(self.m):
            inn = innovation.copy()
            inn[(np.asarray(self.actions[1:steps]) != j) & taken] = np.nan
            innovations.append(inn)
        return innovation, innovations

    def _innovation_stats(self, nlags):
        from .ukf import innovation_stats
        innovation, innovations = self.innovation()
        series = np.stack([innovation] + innovations)            # [1 + m, steps - 1, 3]
        valid = ~np.isnan(series).any(axis=2)
        nlags = min(nlags, series.shape[1] - 1)
        return innovation_stats(series, valid, nlags, self._device), valid

    def autocorrelation(self, nlags=40):
        """SS2:655-668: acf(innovation[:, c], missing='conservative', fft=False) of the overall innovation [3, nlags + 1] and
        of every object's own series (list of [3, nlags + 1]); nlags = 40 is the statsmodels default of the reference's era."""
        (_, acf), _ = self._innovation_stats(nlags)
        return acf[0], [acf[1 + j] for j in range(self.m)]

    def innovation_dw_test(self):
        """SS2:782-832: Durbin-Watson statistic of the innovation, rows = (all observations, min per object, max per object),
        columns = measurement components, rounded to 3 decimals.  Steps without an observation are skipped (the reference's
        sums turn NaN there); objects observed fewer than twice do not enter min / max."""
        (dw, _), valid = self._innovation_stats(1)
        per_obj = dw[1:][valid[1:].sum(axis=1) >= 2]
        lo = np.min(per_obj, axis=0) if len(per_obj) else np.full(3, np.nan)
        hi = np.max(per_obj, axis=0) if len(per_obj) else np.full(3, np.nan)
        table = np.round(np.vstack([dw[0], lo, hi]), 3)
        try:
            import pandas as pd
            cols = ['x', 'y', 'z'] if self.obs_type == 'xyz' else ['Azimuth', 'Elevation', 'Range']
            out = pd.DataFrame(data=table, columns=cols, index=['Durbin-Watson Statistic for All Obs',
                                                                 'Durbin-Watson Statistic for Min per Obj',
                                                                 'Durbin-Watson Statistic for Max per Obj'])
            out.name = 'Durbin-Watson Statistic for Innovation'
            return out
        except ImportError:
            return table

    def close(self):
        if getattr(self, "ukf", None) is not None:
            self.ukf.close()
            self.ukf = None


_orbits_cache = {}


def _default_orbits():
    if "o" not in _orbits_cache:
        from .catalog import synthetic_catalog
        _orbits_cache["o"] = synthetic_catalog(20000, 0)
    return _orbits_cache["o"]
