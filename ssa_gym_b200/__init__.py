"""ssa_gym_b200 — B200-native implementation of ssa-gym's per-step UKF hot path.

Mirrors the reference package `envs` (envs/__init__.py:18-28): `env_config` with the same keys and
defaults, the registered id 'ssa_tasker_simple-v2' and the `SSA_Tasker_Env` class.  Importing the package
does not need a GPU; constructing an environment does (no CPU fallback).
"""
from datetime import datetime

import numpy as np

from .dynamics import (fx_xyz_farnocchia as fx, hx_aer_erfa as hx, mean_z_uvw as mean_z,  # noqa: F401
                       residual_z_aer as residual_z, robust_cholesky)
from .transformations import arcsec2rad

__version__ = "0.1.0"

# envs/__init__.py:23-28 — 'orbits': None selects the built-in synthetic 20 000-orbit catalog
# (ssa_gym_b200/catalog.py); pass the reference's .npy array here to reproduce its episodes.
env_config = {'steps': 480, 'rso_count': 10, 'time_step': 20., 't_0': datetime(2020, 5, 4, 0, 0, 0), 'obs_limit': -90,
              'observer': (38.828198, -77.305352, 20.0), 'update_interval': 1, 'obs_type': 'aer',
              'z_sigma': (1, 1, 1e3), 'x_sigma': tuple([1e5] * 3 + [1e2] * 3), 'q_sigma': 0.000025,
              'P_0': np.diag(([1e5 ** 2] * 3 + [1e2 ** 2] * 3)), 'R': np.diag(([arcsec2rad ** 2] * 2 + [1e3 ** 2])),
              'alpha': 0.0001, 'beta': 2., 'kappa': 3 - 6, 'fx': fx, 'hx': hx, 'mean_z': mean_z, 'residual_z': residual_z,
              'msqrt': robust_cholesky, 'orbits': None, 'obs_returned': 'flatten', 'reward_type': 'jones'}

ENV_ID = 'ssa_tasker_simple-v2'


def make(config=None):
    """`gym.make('ssa_tasker_simple-v2', config=cfg)` equivalent."""
    from .env import SSA_Tasker_Env
    return SSA_Tasker_Env(env_config if config is None else config)


try:  # register with gym when it is installed (envs/__init__.py:18-21)
    from gym.envs.registration import register
    register(id=ENV_ID, entry_point='ssa_gym_b200.env:SSA_Tasker_Env')
except Exception:  # noqa: BLE001 - gym is absent from this image
    pass
