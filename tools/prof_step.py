#!/usr/bin/env python
"""A few plain steps of the C2/C4-style workload for ncu: python tools/prof_step.py N impl[,impl] [steps]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ssa_gym_b200 import _lib as F  # noqa: E402
from ssa_gym_b200.ukf import BatchedUKF  # noqa: E402

n = int(sys.argv[1])
impls = sys.argv[2].split(",")
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cat, x, P0, zn = bench.workload_inputs(n, 0, 2)
cfg = bench.make_cfg(n)
flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE
M = np.array(bench.CEL2TER06AXY)
for impl in impls:
    os.environ["SSA_UKF_KERNEL"] = impl
    os.environ["SSA_UKF_STEP_GRAPH"] = "0"
    ukf = BatchedUKF(n_envs=1, m=n, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                     obs_lla=[np.radians(38.828198), np.radians(-77.305352), 20.0], obs_limit_rad=np.radians(-90.0))
    ukf.reset(cat, x, P0)
    ukf.upload(F.F_Z_NOISE, zn[0])
    for _ in range(steps):
        ukf.step(M, flags)
    ukf.sync()
    print(impl, "ok", int((ukf.download(F.F_STATUS) & 1).sum()), "failed")
    ukf.close()
