#!/usr/bin/env python
"""Where a C3 vector step spends its time on the HOST side (development tool): wall-clock segments of
VecSSATaskerEnv.vector_step in the episodic device mode, and the GPU-side duration of the step graph + copies.
  python tools/c3_host_probe.py [--envs 4096] [--steps 300] [--obs-dtype float32]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--obs-dtype", default="float32")
    a = ap.parse_args()
    import torch
    from ssa_gym_b200 import env_config
    from ssa_gym_b200.catalog import synthetic_catalog
    from ssa_gym_b200.transformations import gcrs2irts_matrix_approx, time_table
    from ssa_gym_b200.vec_env import VecSSATaskerEnv
    cfg = dict(env_config)
    cfg["orbits"] = synthetic_catalog(20000, 0)
    cfg["trans_matrix"] = gcrs2irts_matrix_approx(time_table(cfg["t_0"], cfg["time_step"], cfg["steps"]))
    cfg["obs_dtype"] = a.obs_dtype
    env = VecSSATaskerEnv(cfg, a.envs, seeds=list(range(a.envs)), rng="device")
    f32 = env.obs_dtype == np.float32
    for _ in range(10):
        env.vector_step(env.greedy_actions())
    seg = {k: 0.0 for k in ("greedy_actions", "check+copy actions", "rollout_step (issue)", "sync (GPU + copies)", "post")}
    pc = time.perf_counter
    gpu_ms = 0.0
    t_all = pc()
    for s in range(a.steps):
        t0 = pc()
        act = env.greedy_actions()
        t1 = pc()
        actions = np.ascontiguousarray(act, dtype=np.int32).reshape(env.E)
        assert np.all((actions >= 0) & (actions < env.m))
        env._io["actions"][:] = actions
        t2 = pc()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.ukf.rollout_step(env.auto_reset, obs_f32=f32)
        e1.record()
        t3 = pc()
        env.ukf.sync()
        torch.cuda.synchronize()
        t4 = pc()
        io = env._io
        dones = io["done"].astype(bool)
        rewards = io["reward"].copy()
        env.i += 1
        env.i[dones] = 0
        env.episodes[dones] += 1
        env.obs = env._format_obs(io["obs_f32" if f32 else "obs"].reshape(env.E, env.m * 12))
        t5 = pc()
        gpu_ms += e0.elapsed_time(e1)
        for k, d in zip(seg, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            seg[k] += d
    total = pc() - t_all
    print(f"E={a.envs} obs={a.obs_dtype}: {total / a.steps * 1e3:.4f} ms per step wall (with the probe's events), GPU graph + copies {gpu_ms / a.steps:.4f} ms")
    for k, v in seg.items():
        print(f"  {k:24s} {v / a.steps * 1e6:8.1f} us")
    env.close()


if __name__ == "__main__":
    main()
