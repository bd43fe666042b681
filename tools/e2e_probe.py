#!/usr/bin/env python
"""Where the end-to-end loop of bench.py loses time against the device-resident loop: variants of the per-step result
read-back around BatchedUKF.step_pinned.  python tools/e2e_probe.py [--sizes 1000000,125000] [--steps 100]"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,125000")
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    import torch
    import bench
    from ssa_gym_b200 import _lib as F
    from ssa_gym_b200.ukf import BatchedUKF
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE | F.STEP_CATALOG_STATS
    M = np.array(bench.CEL2TER06AXY)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    side = torch.cuda.Stream()
    for n in [int(v) for v in a.sizes.split(",")]:
        cat, x, P0, zn = bench.workload_inputs(n, 0, 2)
        cfg = bench.make_cfg(n)
        ukf = BatchedUKF(n_envs=1, m=n, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                         obs_lla=[np.radians(38.828198), np.radians(-77.305352), 20.0], obs_limit_rad=np.radians(-90.0))
        ukf.reset(cat, x, P0, stream=sp)
        ukf.catalog_stats(index_offset=0, stream=sp)
        stats = ukf.torch_view(F.F_CATALOG_STATS)
        gathered = torch.zeros(5, dtype=torch.float64, device="cuda")
        ghost = torch.empty(5, dtype=torch.float64).pin_memory()
        io = ukf.host_io()
        for b in range(2):
            io[b]["z_noise"][:] = zn[b]
            io[b]["M"][:] = M.reshape(9)

        def run(variant):
            for _ in range(4):
                ukf.step_pinned(flags | F.STEP_NO_D2H, stream=sp)
            ukf.host_join(stream=sp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for s in range(a.steps):
                ukf.step_pinned(flags | F.STEP_NO_D2H, stream=sp)
                if variant == "inline":
                    gathered.copy_(stats, non_blocking=True)
                    ghost.copy_(gathered, non_blocking=True)
                elif variant == "inline_d2h_only":
                    ghost.copy_(stats, non_blocking=True)
                elif variant == "side":
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    with torch.cuda.stream(side):
                        side.wait_event(ev)
                        ghost.copy_(stats, non_blocking=True)
            ukf.host_join(stream=sp)
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.steps

        def plain():
            ukf.upload(F.F_Z_NOISE, zn[0], stream=sp)
            for _ in range(4):
                ukf.step(M, flags, stream=sp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for s in range(a.steps):
                ukf.step(M, flags, stream=sp)
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.steps

        print(f"N={n}: device-resident step (no flush) {plain():.4f} ms", flush=True)
        for v in ("none", "inline", "inline_d2h_only", "side", "none"):
            print(f"N={n}: step_pinned + read-back '{v}': {run(v):.4f} ms", flush=True)
        ukf.close()


if __name__ == "__main__":
    main()
