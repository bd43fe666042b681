#!/usr/bin/env python
"""Per-kernel CUDA-event times of one step (ssa_ukf_step_profile) and the plain step time for a list of batch sizes
and kernel implementations.  Development tool: python tools/kprof.py [--sizes 20000,125000,1000000] [--impls tile,split]"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="20000,125000,1000000")
    ap.add_argument("--impls", default="tile,split")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    import torch
    import bench
    from ssa_gym_b200 import _lib as F
    from ssa_gym_b200.ukf import BatchedUKF
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE
    M = np.array(bench.CEL2TER06AXY)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for n in [int(v) for v in a.sizes.split(",")]:
        cat, x, P0, zn = bench.workload_inputs(n, 0, 4)
        cfg = bench.make_cfg(n)
        for impl in a.impls.split(","):
            os.environ["SSA_UKF_KERNEL"] = impl
            ukf = BatchedUKF(n_envs=1, m=n, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                             obs_lla=[np.radians(38.828198), np.radians(-77.305352), 20.0], obs_limit_rad=np.radians(-90.0))
            ukf.reset(cat, x, P0, stream=sp)
            ukf.upload(F.F_Z_NOISE, zn[0], stream=sp)
            for _ in range(5):
                ukf.step(M, flags, stream=sp)
            torch.cuda.synchronize()
            ms = []
            for s in range(a.steps):
                flush.fill_(s & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ukf.step(M, flags, stream=sp)
                e1.record(stream)
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            km = []
            for s in range(10):
                flush.fill_(s & 0xFF)
                torch.cuda.synchronize()
                km.append(ukf.step_profile(M, flags, stream=sp))
            km = np.mean(np.array(km), axis=0)
            print(f"N={n:8d} impl={impl:6s} step_ms mean={np.mean(ms):.4f} min={np.min(ms):.4f}  obj/s={n / np.mean(ms) * 1e3:.3e}  "
                  f"kernels_ms factor={km[0]:.4f} fx|predict_tile={km[1]:.4f} ut|refactor={km[2]:.4f} hx={km[3]:.4f} update={km[4]:.4f}",
                  flush=True)
            ukf.close()


if __name__ == "__main__":
    main()
