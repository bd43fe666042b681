"""CPU: the N > 1 host logic (contiguous sharding, the reward/observation gather, catalog-level reductions) with
world_size = 2 over gloo.  No GPU involved."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H  # noqa: F401
from ssa_gym_b200 import dist as D


def test_shard_bounds_partition():
    for total in (1, 7, 10, 4096, 20000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            b = [D.shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == D.shard_sizes(total, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, E, m, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_bounds(E, world, rank)
    rng = np.random.RandomState(0)
    reward_all = rng.uniform(size=E)
    obs_all = rng.normal(size=(E, m * 12))
    r = D.gather_rows(torch.from_numpy(reward_all[lo:hi].copy()), E)
    o = D.gather_rows(torch.from_numpy(obs_all[lo:hi].copy()), E)
    ok = np.array_equal(r.numpy(), reward_all) and np.array_equal(o.numpy(), obs_all)
    # catalog-level reductions over object shards (C4): max dpos, trinary mean, argmax trace with first-max ties
    N = 1001
    dpos = np.random.RandomState(1).uniform(0, 2e7, N)
    trace = np.random.RandomState(2).randint(0, 50, N).astype(float)   # many exact ties
    lo2, hi2 = D.shard_bounds(N, world, rank)
    d, t = dpos[lo2:hi2], trace[lo2:hi2]
    st = D.reduce_catalog_stats(d.max(), float(((d < 1e4) * 1 + (d < 1e7) * 1).sum()), len(d), t.max(), lo2 + int(np.argmax(t)))
    ok = ok and st["max_delta_pos"] == dpos.max() and st["argmax_trace"] == int(np.argmax(trace))
    ok = ok and st["trinary_reward"] == np.mean(((dpos < 1e4) * 1 + (dpos < 1e7) * 1)) / 2
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_and_reductions_world2():
    world, E, m = 2, 13, 10   # ragged: shards of 7 and 6 environments
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, E, m, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(out[r] for r in range(world))
