"""-m gpu: parity at BASELINE.json's FULL sizes through size-independent properties.

C4 is a 1 000 000-object catalog — too many objects for the CPU twin / oracle to step in seconds — but no object of the
path reads another object's state, so the result of object i inside the 1 M batch must be, bit for bit, the
result of object i in any other batch that holds it.  That gives three checks at the full size:

* a random sample of the catalog (plus the first and last tiles) stepped by the host twin as its own small batch
  == the same objects inside the 1 M-object GPU run, every output, every bit;
* one 125 000-object shard (the block an 8-GPU job gives rank 2) stepped alone on the GPU == its slice of the
  monolithic run (what makes the sharding of DESIGN.md section 7 exact);
* whole-catalog invariants: no filter failed, every covariance diagonal positive, the device reduction of the reward
  terms == numpy on the downloaded arrays.
"""
import os
import sys

import numpy as np
import pytest

import helpers as H
from ssa_gym_b200 import _lib
from ssa_gym_b200.ukf import BatchedUKF

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pytestmark = pytest.mark.gpu
F = _lib
FULL = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE | F.STEP_RECORD
N_FULL = 1_000_000
STEPS = 3


def _handle(n, obs_limit_deg):
    cfg = H.make_cfg(n, obs_limit_deg=obs_limit_deg)
    return BatchedUKF(n_envs=1, m=n, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                      obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                      obs_limit_rad=np.radians(obs_limit_deg))


def _run(n, cat, x, P0, zn, obs_limit_deg):
    ukf = _handle(n, obs_limit_deg)
    ukf.reset(cat, x, P0)
    for s in range(STEPS):
        ukf.upload(F.F_Z_NOISE, zn[s])
        ukf.step(H.CEL2TER06AXY, FULL)
    ukf.sync()
    return ukf


FIELDS = (F.F_X_TRUE, F.F_X_FILTER, F.F_OBS, F.F_DELTA_POS, F.F_DELTA_VEL, F.F_SIGMA_POS, F.F_SIGMA_VEL, F.F_TRACE,
          F.F_STATUS, F.F_INFLATIONS, F.F_VISIBLE, F.F_UPDATED, F.F_Z_TRUE)


@pytest.mark.parametrize("obs_limit_deg", [15.0])
def test_c4_full_size_sample_equals_twin_and_shard_equals_whole(obs_limit_deg):
    """15 degree elevation mask: about a third of the catalog is visible, so updated and predicted-only objects are both
    in every tile."""
    cat, x, P0, zn = bench.workload_inputs(N_FULL, 0, STEPS)
    whole = _run(N_FULL, cat, x, P0, zn, obs_limit_deg)
    got = {f: whole.download(f) for f in FIELDS}
    P_whole = H.pack_P(whole.download(F.F_P_FILTER))
    y_w, S_w = whole.download(F.F_Y), whole.download(F.F_S)

    # ---- whole-catalog invariants ----
    assert (got[F.F_STATUS] & 1).sum() == 0
    assert np.isfinite(got[F.F_OBS]).all() and (got[F.F_OBS][:, 6:] > 0).all()
    vis = got[F.F_VISIBLE].astype(bool)
    assert 0.05 < vis.mean() < 0.95 and np.array_equal(got[F.F_UPDATED], got[F.F_VISIBLE])
    whole.catalog_stats(index_offset=0)
    st5 = whole.download(F.F_CATALOG_STATS)
    dpos, tr = got[F.F_DELTA_POS], got[F.F_TRACE]
    assert st5[0] == dpos.max() and st5[2] == N_FULL and st5[3] == tr.max() and st5[4] == int(np.argmax(tr))
    assert st5[1] == float(((dpos < 1e4).astype(int) + (dpos < 1e7).astype(int)).sum())
    whole.close()

    # ---- a sample of the catalog through the host twin, as its own batch ----
    rng = np.random.RandomState(77)
    idx = np.unique(np.concatenate([np.arange(64), np.arange(N_FULL - 64, N_FULL), rng.randint(0, N_FULL, 4000)]))
    cfg = H.make_cfg(len(idx), obs_limit_deg=obs_limit_deg)
    st = H.HostState(cat[idx], x[idx], P0)
    for s in range(STEPS):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, FULL, z_noise=np.ascontiguousarray(zn[s][idx]))
    twin = {F.F_X_TRUE: st.x_true, F.F_X_FILTER: st.x, F.F_OBS: st.obs, F.F_DELTA_POS: st.dpos, F.F_DELTA_VEL: st.dvel,
            F.F_SIGMA_POS: st.spos, F.F_SIGMA_VEL: st.svel, F.F_TRACE: st.trace, F.F_STATUS: st.status,
            F.F_INFLATIONS: st.infl, F.F_VISIBLE: st.visible, F.F_UPDATED: st.updated}
    for f, ref in twin.items():
        assert H.bits_equal(got[f][idx], ref), f
    assert H.bits_equal(P_whole[idx], H.pack_P(st.P))
    upd = st.updated.astype(bool)
    assert upd.any() and (~upd).any()
    assert H.bits_equal(got[F.F_Z_TRUE][idx], st.z_true)
    assert H.bits_equal(y_w[idx][upd], st.y[upd]) and H.bits_equal(S_w[idx][upd], st.S[upd])

    # ---- rank 2's shard of an 8-GPU job, stepped alone ----
    lo, hi = 250_000, 375_000
    shard = _run(hi - lo, cat[lo:hi], x[lo:hi], P0, zn[:, lo:hi], obs_limit_deg)
    for f in FIELDS:
        assert H.bits_equal(shard.download(f), got[f][lo:hi]), f
    assert H.bits_equal(H.pack_P(shard.download(F.F_P_FILTER)), P_whole[lo:hi])
    shard.close()


def test_c3_full_size_vector_step_invariants():
    """C3 at its full size (4 096 environments x 10 objects, RL mode): every environment of the vectorised step is
    bit-identical to the same environment stepped in a batch of 8 environments (environments are independent)."""
    E, m, steps = 4096, 10, 3
    N = E * m
    cat, x, P0, zn = H.c2_inputs(20000, steps)
    from ssa_gym_b200.catalog import tiled_catalog
    cat = tiled_catalog(N, cat, seed=11)
    x = cat + np.random.RandomState(3).normal(size=(N, 6)) * np.array([1e5] * 3 + [1e2] * 3)
    zn = np.random.RandomState(4).normal(size=(steps, N, 3)) * np.array([H.arcsec2rad, H.arcsec2rad, 1e3])
    actions = np.random.RandomState(5).randint(0, m, size=(steps, E)).astype(np.int32)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ACT | F.STEP_EPILOGUE

    def run(e0, e1):
        n_e = e1 - e0
        cfg = H.make_cfg(n_e * m, E=n_e, m=m)
        ukf = BatchedUKF(n_envs=n_e, m=m, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                         obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                         obs_limit_rad=np.radians(-90.0))
        sl = slice(e0 * m, e1 * m)
        ukf.reset(cat[sl], x[sl], P0)
        for s in range(steps):
            ukf.upload(F.F_ACTIONS, actions[s, e0:e1])
            ukf.upload(F.F_Z_NOISE, zn[s][sl])
            ukf.step(H.CEL2TER06AXY, flags)
        ukf.sync()
        out = {f: ukf.download(f) for f in (F.F_X_TRUE, F.F_X_FILTER, F.F_OBS, F.F_TRACE, F.F_STATUS, F.F_UPDATED)}
        out["P"] = H.pack_P(ukf.download(F.F_P_FILTER))
        ukf.close()
        return out

    whole = run(0, E)
    assert whole[F.F_UPDATED].reshape(E, m).sum(axis=1).max() == 1   # one tasked object per environment
    assert (whole[F.F_STATUS] & 1).sum() == 0
    for e0 in (0, 1000, E - 8):
        part = run(e0, e0 + 8)
        sl = slice(e0 * m, (e0 + 8) * m)
        for k, v in part.items():
            assert H.bits_equal(v, whole[k][sl]), (e0, k)
