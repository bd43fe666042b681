"""-m gpu: the sm_100a kernels against the host twin — BIT-EXACT, through the C ABI.

The twin (tests/twin/twin.cpp) is the host build of the product's own __host__ __device__ headers; the
independent oracle checks live in test_gpu_vs_oracle.py / test_twin_vs_oracle.py.  Bit-exactness is the
bar for every output (float and integer) at BASELINE.json's full C2 size (20 000 objects).
"""
import numpy as np
import pytest

import helpers as H
from ssa_gym_b200 import _lib
from ssa_gym_b200.ukf import BatchedUKF

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tile", "tile2", "fused", "split", "team"], autouse=False)
def kernel(request, monkeypatch):
    """Every device implementation of the step: the tile kernels (default: sigma sets in shared memory), the split
    five-kernel pipeline (SSA_UKF_KERNEL=split) and the fused 16-lane team kernel (SSA_UKF_KERNEL=team).  The handle
    reads the variable at creation."""
    monkeypatch.setenv("SSA_UKF_KERNEL", request.param)
    return request.param

F = _lib
FULL = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE | F.STEP_RECORD


def _rng_args(rng, n):
    return np.concatenate([rng.uniform(-7, 7, n), rng.uniform(-1e3, 1e3, n), rng.standard_normal(n) * 1e-3,
                           np.pi * np.arange(-8, 9) / 2, [0.0, -0.0, 1.0, -1.0, 0.5, -0.5]])


@pytest.mark.parametrize("op", ["sin", "cos", "tan", "atan", "exp", "sinh", "cosh", "tanh", "asinh"])
def test_math_unary_bitexact(op):
    x = _rng_args(np.random.default_rng(1), 20000)
    assert H.bits_equal(H.lib_math("gpu", op, x), H.lib_math("twin", op, x))


@pytest.mark.parametrize("op,lo,hi", [("asin", -1, 1), ("acos", -1, 1), ("atanh", -0.999, 0.999), ("log", 1e-300, 1e10),
                                      ("acosh", 1, 1e6), ("pow23", 1e-9, 1e9)])
def test_math_domain_bitexact(op, lo, hi):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(lo, hi, 20000), [lo, hi], 1 - 10.0 ** rng.uniform(-16, 0, 500) if op in ("asin", "acos") else []])
    assert H.bits_equal(H.lib_math("gpu", op, x), H.lib_math("twin", op, x))


def test_math_binary_bitexact():
    rng = np.random.default_rng(3)
    a = rng.standard_normal(50000) * 10.0 ** rng.uniform(-3, 8, 50000)
    b = rng.standard_normal(50000) * 10.0 ** rng.uniform(-3, 8, 50000)
    assert H.bits_equal(H.lib_math("gpu", "atan2", a, b), H.lib_math("twin", "atan2", a, b))
    m = np.full(a.size, 2 * np.pi)
    assert H.bits_equal(H.lib_math("gpu", "pymod", a, m), H.lib_math("twin", "pymod", a, m))
    assert np.array_equal(H.lib_math("gpu", "pymod", a, m), a % (2 * np.pi))  # == numpy's python-mod


def test_fx_bitexact_catalog():
    cat, x, _, _ = H.c2_inputs(20000)
    for dt in (20.0, 600.0, 86400.0):
        for states in (cat, x):
            g, ge = H.lib_fx("gpu", states, dt)
            t, te = H.lib_fx("twin", states, dt)
            assert H.bits_equal(g, t) and np.array_equal(ge, te)


def test_hx_bitexact():
    cat, x, _, _ = H.c2_inputs(20000)
    cfg = H.make_cfg(20000)
    oi, T = np.array(cfg.obs_itrs), np.array(cfg.T)
    assert H.bits_equal(H.lib_hx("gpu", x, H.CEL2TER06AXY, oi, T), H.lib_hx("twin", x, H.CEL2TER06AXY, oi, T))


def _gpu_run(cfg_kwargs, cat, x, P0, zn, flags_seq, actions=None, N=None, E=None, m=None):
    N = len(cat)
    ukf = BatchedUKF(n_envs=E or 1, m=m or N, dt=cfg_kwargs.get("dt", 20.0),
                     Q=np.array(H.make_cfg(N, **cfg_kwargs).Q).reshape(6, 6),
                     R=np.array(H.make_cfg(N, **cfg_kwargs).R).reshape(3, 3),
                     obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                     obs_limit_rad=np.radians(cfg_kwargs.get("obs_limit_deg", -90.0)),
                     obs_type=cfg_kwargs.get("obs_type", "aer"), resample_after_predict=cfg_kwargs.get("resample", True))
    ukf.reset(cat, x, P0)
    for s, flags in enumerate(flags_seq):
        if actions is not None:
            ukf.upload(F.F_ACTIONS, actions[s])
        ukf.upload(F.F_Z_NOISE, zn[s])
        ukf.step(H.CEL2TER06AXY, flags)
    ukf.sync()
    return ukf


def _compare_all(ukf, st, check_update_outputs=True):
    D = ukf.download
    assert H.bits_equal(D(F.F_X_TRUE), st.x_true)
    assert H.bits_equal(D(F.F_X_FILTER), st.x)
    assert H.bits_equal(H.pack_P(D(F.F_P_FILTER)), H.pack_P(st.P))
    assert np.array_equal(D(F.F_STATUS), st.status)
    assert np.array_equal(D(F.F_INFLATIONS), st.infl)
    assert H.bits_equal(D(F.F_OBS), st.obs)
    for f, a in ((F.F_DELTA_POS, st.dpos), (F.F_DELTA_VEL, st.dvel), (F.F_SIGMA_POS, st.spos), (F.F_SIGMA_VEL, st.svel),
                 (F.F_TRACE, st.trace)):
        assert H.bits_equal(D(f), a)
    assert np.array_equal(D(F.F_VISIBLE), st.visible)
    assert np.array_equal(D(F.F_UPDATED), st.updated)
    if check_update_outputs:
        upd = st.updated.astype(bool)
        assert H.bits_equal(D(F.F_Y)[upd], st.y[upd])
        assert H.bits_equal(D(F.F_S)[upd], st.S[upd])
        assert H.bits_equal(D(F.F_SIGMAS_H)[upd], st.sigmas_h[upd])


@pytest.mark.parametrize("obs_limit_deg", [-90.0, 15.0])
def test_fused_step_bitexact_c2(obs_limit_deg, kernel):
    """C2: 20 000 objects, fused truth+predict+update+epilogue, 4 steps, all outputs bit-equal to the twin."""
    N, steps = 20000, 4
    cat, x, P0, zn = H.c2_inputs(N, steps)
    kw = dict(obs_limit_deg=obs_limit_deg)
    cfg = H.make_cfg(N, **kw)
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, FULL, z_noise=zn[s])
    ukf = _gpu_run(kw, cat, x, P0, zn, [FULL] * steps)
    _compare_all(ukf, st)
    ukf.close()


def test_split_predict_update_equals_fused(kernel):
    """predict() then update() as two launches (the reference's call structure) == the fused launch."""
    N, steps = 4096, 3
    cat, x, P0, zn = H.c2_inputs(N, steps)
    cfg = H.make_cfg(N)
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, FULL, z_noise=zn[s])
    seq = []
    for s in range(steps):
        seq += [F.STEP_TRUTH | F.STEP_PREDICT, F.STEP_UPDATE_ALL | F.STEP_EPILOGUE | F.STEP_RECORD]
    zn2 = np.repeat(zn, 2, axis=0)
    ukf = _gpu_run({}, cat, x, P0, zn2, seq)
    _compare_all(ukf, st)
    ukf.close()


def test_rl_mode_actions_bitexact(kernel):
    """E envs x m objects, one tasked object per env (SS2:292-315), ragged N (not a multiple of 8/32)."""
    E, m, steps = 1037, 10, 5
    N = E * m
    cat, x, P0, zn = H.c2_inputs(N, steps)
    rng = np.random.RandomState(5)
    actions = rng.randint(0, m, size=(steps, E)).astype(np.int32)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ACT | F.STEP_EPILOGUE | F.STEP_RECORD
    cfg = H.make_cfg(N, E=E, m=m, obs_limit_deg=10.0)
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, flags, actions=actions[s], z_noise=zn[s])
    ukf = _gpu_run(dict(obs_limit_deg=10.0), cat, x, P0, zn, [flags] * steps, actions=actions, E=E, m=m)
    _compare_all(ukf, st, check_update_outputs=False)
    upd = st.updated.astype(bool)
    assert 0 < upd.sum() < E  # some tasked objects are below the elevation mask
    assert H.bits_equal(ukf.download(F.F_Y)[upd], st.y[upd])
    ukf.close()


@pytest.mark.parametrize("act", ["tile", "split"])
@pytest.mark.parametrize("chunk", [None, "1000"])
def test_rl_mode_tile_update_equals_split_update(act, chunk, monkeypatch):
    """RL mode runs the update of the tasked objects through k_update_tile (the other objects of a tile only get the truth
    measurement and the epilogue); SSA_UKF_ACT=split keeps k_hx + k_update.  Both equal the twin in every output, including
    the `updated` flags, environments that task nothing (action -1) and environment-straddling chunks."""
    if act == "split":
        monkeypatch.setenv("SSA_UKF_ACT", "split")
    if chunk:
        monkeypatch.setenv("SSA_UKF_CHUNK", chunk)
    E, m, steps = 517, 7, 4
    N = E * m
    cat, x, P0, zn = H.c2_inputs(N, steps)
    rng = np.random.RandomState(6)
    actions = rng.randint(-1, m, size=(steps, E)).astype(np.int32)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ACT | F.STEP_EPILOGUE | F.STEP_RECORD
    cfg = H.make_cfg(N, E=E, m=m, obs_limit_deg=10.0)
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, flags, actions=actions[s], z_noise=zn[s])
    ukf = _gpu_run(dict(obs_limit_deg=10.0), cat, x, P0, zn, [flags] * steps, actions=actions, E=E, m=m)
    _compare_all(ukf, st, check_update_outputs=False)
    upd = st.updated.astype(bool)
    assert np.array_equal(ukf.download(F.F_UPDATED).astype(bool), upd) and 0 < upd.sum() < E
    assert H.bits_equal(ukf.download(F.F_Y)[upd], st.y[upd]) and H.bits_equal(ukf.download(F.F_S)[upd], st.S[upd])
    assert H.bits_equal(ukf.download(F.F_Z_TRUE)[upd], st.z_true[upd])
    ukf.close()
    # stand-alone update of the tasked objects (no predict in the step): factor, then update
    st2 = H.HostState(cat, x, P0)
    f2 = F.STEP_UPDATE_ACT | F.STEP_EPILOGUE | F.STEP_RECORD
    H.cpu_step("twin", cfg, st2, H.CEL2TER06AXY, f2, actions=actions[0], z_noise=zn[0])
    ukf = _gpu_run(dict(obs_limit_deg=10.0), cat, x, P0, zn, [f2], actions=actions[:1], E=E, m=m)
    _compare_all(ukf, st2, check_update_outputs=False)
    assert np.array_equal(ukf.download(F.F_UPDATED).astype(bool), st2.updated.astype(bool))
    ukf.close()


@pytest.mark.parametrize("obs_type,resample", [("xyz", True), ("aer", False), ("xyz", False)])
def test_variants_bitexact(obs_type, resample, kernel):
    N, steps = 2048, 3
    cat, x, P0, zn = H.c2_inputs(N, steps)
    R = np.diag([125.0] * 3) if obs_type == "xyz" else None
    kw = dict(obs_type=obs_type, resample=resample, R=R)
    cfg = H.make_cfg(N, **kw)
    if obs_type == "xyz":
        zn = np.random.RandomState(7).normal(size=(steps, N, 3)) * 10.0
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, FULL, z_noise=zn[s])
    ukf = _gpu_run(kw, cat, x, P0, zn, [FULL] * steps)
    _compare_all(ukf, st)
    ukf.close()


def test_pipelined_host_step_equals_twin(kernel):
    """ssa_ukf_step_host: double-buffered, asynchronous H2D / kernels / D2H.  Six pipelined steps with per-step
    host result buffers must deliver exactly what the twin computes step by step."""
    N, steps = 3000, 6
    cat, x, P0, zn = H.c2_inputs(N, steps)
    cfg = H.make_cfg(N)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE
    ukf = BatchedUKF(n_envs=1, m=N, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                     obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                     obs_limit_rad=np.radians(-90.0))
    ukf.reset(cat, x, P0)
    obs = np.zeros((steps, N, 12)); dpos = np.zeros((steps, N)); stat = np.zeros((steps, N), np.int32)
    zn = np.ascontiguousarray(zn)
    for s in range(steps):
        ukf.step_host(H.CEL2TER06AXY, flags, z_noise=zn[s], obs_out=obs[s], dpos_out=dpos[s], status_out=stat[s])
    ukf.host_join()
    ukf.sync()
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, flags, z_noise=zn[s])
        assert H.bits_equal(obs[s], st.obs) and H.bits_equal(dpos[s], st.dpos) and np.array_equal(stat[s], st.status), s
    assert H.bits_equal(ukf.download(F.F_X_FILTER), st.x)
    ukf.close()


@pytest.mark.parametrize("graph", ["1", "0"])
def test_pinned_pipelined_step_equals_twin(graph, monkeypatch):
    """ssa_ukf_step_pinned: the handle's pinned I/O blocks, ONE copy each way, the kernel chain as a captured CUDA
    graph (graph=1) or plain launches (graph=0), trans_matrix read from the uploaded block.  Catalog mode (update
    all, a different trans_matrix every step) and RL mode (actions through the pinned block) against the twin."""
    monkeypatch.setenv("SSA_UKF_GRAPH", graph)
    N, steps = 3000, 7
    cat, x, P0, zn = H.c2_inputs(N, steps)
    cfg = H.make_cfg(N)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE
    rng = np.random.RandomState(3)
    Ms = [H.CEL2TER06AXY @ _rotz(1e-3 * s) for s in range(steps)]
    ukf = BatchedUKF(n_envs=1, m=N, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                     obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                     obs_limit_rad=np.radians(-90.0))
    ukf.reset(cat, x, P0)
    io = ukf.host_io()
    st = H.HostState(cat, x, P0)
    pending = None
    for s in range(steps + 1):
        if s < steps:
            b = ukf.next_parity
            if s >= 2:  # this parity's previous results are about to be overwritten: consume them first
                ukf.host_join(); ukf.sync()
            io[b]["z_noise"][:] = zn[s]
            io[b]["M"][:] = Ms[s].reshape(9)
            assert ukf.step_pinned(flags) == b
        if s >= 1:  # check step s-1
            ukf.host_join(); ukf.sync()
            bp = (s - 1) & 1
            H.cpu_step("twin", cfg, st, Ms[s - 1], flags, z_noise=zn[s - 1])
            assert H.bits_equal(io[bp]["obs"], st.obs) and H.bits_equal(io[bp]["delta_pos"], st.dpos), s
            assert np.array_equal(io[bp]["status"], st.status), s
    assert H.bits_equal(ukf.download(F.F_X_FILTER), st.x)
    ukf.close()
    # RL mode: E envs x m objects, update only the tasked object of every env
    E, m = 300, 7
    N = E * m
    cat, x, P0, zn = H.c2_inputs(N, 4)
    cfg = H.make_cfg(N, E=E, m=m)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ACT | F.STEP_EPILOGUE
    ukf = BatchedUKF(n_envs=E, m=m, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                     obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                     obs_limit_rad=np.radians(-90.0))
    ukf.reset(cat, x, P0)
    io = ukf.host_io()
    st = H.HostState(cat, x, P0)
    for s in range(4):
        act = rng.randint(0, m, size=E).astype(np.int32)
        b = ukf.next_parity
        io[b]["z_noise"][:] = zn[s]
        io[b]["M"][:] = H.CEL2TER06AXY.reshape(9)
        io[b]["actions"][:] = act
        ukf.step_pinned(flags)
        ukf.host_join(); ukf.sync()
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, flags, actions=act, z_noise=zn[s])
        assert H.bits_equal(io[b]["obs"], st.obs) and np.array_equal(io[b]["status"], st.status), s
    ukf.close()


def _rotz(a):
    c, s_ = np.cos(a), np.sin(a)
    return np.array([[c, s_, 0.0], [-s_, c, 0.0], [0.0, 0.0, 1.0]])


@pytest.mark.parametrize("impl", ["tile", "tile2", "fused", "split"])
@pytest.mark.parametrize("chunk", ["640", "4000"])
def test_chunked_execution_bitexact(chunk, impl, monkeypatch):
    """Large batches run in L2-sized chunks (SSA_UKF_CHUNK objects per chunk, env-aligned in RL mode).  Forcing
    tiny chunks (ragged last chunk) must not change a single bit, in catalog mode and in RL mode."""
    monkeypatch.setenv("SSA_UKF_KERNEL", impl)
    monkeypatch.setenv("SSA_UKF_CHUNK", chunk)
    N, steps = 9001, 3
    cat, x, P0, zn = H.c2_inputs(N, steps)
    cfg = H.make_cfg(N, obs_limit_deg=5.0)
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, FULL, z_noise=zn[s])
    ukf = _gpu_run(dict(obs_limit_deg=5.0), cat, x, P0, zn, [FULL] * steps)
    _compare_all(ukf, st)
    ukf.close()
    E, m = 903, 7
    N = E * m
    cat, x, P0, zn = H.c2_inputs(N, steps)
    actions = np.random.RandomState(5).randint(0, m, size=(steps, E)).astype(np.int32)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ACT | F.STEP_EPILOGUE | F.STEP_RECORD
    cfg = H.make_cfg(N, E=E, m=m)
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, flags, actions=actions[s], z_noise=zn[s])
    ukf = _gpu_run({}, cat, x, P0, zn, [flags] * steps, actions=actions, E=E, m=m)
    _compare_all(ukf, st, check_update_outputs=False)
    ukf.close()


def test_diagnostics_kernel_equals_twin():
    """ssa_ukf_diagnostics (NEES, NIS, innovation-bound flags) on the device == the twin, bit for bit."""
    N, steps = 5000, 3
    cat, x, P0, zn = H.c2_inputs(N, steps)
    cfg = H.make_cfg(N, obs_limit_deg=5.0)
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE | F.STEP_RECORD
    st = H.HostState(cat, x, P0)
    for s in range(steps):
        H.cpu_step("twin", cfg, st, H.CEL2TER06AXY, flags, z_noise=zn[s])
    ukf = _gpu_run(dict(obs_limit_deg=5.0), cat, x, P0, zn, [flags] * steps)
    nees_g, nis_g, fl_g = ukf.diagnostics()
    nees = np.zeros(N); nis = np.zeros(N); fl = np.zeros(N, np.uint8)
    H.twin().twin_diagnostics(N, H.p(st.x_true), H.p(st.x), H.p(H.pack_P(st.P)), H.p(np.nan_to_num(st.y)),
                              H.p(np.nan_to_num(st.S)), H.p(st.updated), H.p(nees), H.p(nis), H.p(fl))
    assert st.updated.sum() > 0 and (st.updated == 0).sum() > 0
    assert H.bits_equal(nees_g, nees) and H.bits_equal(nis_g, nis) and np.array_equal(fl_g, fl)
    ukf.close()


def test_catalog_stats_reduction_exact():
    """ssa_ukf_catalog_stats (C4 reward terms of a shard: max delta_pos, trinary count sum, argmax trace with
    first-maximum-wins and a global index offset) against numpy on the downloaded arrays: integer and max work, exact."""
    N, steps = 70001, 2
    cat, x, P0, zn = H.c2_inputs(20000, steps)
    from ssa_gym_b200.catalog import tiled_catalog
    cat = tiled_catalog(N, cat, seed=5)
    x = cat + np.random.RandomState(0).normal(size=(N, 6)) * np.array([1e5] * 3 + [1e2] * 3)
    zn = np.random.RandomState(1).normal(size=(steps, N, 3)) * np.array([H.arcsec2rad, H.arcsec2rad, 1e3])
    ukf = _gpu_run({}, cat, x, P0, zn, [FULL] * steps)
    ukf.catalog_stats(index_offset=1000000)
    got = ukf.download(F.F_CATALOG_STATS)
    dpos, tr = ukf.download(F.F_DELTA_POS), ukf.download(F.F_TRACE)
    assert got[0] == dpos.max() and got[1] == float(((dpos < 1e4).astype(int) + (dpos < 1e7).astype(int)).sum())
    assert got[2] == N and got[3] == tr.max() and got[4] == 1000000 + int(np.argmax(tr))
    # ties: first maximum wins
    tv = ukf.torch_view(F.F_TRACE)
    tv[123] = tv[60000] = float(tr.max()) * 2
    ukf.catalog_stats(index_offset=0)
    assert ukf.download(F.F_CATALOG_STATS)[4] == 123
    ukf.close()


def test_pinned_step_reward_terms_per_parity():
    """ssa_ukf_step_pinned with SSA_STEP_CATALOG_STATS: each call leaves its shard reward terms in the device slot and the
    pinned host mirror of ITS parity (ssa_ukf_host_stats), copied on the download stream — equal to numpy on the same
    call's downloaded outputs; the slot of the other parity still holds the previous step's terms."""
    N, steps = 9001, 5
    cat, x, P0, zn = H.c2_inputs(N, steps)
    cfg = H.make_cfg(N)
    ukf = BatchedUKF(n_envs=1, m=N, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                     obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                     obs_limit_rad=np.radians(-90.0))
    ukf.reset(cat, x, P0)
    ukf.catalog_stats(index_offset=700)
    io = ukf.host_io()
    flags = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE | F.STEP_CATALOG_STATS
    prev = {}
    for s in range(steps):
        b = ukf.next_parity
        io[b]["z_noise"][:] = zn[s]
        io[b]["M"][:] = np.asarray(H.CEL2TER06AXY).reshape(9)
        assert ukf.step_pinned(flags) == b
        ukf.host_join()
        ukf.sync()
        host, _ = ukf.host_stats(b)
        dpos = io[b]["delta_pos"]
        assert host[0] == dpos.max() and host[2] == N
        assert host[1] == float(((dpos < 1e4).astype(int) + (dpos < 1e7).astype(int)).sum())
        obs = io[b]["obs"]
        trace = ((((obs[:, 6] + obs[:, 7]) + obs[:, 8]) + obs[:, 9]) + obs[:, 10]) + obs[:, 11]
        assert host[3] == trace.max() and host[4] == 700 + int(np.argmax(trace))
        dev = ukf.host_stats_torch(b).cpu().numpy()
        assert np.array_equal(dev, host)
        if (b ^ 1) in prev:
            assert np.array_equal(ukf.host_stats(b ^ 1)[0], prev[b ^ 1])
        prev[b] = np.array(host)
    ukf.close()


@pytest.mark.parametrize("graph", ["1", "0"])
def test_step_graph_cache_many_flag_combinations(graph, monkeypatch):
    """ssa_ukf_step replays the kernel chain as a cached CUDA graph per flag combination (4 slots, the k_hx node
    re-parameterised with every step's trans_matrix).  Seven different flag combinations interleaved (evictions), a
    different trans_matrix every step: bit-exact against the twin, and identical with plain launches (graph=0)."""
    monkeypatch.setenv("SSA_UKF_STEP_GRAPH", graph)
    N = 2500
    combos = [FULL, F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_EPILOGUE, F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE,
              F.STEP_UPDATE_ALL | F.STEP_EPILOGUE, F.STEP_TRUTH | F.STEP_EPILOGUE, F.STEP_EPILOGUE,
              F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE | F.STEP_RECORD]
    seq = [combos[i % len(combos)] for i in range(16)] + [FULL, combos[1], FULL]
    cat, x, P0, zn = H.c2_inputs(N, len(seq))
    cfg = H.make_cfg(N, obs_limit_deg=10.0)
    Ms = [H.CEL2TER06AXY @ _rotz(2e-3 * s) for s in range(len(seq))]
    ukf = BatchedUKF(n_envs=1, m=N, dt=20.0, Q=np.array(cfg.Q).reshape(6, 6), R=np.array(cfg.R).reshape(3, 3),
                     obs_lla=[np.radians(H.OBSERVER_DEG[0]), np.radians(H.OBSERVER_DEG[1]), H.OBSERVER_DEG[2]],
                     obs_limit_rad=np.radians(10.0))
    ukf.reset(cat, x, P0)
    st = H.HostState(cat, x, P0)
    for s, flags in enumerate(seq):
        ukf.upload(F.F_Z_NOISE, zn[s])
        ukf.step(Ms[s], flags)
        H.cpu_step("twin", cfg, st, Ms[s], flags, z_noise=zn[s])
        if s % 5 == 4 or s == len(seq) - 1:
            assert H.bits_equal(ukf.download(F.F_OBS), st.obs) and H.bits_equal(ukf.download(F.F_X_TRUE), st.x_true), s
            assert np.array_equal(ukf.download(F.F_VISIBLE), st.visible) and np.array_equal(ukf.download(F.F_STATUS), st.status), s
    _compare_all(ukf, st)
    ukf.close()
