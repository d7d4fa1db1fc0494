"""Parity tests proper: the CUDA path (through the C-ABI library) against the C oracle and against the
fixtures recorded from the unmodified reference (tests/golden, oracle/make_golden.py).  Need a B200: -m gpu.

Tolerances (BASELINE.json north_star): serving-BS indices, handover counts and outage counts bit-exact;
SINR within 1e-3 dB; reward within 1e-5 relative.  The float64 parity kernels are held to far tighter bounds
(1e-9 dB / 1e-9 relative); the fp32 kernels to the stated ones."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.make_golden import load_trace_10k, state_checksum

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SINR_TOL_DB = 1e-3        # north_star tolerance
REWARD_RTOL = 1e-5        # north_star tolerance
F64_SINR_TOL_DB = 1e-9    # what the float64 kernels actually achieve (libm vs CUDA log10/pow: ~1e-13)


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import drl_uav_cellularnet_b200 as p
    return p


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _np(t):
    return t.detach().cpu().numpy()


def _pad(u, n=52):
    """one tick's uniforms padded to the fixed [nUE + 3*nG] injection row"""
    out = np.zeros(n)
    out[:min(len(u), n)] = u[:n]
    return out


# ---------------------------------------------------------------------------------------------------------
def test_trace_replay_matches_reference_fixture(pkg, golden_dir):
    """BASELINE config 1: MobiEnvironment(4,40,100,"read_trace") replaying the regenerated trace for 2001
    step_test calls with a fixed action sequence and the recorded fading, against outputs recorded from the
    UNMODIFIED reference.  Decisions bit-exact; SINR / reward within the float64 bounds."""
    g = _load(golden_dir, "ref_trace_replay.npz")
    T = len(g["actions"])
    fade = np.random.RandomState(int(g["fade_seed"])).normal(0, 2, size=(T + 2, 40, 4))
    env = pkg.MobiEnvironment(4, 40, 100, "read_trace", trace=g["trace"], fading="injected")
    assert float(env.state.sum()) == 0.0                           # zeros until the first reset (mobile_env.py:107)
    env._b.ctor_pass(fading=fade[0][None])
    assert np.array_equal(env.channel.current_BS, g["ctor_cur"])
    s0 = env.reset(fading=fade[1])
    assert np.array_equal(env.channel.current_BS, g["reset_cur"])
    assert np.max(np.abs(env.channel.current_BS_sinr - g["reset_sinr"])) < F64_SINR_TOL_DB
    assert state_checksum(s0) == float(g["reset_state_chk"])
    sinr_rows = {int(t): i for i, t in enumerate(g["sinr_idx"])}
    max_sinr_err = max_rew_rel = 0.0
    for t in range(T):
        s, r, d, info = env.step_test(np.array([g["actions"][t]]), fading=fade[2 + t])   # shape-(1,) as main_test.py:73
        assert np.array_equal(env.channel.current_BS, g["cur"][t]), t
        assert int(env._b.n_out[0]) == g["n_out"][t], t
        assert int(env._b.n_ho[0]) == g["n_ho"][t], t
        assert np.array_equal(info.bs_loc[:, :2], g["bs_xy"][t]), t
        assert info.bs_actions.tolist() == g["digits"][t].tolist()
        assert d == bool(g["done"][t])
        assert info.step_n == t + 1
        assert state_checksum(s) == float(g["state_chk"][t]), t
        assert abs(float(env._b.mean_sinr[0]) - g["mean_sinr"][t]) < F64_SINR_TOL_DB
        assert abs(info.outage_fraction - g["n_out"][t] / 40.0) < 1e-15
        if abs(g["reward"][t]) > 0:
            max_rew_rel = max(max_rew_rel, abs(r - g["reward"][t]) / abs(g["reward"][t]))
        if t in sinr_rows:
            max_sinr_err = max(max_sinr_err,
                               float(np.max(np.abs(env.channel.current_BS_sinr - g["cur_sinr"][sinr_rows[t]]))))
    assert max_sinr_err < F64_SINR_TOL_DB, max_sinr_err
    assert max_rew_rel < 1e-9, max_rew_rel
    with pytest.raises(IndexError):                                # past the end of the trace (mobile_env.py:203)
        for _ in range(200):
            env.step_test(0, fading=fade[0])


def _strict_reward_report(got, want, name):
    """North star: reward within 1e-5 RELATIVE.  An fp32 SINR pass cannot hold that where the reward crosses zero
    (|d reward| = |d meanSINR| / 20, whatever the reward's size), so the strict criterion is counted and
    reported instead of being hidden behind an absolute floor: returns (violations, samples, worst absolute error,
    worst relative error among the samples with |reward| >= 0.1)."""
    got, want = np.asarray(got, dtype=np.float64).ravel(), np.asarray(want, dtype=np.float64).ravel()
    err = np.abs(got - want)
    viol = err > REWARD_RTOL * np.abs(want)
    big = np.abs(want) >= 0.1
    rep = {"what": name, "samples": int(err.size), "strict_rel_1e-5_violations": int(viol.sum()),
           "violations_with_abs_reward_ge_0.1": int((viol & big).sum()), "worst_abs_err": float(err.max()),
           "worst_rel_err_abs_reward_ge_0.1": float((err[big] / np.abs(want[big])).max()) if big.any() else 0.0,
           "largest_abs_reward_among_violations": float(np.abs(want[viol]).max()) if viol.any() else 0.0}
    _record(name, rep)
    return rep


def _record(name, rep):
    """parity statistics of this run -> gpurun_out/parity_stats.json (copied to profiles/ by hand after a GPU session)"""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        fn = os.path.join(d, "parity_stats.json")
        cur = {}
        if os.path.isfile(fn):
            with open(fn) as f:
                cur = json.load(f)
        cur[name] = rep
        with open(fn, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print("PARITY-STATS", name, rep)


@pytest.mark.parametrize("precision", ["fp32_guarded", "fp32"])
def test_trace_replay_fp32_full_fixture(pkg, golden_dir, precision):
    """The whole 2001-step reference fixture (config[0]) through the fp32 kernels, against outputs recorded from the
    UNMODIFIED reference.  fp32_guarded (the mode bench.py measures): serving BS, new-outage and handover counts, BS cells
    and the observation bit-exact at every step; serving SINR within 1e-3 dB; reward within 1e-5 relative wherever
    |reward| >= 0.1, strict violations counted (see _strict_reward_report).  Plain fp32: the same tolerances, and the
    number of steps whose serving cells differ from the reference is REPORTED (north star: 0; SURVEY H2 predicts about
    4e-6 flips per UE-step, i.e. 0.3 expected over these 80 k UE-steps)."""
    g = _load(golden_dir, "ref_trace_replay.npz")
    T = len(g["actions"])
    fade = np.random.RandomState(int(g["fade_seed"])).normal(0, 2, size=(T + 2, 40, 4))
    env = pkg.MobiEnvironment(4, 40, 100, "read_trace", trace=g["trace"], fading="injected", precision=precision)
    env._b.ctor_pass(fading=fade[0][None])
    guarded = precision == "fp32_guarded"
    if guarded:
        assert np.array_equal(env.channel.current_BS, g["ctor_cur"])
    s0 = env.reset(fading=fade[1])
    if guarded:
        assert np.array_equal(env.channel.current_BS, g["reset_cur"])
        assert state_checksum(s0) == float(g["reset_state_chk"])
    sinr_rows = {int(t): i for i, t in enumerate(g["sinr_idx"])}
    worst = worst_mean = 0.0
    first_div = -1
    rew, ref_rew = [], []
    for t in range(T):
        s, r, d, info = env.step_test(int(g["actions"][t]), fading=fade[2 + t])
        same = (np.array_equal(env.channel.current_BS, g["cur"][t]) and int(env._b.n_out[0]) == g["n_out"][t]
                and int(env._b.n_ho[0]) == g["n_ho"][t] and state_checksum(s) == float(g["state_chk"][t]))
        if guarded:
            assert same, t
            assert np.array_equal(info.bs_loc[:, :2], g["bs_xy"][t]), t
        elif not same:
            first_div = t                        # from here on the run is another trajectory: stop comparing
            break
        if t in sinr_rows:
            worst = max(worst, float(np.max(np.abs(env.channel.current_BS_sinr - g["cur_sinr"][sinr_rows[t]]))))
        worst_mean = max(worst_mean, abs(float(env._b.mean_sinr[0]) - g["mean_sinr"][t]))
        rew.append(r)
        ref_rew.append(g["reward"][t])
    assert worst < SINR_TOL_DB and worst_mean < SINR_TOL_DB, (worst, worst_mean)
    rep = _strict_reward_report(rew, ref_rew, "trace_fixture_reward_" + precision)
    assert rep["worst_rel_err_abs_reward_ge_0.1"] < REWARD_RTOL, rep
    assert rep["worst_abs_err"] < SINR_TOL_DB / 20, rep
    _record("trace_fixture_" + precision, {"steps_compared": len(rew), "ue_steps": 40 * len(rew), "first_divergent_step": first_div,
                                           "worst_serving_sinr_err_db": worst, "worst_mean_sinr_err_db": worst_mean,
                                           "guard_hits": env._b.guard_hits if guarded else None})
    if guarded:
        assert env._b.guard_hits > 0                  # the float64 re-evaluation ran (and is rare)
        assert env._b.guard_hits < 0.01 * 40 * T


# ---------------------------------------------------------------------------------------------------------
def test_group_mobility_matches_reference_fixture(pkg, orc, golden_dir):
    """Group mobility + channel driven by the uniforms / fading the UNMODIFIED reference drew
    (ref_group_replay.npz): the GPU generator is initialised from the reference's init draws, ticked 201 times
    (reset() is one tick) with the recorded uniforms, then reset + stepped against the recorded outputs."""
    g = _load(golden_dir, "ref_group_replay.npz")
    cfg = orc.default_cfg()
    u = g["u_ctor"]
    n, ng, G = 40, 4, 100.0
    # ue_mobility.py:434-448 init draws, in reference order
    x0, y0, th0 = u[0:n] * G, u[n:2 * n] * G, u[2 * n:3 * n]
    k = 3 * n
    gx, gy, gfl, gv = u[k:k + ng] * G, u[k + ng:k + 2 * ng] * G, u[k + 2 * ng:k + 3 * ng] * G, u[k + 3 * ng:k + 4 * ng]
    gth = u[k + 4 * ng:k + 5 * ng] * (2 * np.pi)
    k += 5 * ng
    env = pkg.BatchedMobiEnvironment(1, 4, 40, 100, "group", fading="injected", precision="fp64", warmup_ticks=-1)
    st = env.get_state()
    st["xy"][0, :, 0], st["xy"][0, :, 1], st["theta_u"][0] = x0, y0, th0
    st["group"][0] = np.stack([gx, gy, gfl, gv, np.cos(gth), np.sin(gth)])
    st["counters"][0, :5] = [0, 0, 0, 200, 100]                    # tick, epoch, step_n, aggregating, deaggregating
    env.set_state(st)
    m = orc.Mobility(cfg, [10, 10, 10, 10])                        # the oracle walks alongside to size the draws
    assert m.init(u) == k
    dummy_fade = np.zeros((1, 40, 4))
    xy = None
    for tick in range(201):
        uu = _pad(u[k:k + 52])
        xy, used = m.tick(uu)
        env.reset(fading=dummy_fade, mob_uniforms=uu[None])
        k += used
        if tick % 25 == 0 or tick == 200:
            s = env.get_state()
            assert np.max(np.abs(s["xy"][0] - xy)) < 1e-9, tick
    assert k == u.size
    assert np.array_equal(env.get_state()["ue_cell"][0], g["ue0"])
    # constructor channel pass, then reset + steps with the recorded fading / uniforms
    env.ctor_pass(fading=g["f_ctor"][None])
    obs = env.reset(fading=g["f_reset"][None], mob_uniforms=_pad(g["u_reset"])[None])
    assert np.array_equal(_np(env.serving[0]), g["reset_cur"])
    assert state_checksum(_np(obs[0]).astype(np.float64)) == float(g["reset_state_chk"])
    off = 0
    for t in range(len(g["actions"])):
        nn = int(g["u_len"][t])
        uu = _pad(g["u_steps"][off:off + nn])[None]
        off += nn
        obs, r, d, info = env.step(np.array([g["actions"][t]]), fading=g["f_steps"][t][None], mob_uniforms=uu)
        assert np.array_equal(_np(info["ue_xy"][0]), g["ue"][t]), t
        assert np.array_equal(_np(info["bs_xy"][0]), g["bs_xy"][t]), t
        assert np.array_equal(_np(info["serving"][0]), g["cur"][t]), t
        assert int(info["n_out"][0]) == g["n_out"][t] and int(info["n_ho"][0]) == g["n_ho"][t], t
        assert np.max(np.abs(_np(info["serving_sinr"][0]) - g["cur_sinr"][t])) < F64_SINR_TOL_DB, t
        assert abs(float(r[0]) - g["reward"][t]) <= 1e-9 * max(1.0, abs(g["reward"][t])), t
        assert state_checksum(_np(obs[0]).astype(np.float64)) == float(g["state_chk"][t]), t
    assert env.check() == 0


# ---------------------------------------------------------------------------------------------------------
def _oracle_envs(orc, E, seed, env_offset=0, **cfg_over):
    return [orc.OracleEnv(orc.default_cfg(**cfg_over), seed=seed, env_id=env_offset + e) for e in range(E)]


@pytest.mark.parametrize("E,steps", [(6, 260)])
def test_philox_fp64_matches_oracle(pkg, orc, E, steps):
    """Synthetic mode (Philox mobility + fading), float64 kernels vs the oracle with the same counters: every
    integer output exact, SINR 1e-9 dB, reward 1e-9; covers the 200/100/10 aggregation phase flips and reset."""
    seed = 77
    env = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", precision="fp64", seed=seed, diagnostics=True)
    oenvs = _oracle_envs(orc, E, seed)
    st = env.get_state()
    for e in range(E):
        assert np.array_equal(st["ue_cell"][e], oenvs[e].ue_xy), e           # constructor positions (201 ticks)
        assert np.array_equal(st["ho_word"][e] & 31, oenvs[e].current_BS), e  # constructor association
    obs = env.reset()
    for e in range(E):
        assert np.array_equal(_np(obs[e]).astype(np.float64), oenvs[e].reset())
    rs = np.random.RandomState(5)
    for t in range(steps):
        act = rs.randint(0, 625, size=E)
        if t == 130:                                                          # reset half of the envs mid-run
            mask = (np.arange(E) % 2).astype(np.uint8)
            obs = env.reset(env_mask=mask)
            for e in range(E):
                if mask[e]:
                    assert np.array_equal(_np(obs[e]).astype(np.float64), oenvs[e].reset())
        obs, r, d, info = env.step(act)
        for e in range(E):
            s, orw, od, oi = oenvs[e].step(int(act[e]))
            assert np.array_equal(_np(info["ue_xy"][e]), oenvs[e].ue_xy), (t, e)
            assert np.array_equal(_np(info["bs_xy"][e]), oenvs[e].bs_xy), (t, e)
            assert np.array_equal(_np(info["serving"][e]), oenvs[e].current_BS), (t, e)
            assert int(info["n_out"][e]) == oi["n_out"] and int(info["n_ho"][e]) == oi["n_ho"], (t, e)
            assert int(info["n_blocked"][e]) == oi["n_blocked"], (t, e)
            assert int(info["step_n"][e]) == oi["step_n"] and bool(d[e]) == od
            assert np.max(np.abs(_np(env.sinr_all[e]) - oenvs[e].last_sinr)) < F64_SINR_TOL_DB, (t, e)
            assert abs(float(r[e]) - orw) <= 1e-9 * max(1.0, abs(orw)), (t, e)
            assert np.array_equal(_np(obs[e]).astype(np.float64), s), (t, e)
    assert env.check() == 0


def _chan_replay(orc, cfg, sinr_steps, sinr_reset):
    """Drive the oracle's LTEChannel state machine (channel.py:113-116,138-176) with a given SINR history."""
    L = orc.lib()
    n_ue, n_bs = sinr_reset.shape
    ch = L.orc_chan_create(n_ue, n_bs)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    out = []
    try:
        s = np.ascontiguousarray(sinr_reset, dtype=np.float64)
        L.orc_chan_reset(C.byref(cfg), ch, dp(s))
        cur = np.ctypeslib.as_array(ch.contents.cur, shape=(n_ue,))
        cs = np.ctypeslib.as_array(ch.contents.cur_sinr, shape=(n_ue,))
        for s in sinr_steps:
            s = np.ascontiguousarray(s, dtype=np.float64)
            ms, no, nh = C.c_double(), C.c_int32(), C.c_int32()
            L.orc_chan_update(C.byref(cfg), ch, dp(s), C.byref(ms), C.byref(no), C.byref(nh))
            out.append((cur.copy(), cs.copy(), ms.value, no.value, nh.value))
    finally:
        L.orc_chan_destroy(ch)
    return out


@pytest.mark.parametrize("E,steps,nBS,nUE", [(8, 200, 4, 40), (3, 40, 32, 256), (3, 60, 12, 100), (3, 60, 7, 64)])
def test_philox_fp32_matches_oracle(pkg, orc, E, steps, nBS, nUE):
    """fp32 kernels vs the oracle in synthetic mode, for the thread-per-UE mapping (nBS <= 4) and the
    4-BSs-per-lane mapping with shuffle reductions (nBS > 4, incl. BS counts that do not fill the lanes).
    UE / BS cells are exact (mobility is float64 and does not depend on SINR decisions).  The SINR matrix of every
    pass is within 1e-3 dB of the oracle's float64 SINR for the same cells and the same Philox fading, the fading
    itself within 1e-4 dB, and the handover / outage state machine run on the GPU's own SINR values reproduces the
    GPU's serving cells and counts exactly."""
    seed = 4242
    nG = 4
    gs = [nUE // nG] * nG
    kw, okw = {}, {}
    if nBS != 4:
        side = int(np.ceil(np.sqrt(nBS)))
        layout = [[max(2, (b // side + 1) * 100 // (side + 1)), max(2, (b % side + 1) * 100 // (side + 1))] for b in range(nBS)]
        kw, okw = dict(init_bs_xy=layout, group_sizes=gs), dict(init_bs_xy=layout, group_sizes=gs)
    env = pkg.BatchedMobiEnvironment(E, nBS, nUE, 100, "group", precision="fp32", seed=seed, diagnostics=True, **kw)
    cfg = orc.default_cfg(nBS, nUE, 100, nG)
    oenvs = [orc.OracleEnv(cfg, seed=seed, env_id=e, **okw) for e in range(E)]
    env.reset()
    for o in oenvs:
        o.reset()
    sinr_reset = _np(env.sinr_all).astype(np.float64)
    hist, gpu = [], []
    rs = np.random.RandomState(9)
    worst_sinr = worst_fade = worst_rew = 0.0
    for t in range(steps):
        digits = rs.randint(0, 5, size=(E, nBS)).astype(np.uint8)
        obs, r, d, info = env.step(digits)
        sa = _np(env.sinr_all).astype(np.float64)
        hist.append(sa)
        gpu.append((_np(info["serving"]).copy(), _np(info["serving_sinr"]).astype(np.float64), _np(info["mean_sinr"]).copy(),
                    _np(info["n_out"]).copy(), _np(info["n_ho"]).copy(), _np(r).copy()))
        fu = _np(env.fading_used).astype(np.float64)
        for e in range(E):
            oenvs[e].step(digits[e].astype(np.int32))
            assert np.array_equal(_np(info["ue_xy"][e]), oenvs[e].ue_xy), (t, e)
            assert np.array_equal(_np(info["bs_xy"][e]), oenvs[e].bs_xy), (t, e)
            worst_sinr = max(worst_sinr, float(np.max(np.abs(sa[e] - oenvs[e].last_sinr))))
            # oracle epoch of this pass: ctor 0, reset 1, step t -> 2 + t
            of = np.empty((nUE, nBS))
            orc.lib().orc_philox_fading(C.byref(cfg), seed, e, 2 + t, of.ctypes.data_as(C.POINTER(C.c_double)))
            worst_fade = max(worst_fade, float(np.max(np.abs(fu[e] - of))))
    assert worst_fade < 2e-4, worst_fade          # MUFU Box-Muller (lg2, sqrt, sin, cos) on N(0, 2 dB)
    assert worst_sinr < SINR_TOL_DB, worst_sinr
    for e in range(E):
        rep = _chan_replay(orc, cfg, [h[e] for h in hist], sinr_reset[e])
        for t, (cur, cs, ms, no, nh) in enumerate(rep):
            srv, ssinr, mean, n_out, n_ho, rew = gpu[t]
            assert np.array_equal(srv[e], cur), (t, e)
            assert np.array_equal(ssinr[e], cs), (t, e)
            assert n_out[e] == no and n_ho[e] == nh, (t, e)
            assert abs(mean[e] - ms) < 1e-9
            want = max(ms / 20 - no / float(nUE), -1.0)
            worst_rew = max(worst_rew, abs(rew[e] - want) / max(abs(want), 1e-12))
    assert worst_rew < 1e-9, worst_rew


# ---------------------------------------------------------------------------------------------------------
def test_dense_channel_matches_reference_fixture(pkg, golden_dir):
    """Config-4 sizes (32 BS x 2048 UE): the reference's LTEChannel / BS_move driven directly (fixture), UE
    cells from a per-step trace, per-BS digit actions, injected fading; float64 kernels."""
    g = _load(golden_dir, "ref_dense_channel.npz")
    n_steps, n_ue = g["cur"].shape
    n_bs = g["init_bs"].shape[0]
    fade = np.random.RandomState(int(g["seed"]) + 1).normal(0, 2, size=(n_steps + 1, n_ue, n_bs))
    # trace row 0 feeds the ctor pass; the fixture moved the UEs to ue[t+1] before update t, while step_test
    # reads trace[step_n] (mobile_env.py:203) and step_n is 0 at the first step: start the replay at step_n = 1
    env = pkg.BatchedMobiEnvironment(1, n_bs, n_ue, 100, "read_trace", trace=g["ue"], fading="injected",
                                     precision="fp64", init_bs_xy=g["init_bs"][:, :2])
    env.ctor_pass(fading=fade[0][None])
    assert np.array_equal(_np(env.serving[0]), g["ctor_cur"])
    assert np.max(np.abs(_np(env.serving_sinr[0]) - g["ctor_sinr"])) < F64_SINR_TOL_DB
    st = env.get_state()
    st["counters"][0, 2] = 1
    env.set_state(st)
    for t in range(n_steps):
        obs, r, d, info = env.step(g["digits"][t].astype(np.uint8)[None], fading=fade[t + 1][None])
        assert np.array_equal(_np(info["bs_xy"][0]), g["bs_xy"][t][:, :2]), t
        assert np.array_equal(_np(info["serving"][0]), g["cur"][t]), t
        assert int(info["n_out"][0]) == g["n_out"][t], t
        assert np.max(np.abs(_np(info["serving_sinr"][0]) - g["cur_sinr"][t])) < F64_SINR_TOL_DB, t
        assert abs(float(info["mean_sinr"][0]) - g["mean_sinr"][t]) < F64_SINR_TOL_DB, t
        o = _np(obs[0]).astype(np.float64)
        assert o[0].sum() == n_bs and o[1:].sum() == n_ue
        amap = o[1:]
        assert float(np.sum(amap * (1 + (np.arange(amap.size).reshape(amap.shape) % 8191)))) == float(g["amap_chk"][t]), t
    assert env.check() == 0


# ---------------------------------------------------------------------------------------------------------
_SWEEP_ORACLE = {}


def _sweep_oracle(orc, E, T, seed, golden_dir):
    """the C oracle's side of the config[4] sweep, computed once per session and shared by the precisions.  The trace is
    ue_trace_10k regenerated from the REFERENCE's own generator (tests/golden/ue_trace_10k.npz, oracle/make_golden.py
    trace10k: seed + SHA-256 inside; its first 2100 rows are the trace of the reference fixture)."""
    from concurrent.futures import ThreadPoolExecutor
    key = (E, T, seed)
    if key not in _SWEEP_ORACLE:
        cfg = orc.default_cfg()
        trace = load_trace_10k(golden_dir).astype(np.int32)
        assert trace.shape[0] >= T + 1
        acts = np.stack([np.random.RandomState(1000 + e).randint(0, 625, size=T) for e in range(E)], axis=1)   # [T, E]
        with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
            outs = list(ex.map(lambda e: orc.replay_run(cfg, trace, seed, e, acts[:, e]), range(E)))
        _SWEEP_ORACLE[key] = (trace, acts, [np.stack([o[k] for o in outs], axis=1) for k in range(4)])
    return _SWEEP_ORACLE[key]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("precision", ["fp32_guarded", "fp64", "fp32"])
def test_trace_replay_equivalence_sweep_config5(pkg, orc, golden_dir, precision):
    """BASELINE config[4] ("trace-replay equivalence sweep"): 1024 independent read_trace envs x 10 000 step_test
    calls over ue_trace_10k regenerated from the reference's generator, env e driven by its own fixed action stream RandomState(1000+e) and its
    own Philox fading stream, against the C oracle (float64; pinned to the reference by tests/golden).  4.1e8 UE-steps;
    MAXSTEP/done is ignored like main_test.py:70-103 ignores it.
      fp64, fp32_guarded (the mode bench.py measures): new-outage counts, handover counts and a hash of every UE's serving
        BS bit-exact at every step of every env; reward within 1e-9 (fp64) / the north star's 1e-5 relative (guarded;
        strict violations near reward = 0 are counted and reported, see _strict_reward_report);
      fp32 (no guard): how many envs ever leave the oracle's trajectory, and after how many UE-steps, is REPORTED --
        the measured decision-flip rate of a plain fp32 pass (SURVEY H2 predicted about 4e-6 per UE-step)."""
    E, T, seed = 1024, 10000, 555
    trace, acts, (o_out, o_ho, o_rew, o_hsh) = _sweep_oracle(orc, E, T, seed, golden_dir)
    env = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "read_trace", trace=trace, fading="philox", precision=precision,
                                     obs="none", seed=seed)
    env.reset()
    dev = env.device
    acts_d = torch.from_numpy(acts).to(dev)
    n_out = torch.empty((T, E), dtype=torch.int32, device=dev)
    n_ho = torch.empty((T, E), dtype=torch.int32, device=dev)
    rew = torch.empty((T, E), dtype=torch.float64, device=dev)
    hsh = torch.empty((T, E), dtype=torch.int64, device=dev)
    w = torch.arange(1, 41, device=dev, dtype=torch.int64)
    for t in range(T):
        _, r, _, info = env.step(acts_d[t])
        n_out[t].copy_(info["n_out"])
        n_ho[t].copy_(info["n_ho"])
        rew[t].copy_(r)
        hsh[t] = ((info["serving"].long() + 1) * w).sum(dim=1)
    assert env.check() == 0
    n_out, n_ho, rew, hsh = _np(n_out), _np(n_ho), _np(rew), _np(hsh)
    assert int(o_ho.sum()) > 100000 and int(o_out.sum()) > 100000          # the sweep exercises the state machine
    if precision == "fp32":
        # a flipped decision changes the env's trajectory for good: count envs by their first divergent step
        bad = (hsh != o_hsh) | (n_out != o_out) | (n_ho != o_ho)
        first = np.where(bad.any(axis=0), bad.argmax(axis=0), T)           # [E] first divergent step (T = never)
        clean_ue_steps = int(first.sum()) * 40
        n_div = int((first < T).sum())
        _record("sweep_config5_fp32_unguarded", {
            "envs": E, "steps": T, "envs_that_left_the_oracle_trajectory": n_div,
            "ue_steps_before_divergence": clean_ue_steps,
            "decision_flips_per_ue_step": n_div / max(clean_ue_steps, 1),
            "median_first_divergent_step": float(np.median(first[first < T])) if n_div else None})
        assert n_div / max(clean_ue_steps, 1) < 1e-4       # sanity bound only; the guarded mode is the exact one
        return
    assert np.array_equal(hsh, o_hsh), "serving BS mismatch at %d (step, env) pairs" % int((hsh != o_hsh).sum())
    assert np.array_equal(n_out, o_out)
    assert np.array_equal(n_ho, o_ho)
    if precision == "fp64":
        # float64 against float64 (libm vs CUDA log10 / pow differ by ~1e-13 dB): 1e-9 relative, and absolute where the
        # reward passes through zero (10 M samples: some land within 1e-7 of it)
        err = np.abs(rew - o_rew)
        assert float((err / np.maximum(np.abs(o_rew), 1.0)).max()) < 1e-9
        rel = err / np.maximum(np.abs(o_rew), 1e-6)
        assert rel.max() < 1e-8, rel.max()
    else:
        rep = _strict_reward_report(rew, o_rew, "sweep_config5_reward_fp32_guarded")
        assert rep["worst_rel_err_abs_reward_ge_0.1"] < REWARD_RTOL, rep
        assert rep["worst_abs_err"] < SINR_TOL_DB / 20, rep
        hits = env.guard_hits
        _record("sweep_config5_fp32_guarded", {"envs": E, "steps": T, "ue_steps": E * T * 40, "decision_mismatches": 0,
                                               "guard_hits": hits, "guard_hits_per_ue_step": hits / (E * T * 40.0),
                                               "guard_db": float(env.cfg.guard_db)})
        assert 0 < hits < 0.01 * E * T * 40


@pytest.mark.timeout(900)
@pytest.mark.parametrize("precision", ["fp32_guarded", "fp64"])
def test_group_mode_full_size_config1_matches_oracle(pkg, orc, precision):
    """BASELINE config[1] at full size -- 4096 envs, 4 x 40, grid 100, group mobility, Philox fading -- against the oracle
    for every env: 260 steps across the 200/100/10 aggregation phase flips with MAXSTEP lowered to 120 so that every env is
    reset twice (main.py:188-190).  UE-cell hash, serving-BS hash, new-outage and handover counts bit-exact at every step of
    every env (4.3e7 UE-steps) for the float64 kernels AND the guarded fp32 kernels (the benchmarked mode); reward within
    1e-9 relative (fp64) / the north star's 1e-5 (guarded)."""
    from concurrent.futures import ThreadPoolExecutor
    E, T, seed = 4096, 260, 909
    cfg = orc.default_cfg(max_step=120)
    acts = np.random.RandomState(12).randint(0, 625, size=(T, E))
    env = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", precision=precision, obs="none", seed=seed, max_step=120)
    env.reset()
    dev = env.device
    acts_d = torch.from_numpy(acts).to(dev)
    n_out = torch.empty((T, E), dtype=torch.int32, device=dev)
    n_ho = torch.empty((T, E), dtype=torch.int32, device=dev)
    rew = torch.empty((T, E), dtype=torch.float64, device=dev)
    hsh = torch.empty((T, E), dtype=torch.int64, device=dev)
    hc = torch.empty((T, E), dtype=torch.int64, device=dev)
    w = torch.arange(1, 41, device=dev, dtype=torch.int64)
    for t in range(T):
        _, r, done, info = env.step(acts_d[t])
        n_out[t].copy_(info["n_out"])
        n_ho[t].copy_(info["n_ho"])
        rew[t].copy_(r)
        hsh[t] = ((info["serving"].long() + 1) * w).sum(dim=1)
        ue = info["ue_xy"].long()
        hc[t] = ((ue[..., 0] * 100 + ue[..., 1] + 1) * w).sum(dim=1)
        env.reset(env_mask=env.done_u8)
    assert env.check() == 0
    key = ("group", E, T, seed)
    if key not in _SWEEP_ORACLE:
        with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
            _SWEEP_ORACLE[key] = list(ex.map(lambda e: orc.group_run(cfg, seed, e, acts[:, e]), range(E)))
    outs = _SWEEP_ORACLE[key]
    for k, got in enumerate((n_out, n_ho, rew, hsh, hc)):
        want = np.stack([o[k] for o in outs], axis=1)
        if k == 2 and precision == "fp64":
            rel = np.abs(_np(got) - want) / np.maximum(np.abs(want), 1e-9)
            assert rel.max() < 1e-9, rel.max()
        elif k == 2:
            rep = _strict_reward_report(_np(got), want, "config1_full_size_reward_fp32_guarded")
            assert rep["worst_rel_err_abs_reward_ge_0.1"] < REWARD_RTOL, rep
        else:
            assert np.array_equal(_np(got), want), ("n_out", "n_ho", "reward", "serving", "cells")[k]


@pytest.mark.parametrize("E,steps,nBS,nUE", [(8, 300, 4, 40), (3, 60, 32, 256), (2, 30, 32, 2048), (3, 80, 12, 100),
                                             (3, 80, 7, 64), (3, 80, 3, 33)])
def test_philox_fp32_guarded_decisions_match_oracle(pkg, orc, E, steps, nBS, nUE):
    """precision="fp32_guarded" against the float64 oracle in synthetic mode, both BS mappings (thread-per-UE for nBS <= 4,
    4 BSs per lane beyond, incl. the guard list's warp-per-UE re-evaluation): UE / BS cells, serving BS of every UE,
    new-outage and handover counts and the observation bit-exact at every step; serving SINR within 1e-3 dB.  A widened
    guard (guard_db = 0.5: about one UE in ten re-evaluated, which also overflows the 64-entry guard list at 2048 UEs)
    must give the same decisions."""
    seed = 9191
    nG = 4 if nUE % 4 == 0 else 3
    gs = [nUE // nG] * nG
    kw = {}
    if nBS != 4:
        side = int(np.ceil(np.sqrt(nBS)))
        layout = [[max(2, (b // side + 1) * 100 // (side + 1)), max(2, (b % side + 1) * 100 // (side + 1))] for b in range(nBS)]
        kw = dict(init_bs_xy=layout, group_sizes=gs)
    elif nG != 4:
        kw = dict(group_sizes=gs)
    env = pkg.BatchedMobiEnvironment(E, nBS, nUE, 100, "group", precision="fp32_guarded", seed=seed, **kw)
    wide = pkg.BatchedMobiEnvironment(E, nBS, nUE, 100, "group", precision="fp32_guarded", seed=seed, guard_db=0.5, **kw)
    cfg = orc.default_cfg(nBS, nUE, 100, nG)
    oenvs = [orc.OracleEnv(cfg, seed=seed, env_id=e, **kw) for e in range(E)]
    obs = env.reset()
    wobs = wide.reset()
    for e, o in enumerate(oenvs):
        want = o.reset()
        assert np.array_equal(_np(obs[e]).astype(np.float64), want), e
        assert np.array_equal(_np(wobs[e]).astype(np.float64), want), e
    rs = np.random.RandomState(3)
    worst = 0.0
    for t in range(steps):
        digits = rs.randint(0, 5, size=(E, nBS)).astype(np.uint8)
        obs, r, d, info = env.step(digits)
        got = {k: _np(info[k]).copy() for k in ("ue_xy", "bs_xy", "serving", "n_out", "n_ho", "serving_sinr")}
        obs_np = _np(obs).astype(np.float64)
        wobs, _, _, winfo = wide.step(digits)
        for e in range(E):
            s, orw, od, oi = oenvs[e].step(digits[e].astype(np.int32))
            assert np.array_equal(got["ue_xy"][e], oenvs[e].ue_xy), (t, e)
            assert np.array_equal(got["bs_xy"][e], oenvs[e].bs_xy), (t, e)
            assert np.array_equal(got["serving"][e], oenvs[e].current_BS), (t, e)
            assert int(got["n_out"][e]) == oi["n_out"] and int(got["n_ho"][e]) == oi["n_ho"], (t, e)
            assert np.array_equal(obs_np[e], s), (t, e)
            worst = max(worst, float(np.max(np.abs(got["serving_sinr"][e] - oenvs[e].current_BS_sinr))))
            assert abs(float(r[e]) - orw) <= REWARD_RTOL * abs(orw) or abs(orw) < 0.1, (t, e)
        for k in ("serving", "n_out", "n_ho"):
            assert np.array_equal(_np(winfo[k]), got[k]), (k, t)
        assert torch.equal(wobs, obs), t
    assert worst < SINR_TOL_DB, worst
    assert env.check() == 0 and wide.check() == 0
    n = E * nUE * (steps + 1)
    assert wide.guard_hits > 0.02 * n, (wide.guard_hits, n)
    assert env.guard_hits < 0.02 * n, (env.guard_hits, n)
    _record("guarded_vs_oracle_%dx%d" % (nBS, nUE), {"ue_passes": n, "guard_hits": env.guard_hits,
                                                     "guard_hits_wide_0.5dB": wide.guard_hits,
                                                     "worst_serving_sinr_err_db": worst})
