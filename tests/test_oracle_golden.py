"""Pins the C parity oracle (oracle/mobi_oracle.c) against fixtures recorded from the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle.make_golden import state_checksum


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_philox_known_answers(orc):
    # Random123 kat_vectors, philox4x32-10
    assert orc.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox4x32([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("n", [0, 1, 3, 7, 8, 9, 31, 40, 128, 129, 2047, 2048])
def test_np_sum_order(orc, n):
    rs = np.random.RandomState(n)
    a = rs.rand(n) * 10.0 ** rs.randint(-12, 12, size=n)
    assert orc.np_sum(a) == float(np.sum(a))


def test_action_digits_msb_first(orc):
    assert orc.action_digits(624).tolist() == [4, 4, 4, 4]
    assert orc.action_digits(0).tolist() == [0, 0, 0, 0]
    assert orc.action_digits(125 * 1 + 25 * 2 + 5 * 3 + 4).tolist() == [1, 2, 3, 4]
    assert orc.action_digits(7).tolist() == [0, 0, 1, 2]
    with pytest.raises(ValueError):
        orc.action_digits(625)


@pytest.mark.parametrize("run", ["uniform", "biased", "walls"])
def test_bs_move_golden(orc, golden_dir, run):
    g = _load(golden_dir, "ref_bs_move.npz")
    cfg = orc.default_cfg()
    loc = g["init"].astype(np.int64)
    acts, want, wdig = g[run + "_actions"], g[run + "_loc"], g[run + "_digits"]
    blocked_total = 0
    for t in range(len(acts)):
        d = orc.action_digits(int(acts[t]))
        assert d.tolist() == wdig[t].tolist()
        loc, blocked = orc.bs_move(cfg, loc, d)
        blocked_total += blocked
        assert np.array_equal(loc, want[t]), (run, t)
    if run != "walls":
        assert blocked_total > 0  # the lock quirk is exercised


def test_mobility_golden(orc, golden_dir):
    g = _load(golden_dir, "ref_mobility.npz")
    n_ticks = int(g["n_ticks"])
    flat = np.random.RandomState(int(g["seed"])).rand(int(g["n_uniforms"]) + 200)
    cfg = orc.default_cfg()
    m = orc.Mobility(cfg, [10, 10, 10, 10])
    k = m.init(flat)
    assert k == 3 * 40 + 5 * 4
    keep = {int(t): i for i, t in enumerate(g["keep"])}
    w1, w2 = np.arange(40) + 1, (np.arange(40) + 41) * 101
    max_err = 0.0
    for t in range(n_ticks):
        xy, used = m.tick(flat[k:k + 40 + 12])
        k += used
        c = xy.astype(int)
        assert int(np.sum(c[:, 0] * w1 + c[:, 1] * w2)) == int(g["cell_chk"][t]), t
        if t in keep:
            max_err = max(max_err, float(np.max(np.abs(xy - g["pos"][keep[t]]))))
    assert k == int(g["n_uniforms"])     # consumed exactly as many uniforms as the reference
    assert max_err < 1e-9, max_err       # libm vs numpy SIMD sin/cos may differ in the last ulp


def test_group_replay_golden(orc, golden_dir):
    g = _load(golden_dir, "ref_group_replay.npz")
    cfg = orc.default_cfg()
    # rebuild the constructor's mobility state from the recorded uniforms (mobile_env.py:76-79,93-97)
    m = orc.Mobility(cfg, [10, 10, 10, 10])
    u = g["u_ctor"]
    k = m.init(u)
    xy = None
    for _ in range(201):
        xy, used = m.tick(u[k:k + 52] if k + 52 <= u.size else np.concatenate([u[k:], np.zeros(52)]))
        k += used
    assert k == u.size
    assert np.array_equal(xy.astype(int), g["ue0"])
    env = orc.OracleEnv(cfg, mobility=orc.MOB_GROUP, fading=orc.FADE_INJECTED, mob_state=m.get_state(), ue_float=xy,
                        ctor_fading=g["f_ctor"])
    s0 = env.reset(fading=g["f_reset"], mob_uniforms=np.concatenate([g["u_reset"], np.zeros(12)]))
    assert np.array_equal(env.current_BS, g["reset_cur"])
    assert state_checksum(s0) == float(g["reset_state_chk"])
    off = 0
    for t in range(len(g["actions"])):
        n = int(g["u_len"][t])
        uu = np.concatenate([g["u_steps"][off:off + n], np.zeros(12)])
        off += n
        before = env.current_BS
        s, r, d, info = env.step(int(g["actions"][t]), fading=g["f_steps"][t], mob_uniforms=uu)
        assert np.array_equal(env.ue_xy, g["ue"][t]), t
        assert np.array_equal(env.bs_xy, g["bs_xy"][t]), t
        assert np.array_equal(env.current_BS, g["cur"][t]), t
        assert info["n_out"] == g["n_out"][t] and info["n_ho"] == g["n_ho"][t], t
        assert info["n_ho"] == int(np.sum(before != env.current_BS))
        assert np.max(np.abs(env.current_BS_sinr - g["cur_sinr"][t])) < 1e-11, t
        assert abs(r - g["reward"][t]) <= 1e-12 * max(1.0, abs(g["reward"][t])), t
        assert state_checksum(s) == float(g["state_chk"][t]), t
    assert env.n_clamped == 0


def test_trace_replay_golden(orc, golden_dir):
    """BASELINE config 1 with a fixed action sequence: 2001 step_test calls on a regenerated trace."""
    g = _load(golden_dir, "ref_trace_replay.npz")
    T = len(g["actions"])
    fade = np.random.RandomState(int(g["fade_seed"])).normal(0, 2, size=(T + 2, 40, 4))
    assert np.array_equal(fade[:3, :2, :], g["fade_probe"])  # numpy's legacy stream is what the fixture used
    cfg = orc.default_cfg()
    env = orc.OracleEnv(cfg, mobility=orc.MOB_TRACE, fading=orc.FADE_INJECTED, trace=g["trace"], ctor_fading=fade[0])
    assert np.array_equal(env.current_BS, g["ctor_cur"])
    assert float(env.state.sum()) == 0.0           # zeros until the first reset (mobile_env.py:107)
    s0 = env.reset(fading=fade[1])
    assert np.array_equal(env.current_BS, g["reset_cur"])
    assert np.max(np.abs(env.current_BS_sinr - g["reset_sinr"])) < 1e-11
    assert state_checksum(s0) == float(g["reset_state_chk"])
    sinr_rows = {int(t): i for i, t in enumerate(g["sinr_idx"])}
    max_sinr_err = max_rew_rel = 0.0
    for t in range(T):
        s, r, d, info = env.step(int(g["actions"][t]), fading=fade[2 + t])
        assert np.array_equal(env.current_BS, g["cur"][t]), t
        assert info["n_out"] == g["n_out"][t], t
        assert info["n_ho"] == g["n_ho"][t], t
        assert np.array_equal(env.bs_xy, g["bs_xy"][t]), t
        assert info["digits"].tolist() == g["digits"][t].tolist()
        assert d == bool(g["done"][t])
        assert state_checksum(s) == float(g["state_chk"][t]), t
        assert abs(info["mean_sinr"] - g["mean_sinr"][t]) < 1e-11
        if abs(g["reward"][t]) > 0:
            max_rew_rel = max(max_rew_rel, abs(r - g["reward"][t]) / abs(g["reward"][t]))
        if t in sinr_rows:
            max_sinr_err = max(max_sinr_err, float(np.max(np.abs(env.current_BS_sinr - g["cur_sinr"][sinr_rows[t]]))))
    assert g["done"][1999] == 1 and g["done"][1998] == 0   # MAXSTEP = 2000 (mobile_env.py:18,186)
    assert max_sinr_err < 1e-11, max_sinr_err
    assert max_rew_rel < 1e-10, max_rew_rel
    with pytest.raises(IndexError):                        # past the end of the trace (mobile_env.py:203)
        env.step_n = 2100
        env.step(0, fading=fade[0])


def test_dense_channel_golden(orc, golden_dir):
    """Config-4 sizes (32 BS x 2048 UE) against LTEChannel/BS_move driven directly."""
    g = _load(golden_dir, "ref_dense_channel.npz")
    n_steps, n_ue = g["cur"].shape
    n_bs = g["init_bs"].shape[0]
    fade = np.random.RandomState(int(g["seed"]) + 1).normal(0, 2, size=(n_steps + 1, n_ue, n_bs))
    cfg = orc.default_cfg(n_bs=n_bs, n_ue=n_ue, grid_n=100, n_groups=32)
    import ctypes as C
    L = orc.lib()
    ch = L.orc_chan_create(n_ue, n_bs)
    try:
        s = orc.sinr_all(cfg, g["ue"][0], g["init_bs"], fade[0])
        L.orc_chan_reset(C.byref(cfg), ch, s.ctypes.data_as(C.POINTER(C.c_double)))
        cur = np.ctypeslib.as_array(ch.contents.cur, shape=(n_ue,))
        cur_sinr = np.ctypeslib.as_array(ch.contents.cur_sinr, shape=(n_ue,))
        assert np.array_equal(cur, g["ctor_cur"])
        assert np.max(np.abs(cur_sinr - g["ctor_sinr"])) < 1e-11
        loc = g["init_bs"].astype(np.int64)
        for t in range(n_steps):
            loc, _ = orc.bs_move(cfg, loc, g["digits"][t].astype(np.int32))
            assert np.array_equal(loc, g["bs_xy"][t])
            s = orc.sinr_all(cfg, g["ue"][t + 1], loc, fade[t + 1])
            ms, no, nh = C.c_double(), C.c_int32(), C.c_int32()
            L.orc_chan_update(C.byref(cfg), ch, s.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ms), C.byref(no),
                              C.byref(nh))
            assert np.array_equal(cur, g["cur"][t]), t
            assert no.value == g["n_out"][t]
            assert np.max(np.abs(cur_sinr - g["cur_sinr"][t])) < 1e-11
            assert abs(ms.value - g["mean_sinr"][t]) < 1e-11
    finally:
        L.orc_chan_destroy(ch)


def test_replay_run_matches_env_stepping(orc):
    """The config-5 sweep helper (orc_replay_run: one C call per env) == stepping the oracle env from Python."""
    cfg = orc.default_cfg()
    tr = orc.make_trace(cfg, 7, 0, 120)
    assert tr.shape == (120, 40, 2) and tr.min() >= 0 and tr.max() <= 99
    acts = np.random.RandomState(0).randint(0, 625, size=119)
    n_out, n_ho, rew, hsh = orc.replay_run(cfg, tr, 7, 3, acts)
    o = orc.OracleEnv(cfg, mobility=orc.MOB_TRACE, seed=7, env_id=3, trace=tr, warmup_ticks=0)
    o.reset()
    for t in range(119):
        s, r, d, info = o.step(int(acts[t]), want_state=False)
        assert (info["n_out"], info["n_ho"], r) == (n_out[t], n_ho[t], rew[t]), t
        assert int(((np.arange(40) + 1) * (o.current_BS + 1)).sum()) == hsh[t], t
    with pytest.raises(IndexError):
        orc.replay_run(cfg, tr, 7, 3, np.zeros(121, dtype=np.int64))


def test_sinr_in_area_golden(orc, golden_dir):
    """GetSinrInArea (channel.py:411-433) of the unmodified reference, three BS layouts (one with equidistant BSs),
    fading draws replayed in the reference's call order: the oracle's coverage map is bit-identical."""
    g = _load(golden_dir, "ref_sinr_area.npz")
    cfg = orc.default_cfg()
    for i in range(g["bs"].shape[0]):
        out = orc.sinr_in_area(cfg, g["bs"][i], fading=g["fading"][i])
        assert np.array_equal(out[0], np.zeros(100)) and np.array_equal(out[:, 0], np.zeros(100))
        assert np.max(np.abs(out - g["sinr"][i])) <= 1e-12, i
    # per-(cell, BS) draws are the same numbers in a different layout
    by_bs = orc.philox_area_fading(cfg, 5, 0, 0)
    assert by_bs.shape == (99 * 99, 4) and abs(by_bs.std() - 2.0) < 0.02 and abs(by_bs.mean()) < 0.02


def test_free_running_statistics_match_reference(orc, golden_dir):
    """Distributional pin: the Philox-driven oracle cannot follow the reference's Mersenne-Twister
    stream draw for draw, so its free-running statistics -- reward, new outages, handovers, serving SINR, UE movement,
    outage fraction per step, group mode, uniform random actions -- are compared with 8 free-running runs of the
    unmodified reference (tests/golden/ref_free_running_stats.npz).  Tolerance: 4 standard errors of the difference of
    the two means (between-env spread of both samples)."""
    g = _load(golden_dir, "ref_free_running_stats.npz")
    ref, T = g["stats"], int(g["n_steps"])
    n_envs = 24
    cfg = orc.default_cfg()
    got = np.zeros((n_envs, 6))
    for e in range(n_envs):
        o = orc.OracleEnv(cfg, seed=4321, env_id=e)
        o.reset()
        rs = np.random.RandomState(50 + e)
        prev = o.ue_xy.copy()
        acc = np.zeros(6)
        for t in range(T):
            s, r, d, info = o.step(int(rs.randint(625)), want_state=False)
            cell = o.ue_xy
            sinr = o.current_BS_sinr
            acc += (r, info["n_out"], info["n_ho"], sinr.mean(), np.abs(cell - prev).mean(), (sinr <= 0).mean())
            prev = cell.copy()
        got[e] = acc / T
    for k, name in enumerate(g["columns"]):
        se = np.sqrt(ref[:, k].var(ddof=1) / len(ref) + got[:, k].var(ddof=1) / n_envs)
        assert abs(ref[:, k].mean() - got[:, k].mean()) < 4 * se, (str(name), ref[:, k].mean(), got[:, k].mean(), se)


def test_ue_trace_10k_fixture_is_the_reference_generator_output(golden_dir):
    """tests/golden/ue_trace_10k.npz (oracle/make_golden.py trace10k): 10 001 rows of the reference's group-reference
    generator under the recorded seed; decodes to its SHA-256, stays on the grid, moves at most 3 cells per step, and its
    first 2100 rows are the trace the reference fixture (ref_trace_replay.npz) was recorded on."""
    from oracle.make_golden import load_trace_10k
    tr = load_trace_10k(golden_dir)
    assert tr.shape == (10001, 40, 2) and tr.min() >= 0 and tr.max() <= 99
    assert np.abs(np.diff(tr, axis=0)).max() <= 3
    g = np.load(os.path.join(golden_dir, "ref_trace_replay.npz"))
    assert np.array_equal(tr[:g["trace"].shape[0]], g["trace"].astype(np.int64))
