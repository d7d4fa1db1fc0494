"""Property tests (hypothesis, CPU) of the oracle's building blocks -- the unit level of SURVEY.md section 4: digit decode,
BS move incl. the lock quirk, path loss at d = 0, exclude-self interference, the time-to-trigger FIFO warm-up, new-outage
semantics, the reward clamp and the heat-map orientation.  The same properties hold for the CUDA path because the GPU
parity tests tie it to the oracle bit for bit."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import example, given, settings, strategies as st


@pytest.fixture(scope="module")
def orc():
    from oracle import mobi_oracle
    mobi_oracle.lib()
    return mobi_oracle


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 27), base=st.integers(2, 9), data=st.data())
def test_action_digits_round_trip(orc, n, base, data):
    """Decimal_to_Base_N (ue_mobility.py:310-336): MSB first, digit 0 <-> BS 0; values needing more digits are an error."""
    hi = min(base ** n, 2 ** 62)
    a = data.draw(st.integers(0, hi - 1))
    d = orc.action_digits(a, base, n)
    assert all(0 <= x < base for x in d)
    assert sum(int(x) * base ** (n - 1 - i) for i, x in enumerate(d)) == a
    if base ** n < 2 ** 62:
        with pytest.raises(ValueError):
            orc.action_digits(base ** n, base, n)


@settings(max_examples=200, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n_bs=st.integers(1, 8), steps=st.integers(1, 30))
def test_bs_move_invariants(orc, seed, n_bs, steps):
    """BS_move (ue_mobility.py:191-271): a BS moves by exactly 0 or 2 cells along one axis, never leaves [2, G-1] once
    inside, `blocked` counts the BSs that had a neighbour within the lock radius BEFORE moving, and a BS that is within
    the radius of another one stays frozen (the quirk: the test uses the pre-move position)."""
    rs = np.random.RandomState(seed)
    G = 30
    cfg = orc.default_cfg(n_bs, 8, G, 4)
    loc = rs.randint(2, G, size=(n_bs, 2)).astype(np.int64)
    for _ in range(steps):
        digits = rs.randint(0, 5, size=n_bs).astype(np.int32)
        new, blocked = orc.bs_move(cfg, loc, digits)
        delta = new - loc
        assert np.all(np.abs(delta).sum(axis=1) <= 2) and np.all((np.abs(delta).sum(axis=1) == 0) | (np.abs(delta).sum(axis=1) == 2))
        assert new.min() >= 2 and new.max() <= G - 1
        # replay the sequential rule literally
        cur = loc.copy()
        nb = 0
        for i in range(n_bs):
            d2 = ((cur - cur[i]) ** 2).sum(axis=1)
            d2[i] = 10 ** 9
            if (d2 <= 16).any():
                nb += 1
                continue
            x, y = cur[i]
            if digits[i] == 0 and x + 2 < G: x += 2
            elif digits[i] == 1 and x - 2 > 1: x -= 2
            elif digits[i] == 2 and y + 2 < G: y += 2
            elif digits[i] == 3 and y - 2 > 1: y -= 2
            cur[i] = (x, y)
        assert np.array_equal(new, cur) and blocked == nb
        loc = new


@settings(max_examples=100, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n_bs=st.integers(2, 12), n_ue=st.integers(1, 20))
@example(seed=15315943, n_bs=5, n_ue=2)      # a second BS lands on BS 0's cell: fading, not path loss, decides between them
def test_sinr_matches_direct_formula_and_handles_d0(orc, seed, n_bs, n_ue):
    """GetChannelGainAll / GetDLSinrAllDb (channel.py:249-269): loss = 38 + 30 log10(5 d) for d > 0 and 0 at d = 0
    (UE on a BS cell), interference = explicit sum over the OTHER BSs -- compared with a direct numpy evaluation."""
    rs = np.random.RandomState(seed)
    G = 40
    cfg = orc.default_cfg(n_bs, n_ue, G, 1)
    bs = rs.randint(2, G, size=(n_bs, 2)).astype(np.int64)
    ue = rs.randint(0, G, size=(n_ue, 2)).astype(np.int64)
    ue[0] = bs[0]                                                         # d = 0
    fade = rs.normal(0, 2, size=(n_ue, n_bs))
    got = orc.sinr_all(cfg, ue, bs, fade)
    d = 5.0 * np.sqrt(((ue[:, None, :] - bs[None, :, :]) ** 2).sum(-1).astype(np.float64))
    loss = np.where(d > 0, 38 + 30 * np.log10(np.maximum(d, 1e-300)), 0.0)
    p = 0.1 * 10 ** ((2 - loss - fade) / 10)
    noise = 10 ** (-12.1) * 1e-3
    want = np.empty_like(p)
    for b in range(n_bs):
        want[:, b] = 10 * np.log10(p[:, b] / (noise + np.delete(p, b, axis=1).sum(axis=1)))
    assert np.max(np.abs(got - want)) < 1e-9
    # zero path loss dominates every BS that is NOT on the same cell (co-located BSs differ by their fading only)
    apart = d[0] > 0
    assert np.all(got[0, 0] > got[0, apart])
    assert got[0].argmax() in np.flatnonzero(~apart)


@settings(max_examples=100, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), steps=st.integers(1, 12))
def test_handover_fifo_and_new_outage_semantics(orc, seed, steps):
    """UpdateDroneNet decisions (channel.py:145-176) against a literal Python replay: FIFO depth 1 -> 2 -> 3 then shift,
    handover iff the FIFO is constant, differs from the current cell and beats it by more than 1 dB, serving SINR read
    BEFORE the handover, n_out = UEs newly at or below 0 dB."""
    rs = np.random.RandomState(seed)
    n_ue, n_bs = 6, 3
    cfg = orc.default_cfg(n_bs, n_ue, 20, 1)
    L = orc.lib()
    ch = L.orc_chan_create(n_ue, n_bs)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    try:
        s0 = np.ascontiguousarray(rs.normal(3, 6, size=(n_ue, n_bs)))
        L.orc_chan_reset(C.byref(cfg), ch, dp(s0))
        cur = s0.argmax(1)
        fifo = [cur.copy()]
        out_prev = s0.max(1) <= 0
        for _ in range(steps):
            s = np.ascontiguousarray(np.round(rs.normal(3, 6, size=(n_ue, n_bs)), 1))   # rounded: exact ties happen
            ms, no, nh = C.c_double(), C.c_int32(), C.c_int32()
            L.orc_chan_update(C.byref(cfg), ch, dp(s), C.byref(ms), C.byref(no), C.byref(nh))
            best, bestS = s.argmax(1), s.max(1)
            curS = s[np.arange(n_ue), cur]
            fifo.append(best.copy())
            fifo = fifo[-3:]
            same = np.all(np.stack(fifo) == fifo[0], axis=0)
            ho = same & (cur != best) & (bestS - curS > 1)
            cur = np.where(ho, best, cur)
            out = curS <= 0
            assert nh.value == int(ho.sum()) and no.value == int((out & ~out_prev).sum())
            assert abs(ms.value - curS.mean()) < 1e-12
            assert np.array_equal(np.ctypeslib.as_array(ch.contents.cur, shape=(n_ue,)), cur)
            out_prev = out
    finally:
        L.orc_chan_destroy(ch)


@settings(max_examples=50, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1))
def test_state_orientation_reward_clamp_and_done(orc, seed):
    """state[0, x, y] = #BS, state[1+b, x, y] = #UE served by b (mobile_env.py:169-170, channel.py:404-406);
    r = max(meanSINR/20 - nOut/nUE, -1) (mobile_env.py:163-167,189); done at step_n >= MAXSTEP (:186-187)."""
    cfg = orc.default_cfg(max_step=3)
    o = orc.OracleEnv(cfg, seed=seed, env_id=1)
    o.reset()
    rs = np.random.RandomState(seed)
    for t in range(4):
        s, r, d, info = o.step(int(rs.randint(625)))
        want = np.zeros_like(s)
        for (x, y) in o.bs_xy:
            want[0, x, y] += 1
        for (x, y), b in zip(o.ue_xy, o.current_BS):
            want[1 + b, x, y] += 1
        assert np.array_equal(s, want)
        assert r == max(info["mean_sinr"] / 20 + (-1.0 * info["n_out"] / 40), -1) and r >= -1
        assert info["r_dissect"] == [info["mean_sinr"] / 20, -1.0 * info["n_out"] / 40]
        assert d == (t + 1 >= 3) and info["step_n"] == t + 1
