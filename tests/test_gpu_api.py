"""GPU tests of the host-facing behaviour of the CUDA path (through the C-ABI library): the host-buffer entry
point, sharding invariance, error behaviour, masked reset, state clone, the fallback observation paths, and
size-independent properties at BASELINE.json's full sizes (config[1]: 4096 envs; config[3]: 32 BS x 2048 UE).
Need a B200: -m gpu."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import drl_uav_cellularnet_b200 as p
    return p


def _np(t):
    return t.detach().cpu().numpy()


def _actions(E, T, seed=3, n=625):
    return np.random.RandomState(seed).randint(0, n, size=(T, E))


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_equals_device_step(pkg, pinned):
    """uavenv_step_host (host actions in, host rewards / done / mean SINR / outage counts out; pinned buffers are
    written by the kernel directly, pageable ones through staging copies) == uavenv_step on device tensors."""
    E, T = 64, 12
    a = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=11)
    b = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=11)
    a.reset()
    b.reset()
    mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    act_h = mk(torch.zeros(E, dtype=torch.int64))
    rew_h, done_h = mk(torch.zeros(E, dtype=torch.float64)), mk(torch.zeros(E, dtype=torch.uint8))
    mean_h, nout_h = mk(torch.zeros(E, dtype=torch.float64)), mk(torch.zeros(E, dtype=torch.int32))
    acts = _actions(E, T)
    for t in range(T):
        act_h.copy_(torch.from_numpy(acts[t]))
        obs_a = a.step_host(act_h, rew_h, done_h, mean_h, nout_h)
        obs_b, r, d, info = b.step(acts[t])
        torch.cuda.synchronize()
        assert np.array_equal(rew_h.numpy(), _np(r)), t
        assert np.array_equal(done_h.numpy().astype(bool), _np(d)), t
        assert np.array_equal(mean_h.numpy(), _np(info["mean_sinr"])), t
        assert np.array_equal(nout_h.numpy(), _np(info["n_out"])), t
        assert torch.equal(obs_a, obs_b), t
    assert a.check() == 0 and b.check() == 0


def test_sharding_invariance(pkg):
    """Environments are keyed by GLOBAL id: one handle with 12 envs == three handles with 4 envs at offsets
    0 / 4 / 8 (how bench.py shards config[2] over ranks), for the fp32 and the fp64 kernels."""
    E, T, seed = 12, 25, 2026
    acts = _actions(E, T)
    for prec in ("fp32", "fp64"):
        full = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=seed, precision=prec)
        parts = [pkg.BatchedMobiEnvironment(4, 4, 40, 100, "group", seed=seed, precision=prec, env_offset=4 * i)
                 for i in range(3)]
        o_full = full.reset()
        o_parts = [p.reset() for p in parts]
        assert torch.equal(o_full, torch.cat(o_parts))
        for t in range(T):
            o_full, r, d, info = full.step(acts[t])
            outs = [p.step(acts[t][4 * i:4 * i + 4]) for i, p in enumerate(parts)]
            assert torch.equal(o_full, torch.cat([o[0] for o in outs])), (prec, t)
            assert torch.equal(r, torch.cat([o[1] for o in outs])), (prec, t)
            for k in ("serving", "serving_sinr", "ue_xy", "bs_xy", "n_out", "n_ho", "mean_sinr"):
                assert torch.equal(info[k], torch.cat([o[3][k] for o in outs])), (prec, t, k)


def test_invalid_action_is_rejected_and_state_kept(pkg):
    """An action >= 5**nBS (ValueError in the reference, ue_mobility.py:323-336) raises on check() and leaves the
    state of that env untouched; the other envs of the batch are stepped."""
    E = 4
    env = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=5, precision="fp64")
    ref = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=5, precision="fp64")
    env.reset()
    ref.reset()
    before = env.get_state()
    env.step(np.array([3, 625, 17, -1]))
    with pytest.raises(ValueError):
        env.check()
    ref.step(np.array([3, 0, 17, 0]))
    after, want = env.get_state(), ref.get_state()
    for k in after:
        assert np.array_equal(after[k][[1, 3]], before[k][[1, 3]]), k          # rejected envs: unchanged
        assert np.array_equal(after[k][[0, 2]], want[k][[0, 2]]), k            # the others: stepped
    assert env.check() == 0                                                    # flags are cleared by check()
    single = pkg.MobiEnvironment(4, 40, 100)
    single.reset()
    with pytest.raises(ValueError):
        single.step(625)
    with pytest.raises(ValueError):
        pkg.MobiEnvironment(4, 40, 100, "random_walk")                          # sys.exit at mobile_env.py:91


def test_trace_exhaustion_raises_index_error(pkg):
    """step_test past the end of the trace: IndexError as in the reference (mobile_env.py:203)."""
    trace = np.random.RandomState(0).randint(0, 100, size=(3, 40, 2))
    env = pkg.MobiEnvironment(4, 40, 100, "read_trace", trace=trace, fading="none")
    env.reset()
    for _ in range(3):
        env.step_test(624)
    with pytest.raises(IndexError):
        env.step_test(624)


def test_masked_reset_and_done(pkg):
    """reset(env_mask) touches only the selected envs (main.py:188-190 resets a worker's env when done);
    done is raised at step_n >= MAXSTEP (mobile_env.py:186-187) and step_n keeps counting."""
    E = 6
    env = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=9, max_step=5)
    env.reset()
    for t in range(5):
        obs, r, d, info = env.step(np.full(E, 624))
        assert bool(d.all()) == (t == 4)
    before = env.get_state()
    obs_before = env.obs.clone()
    mask = np.array([1, 0, 0, 1, 0, 1], dtype=np.uint8)
    env.reset(env_mask=mask)
    after = env.get_state()
    keep = mask == 0
    for k in after:
        assert np.array_equal(after[k][keep], before[k][keep]), k
    assert torch.equal(env.obs[torch.from_numpy(keep)], obs_before[torch.from_numpy(keep)])
    assert np.array_equal(after["counters"][:, 2], np.where(mask, 0, 5))       # step_n
    init_bs = np.array([[25, 25], [25, 75], [75, 25], [75, 75]])               # mobile_env.py:49-50
    assert np.array_equal(after["bs_xy"][mask == 1], np.broadcast_to(init_bs, (3, 4, 2)))


def test_state_clone_what_if_step(pkg):
    """copy.deepcopy(env) + look-ahead step_test(624) of gradient.py:14-17: get_state / set_state round trip."""
    env = pkg.BatchedMobiEnvironment(3, 4, 40, 100, "group", seed=21, precision="fp64")
    env.reset()
    env.step(np.array([1, 2, 3]))
    snap = env.get_state()
    o1, r1, _, i1 = env.step(np.full(3, 624))
    o1, r1, srv1 = o1.clone(), r1.clone(), i1["serving"].clone()
    env.step(np.array([5, 6, 7]))
    env.set_state(snap)
    o2, r2, _, i2 = env.step(np.full(3, 624))
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(srv1, i2["serving"])


def _choose_act_gradient(virtual_env_after_step):
    """gradient.py:19-35 on the virtual env AFTER its look-ahead step_test(624): per BS the direction whose UEs have the
    lowest mean serving SINR (nanargmin over x+, x-, y+, y-), joint action MSB first"""
    import warnings
    sinr, bs_loc, ue_loc = virtual_env_after_step
    act = np.zeros(len(bs_loc))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                      # mean of an empty slice -> nan, as in the reference
        for b in range(len(bs_loc)):
            g = np.array([np.mean(sinr[ue_loc[:, 0] > bs_loc[b][0]]), np.mean(sinr[ue_loc[:, 0] <= bs_loc[b][0]]),
                          np.mean(sinr[ue_loc[:, 1] > bs_loc[b][1]]), np.mean(sinr[ue_loc[:, 1] <= bs_loc[b][1]])])
            act[b] = np.nanargmin(g)
    return int(act[3] + act[2] * 5 + act[1] * 25 + act[0] * 125)


@pytest.mark.parametrize("precision", ["fp64", "fp32_guarded"])
def test_deepcopy_look_ahead_matches_cloned_oracle(pkg, precision):
    """gradient.py:14-37 through the drop-in class: `virtual_env = deepcopy(actual_env)`, `virtual_env.step_test(624)`,
    read `channel.current_BS_sinr / bsLoc / ueLoc`, derive the action, step the ACTUAL env with it -- against the oracle,
    whose clone is a second oracle env replaying the same history (draws are keyed by counters, so a replay is a clone).
    Decisions bit-exact; the actual env must not notice the virtual step."""
    import copy
    from oracle import mobi_oracle as orc
    seed, T = 99, 40
    cfg = orc.default_cfg()
    trace = orc.make_trace(cfg, 7, 3, 2 * T + 8)
    env = pkg.MobiEnvironment(4, 40, 100, "read_trace", trace=trace, precision=precision, seed=seed)

    def make_oracle(history):
        o = orc.OracleEnv(cfg, mobility=orc.MOB_TRACE, fading=orc.FADE_PHILOX, seed=seed, env_id=0, trace=trace)
        o.reset()
        for a in history:
            o.step(a)
        return o

    s = env.reset()
    oenv = make_oracle([])
    history = []
    for t in range(T):
        venv = copy.deepcopy(env)                                         # gradient.py:15
        assert venv is not env and venv._b is not env._b
        assert np.array_equal(venv.channel.current_BS_sinr, env.channel.current_BS_sinr)
        venv.step_test(624, False)                                        # :17 BSs stay, UEs move
        voe = make_oracle(history)
        voe.step(624)
        assert np.array_equal(venv.channel.current_BS, voe.current_BS), t
        assert np.array_equal(venv.ueLoc, voe.ue_xy) and np.array_equal(venv.bsLoc[:, :2], voe.bs_xy), t
        tol = 1e-9 if precision == "fp64" else 1e-3
        assert np.max(np.abs(venv.channel.current_BS_sinr - voe.current_BS_sinr)) < tol, t
        action = _choose_act_gradient((venv.channel.current_BS_sinr, venv.bsLoc, venv.ueLoc))
        o_action = _choose_act_gradient((voe.current_BS_sinr, voe.bs_xy, voe.ue_xy))
        if precision == "fp64":
            assert action == o_action, t
        s, r, d, info = env.step_test(action, False)                      # gradient.py:61
        os_, orw, od, oi = oenv.step(action)
        history.append(action)
        assert np.array_equal(env.channel.current_BS, oenv.current_BS), t
        assert np.array_equal(s, os_), t
        assert int(env._b.n_out[0]) == oi["n_out"] and int(env._b.n_ho[0]) == oi["n_ho"], t
        assert info.step_n == t + 1 == oi["step_n"]


def test_deepcopy_batched_group_env_is_independent(pkg):
    """the batched class clones too (group mobility: positions, group state, phase counters): twin and original produce
    the same next step, and stepping one leaves the other untouched"""
    import copy
    env = pkg.BatchedMobiEnvironment(5, 4, 40, 100, "group", seed=5, precision="fp32_guarded")
    env.reset()
    for t in range(3):
        env.step(np.arange(5) * 100 + t)
    twin = copy.deepcopy(env)
    assert torch.equal(twin.obs, env.obs) and torch.equal(twin.serving, env.serving)
    a = np.array([624, 0, 311, 17, 500])
    o1, r1, _, i1 = twin.step(a)
    before = env.get_state()
    twin.step(a)
    twin.step(a)
    after = env.get_state()
    for k in before:
        assert np.array_equal(before[k], after[k]), k
    o2, r2, _, i2 = env.step(a)
    twin2 = copy.deepcopy(env)
    assert torch.equal(twin2.reward, r2)
    # first step of the twin == the same step of the original
    env3 = pkg.BatchedMobiEnvironment(5, 4, 40, 100, "group", seed=5, precision="fp32_guarded")
    env3.reset()
    for t in range(3):
        env3.step(np.arange(5) * 100 + t)
    o3, r3, _, i3 = env3.step(a)
    assert torch.equal(r3, r2) and torch.equal(i3["serving"], i2["serving"]) and torch.equal(o3, o2)


@pytest.mark.parametrize("nBS,nUE,G", [(4, 40, 25), (3, 17, 10), (4, 40, 100), (6, 100, 64)])
def test_observation_paths_agree_with_oracle(pkg, nBS, nUE, G):
    """The dense observation through every code path -- TMA zero stream + count REDs (whole float4s per env),
    plain-store fallback (odd sizes), incremental updates -- against the oracle's state for the same env."""
    from oracle import mobi_oracle as orc
    seed, E, T = 31, 5, 30
    gs = [nUE // 4 + (1 if g < nUE % 4 else 0) for g in range(4)]
    kw = dict(seed=seed, precision="fp64", group_sizes=gs)
    if nBS != 4:
        kw["init_bs_xy"] = [[2 + (G - 4) * b // nBS, 2 + (G - 4) * ((b * 7) % nBS) // nBS] for b in range(nBS)]
    full = pkg.BatchedMobiEnvironment(E, nBS, nUE, G, "group", obs="f32", **kw)
    inc = pkg.BatchedMobiEnvironment(E, nBS, nUE, G, "group", obs="f32_incremental", **kw)
    plan = full.launch_plan
    assert (plan["tile_bytes"] > 0) == (((nBS + 1) * G * G) % 4 == 0), plan
    cfg = orc.default_cfg(nBS, nUE, G, 4)
    oenvs = [orc.OracleEnv(cfg, group_sizes=gs, init_bs_xy=kw.get("init_bs_xy"), seed=seed, env_id=e) for e in range(E)]
    want = np.stack([o.reset() for o in oenvs])
    assert np.array_equal(_np(full.reset()).astype(np.float64), want)
    assert np.array_equal(_np(inc.reset()).astype(np.float64), want)
    acts = np.random.RandomState(1).randint(0, 5, size=(T, E, nBS)).astype(np.uint8)
    for t in range(T):
        of, *_ = full.step(acts[t])
        oi, *_ = inc.step(acts[t])
        want = np.stack([oenvs[e].step(acts[t][e].astype(np.int32))[0] for e in range(E)])
        assert np.array_equal(_np(of).astype(np.float64), want), t
        assert np.array_equal(_np(oi).astype(np.float64), want), t
    assert full.check() == 0 and inc.check() == 0


def test_unaligned_observation_buffer_uses_fallback(pkg):
    """A caller-provided observation buffer that is not 16-byte aligned cannot be a TMA destination: the same
    call must fall back to plain stores and still produce the identical observation."""
    import ctypes as C
    from drl_uav_cellularnet_b200 import _native as N
    E = 3
    a = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=8)
    b = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=8)
    a.reset()
    b.reset()
    raw = torch.zeros(E * 50000 + 1, dtype=torch.float32, device=b.device)
    shifted = raw[1:]                                                           # base + 4 bytes
    assert shifted.data_ptr() % 16 == 4
    b._out.obs = C.c_void_p(shifted.data_ptr())
    act = np.array([7, 300, 624])
    oa, *_ = a.step(act)
    b.step(act)
    torch.cuda.synchronize()
    assert torch.equal(oa.reshape(-1), shifted)
    assert N.OK == 0


# ---------------------------------------------------------------------------------------------------------
def test_full_size_properties_config1(pkg):
    """config[1] sizes (4096 envs, 4 x 40, grid 100), fp32 kernels, 40 steps + a reset: properties that do not need
    the oracle -- BS plane sums to nBS and the association planes to nUE in every env, every count is where
    (serving, ue_xy, bs_xy) say it is, rewards obey the clamp, incremental observations equal rebuilt ones, and
    two handles with the same seed agree bit for bit (determinism across launches)."""
    E, T = 4096, 40
    a = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=123)
    b = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=123, obs="f32_incremental")
    a.reset()
    b.reset()
    acts = torch.from_numpy(_actions(E, T, seed=77)).cuda()
    ar = torch.arange(E, device="cuda")
    for t in range(T):
        obs, r, d, info = a.step(acts[t])
        obs_b, r_b, *_ = b.step(acts[t])
        assert torch.equal(obs, obs_b), t
        assert torch.equal(r, r_b), t
        assert torch.all(obs[:, 0].sum(dim=(1, 2)) == 4) and torch.all(obs[:, 1:].sum(dim=(1, 2, 3)) == 40), t
        rebuilt = torch.zeros_like(obs)
        ue, srv, bs = info["ue_xy"].long(), info["serving"].long(), info["bs_xy"].long()
        rebuilt.index_put_((ar[:, None].expand(E, 40), 1 + srv, ue[..., 0], ue[..., 1]),
                           torch.ones((), device="cuda"), accumulate=True)
        rebuilt.index_put_((ar[:, None].expand(E, 4), torch.zeros_like(bs[..., 0]), bs[..., 0], bs[..., 1]),
                           torch.ones((), device="cuda"), accumulate=True)
        assert torch.equal(obs, rebuilt), t
        want = torch.clamp(info["mean_sinr"] / 20 - info["n_out"].double() / 40, min=-1.0)
        assert torch.allclose(r, want, rtol=0, atol=1e-12), t
        assert int(info["n_out"].max()) <= 40 and int(info["n_out"].min()) >= 0
        if t == 20:
            a.reset()
            b.reset()
    assert a.check() == 0 and b.check() == 0


def test_full_size_properties_dense(pkg):
    """config[3] sizes (32 UAV-BS x 2048 UE, 64 of the 1024 envs to keep the test short), fp32 kernels: plane sums,
    exact count placement, serving BS = argmax of the SINR the kernel itself reports whenever a handover fired."""
    E, nBS, nUE, T = 64, 32, 2048, 6
    env = pkg.BatchedMobiEnvironment(E, nBS, nUE, 100, "group", seed=5, diagnostics=True)
    env.reset()
    rs = np.random.RandomState(2)
    ar = torch.arange(E, device="cuda")
    prev_srv = env.serving.clone()
    for t in range(T):
        digits = rs.randint(0, 5, size=(E, nBS)).astype(np.uint8)
        obs, r, d, info = env.step(digits)
        assert torch.all(obs[:, 0].sum(dim=(1, 2)) == nBS) and torch.all(obs[:, 1:].sum(dim=(1, 2, 3)) == nUE), t
        ue, srv = info["ue_xy"].long(), info["serving"].long()
        rebuilt = torch.zeros_like(obs[:, 1:])
        rebuilt.index_put_((ar[:, None].expand(E, nUE), srv, ue[..., 0], ue[..., 1]), torch.ones((), device="cuda"),
                           accumulate=True)
        assert torch.equal(obs[:, 1:], rebuilt), t
        changed = srv != prev_srv.long()
        best = env.sinr_all.argmax(dim=2)
        assert torch.equal(srv[changed], best[changed]), t                      # handovers go to the best server
        assert int(info["n_ho"].sum()) == int(changed.sum()), t
        prev_srv = info["serving"].clone()
    assert env.check() == 0


def test_coverage_map_matches_reference_fixture_and_oracle(pkg, golden_dir):
    """GetSinrInArea (channel.py:411-433; saved every 500 evaluation steps, main_test.py:89): float64 kernel vs the
    fixture recorded from the unmodified reference (injected draws, 1e-9 dB), Philox mode vs the oracle with the same
    counters (fp64 1e-9 dB, fp32 1e-3 dB), and the single-env shim's env.channel.GetSinrInArea."""
    import os
    from oracle import mobi_oracle as orc
    g = np.load(os.path.join(golden_dir, "ref_sinr_area.npz"))
    n = g["bs"].shape[0]
    env = pkg.BatchedMobiEnvironment(n, 4, 40, 100, "group", precision="fp64", seed=9)
    out = env.coverage_map(g["bs"], g["fading"].reshape(n, 99 * 99, 4))
    assert float(np.abs(_np(out) - g["sinr"]).max()) < 1e-9
    cfg = orc.default_cfg()
    for prec, tol in (("fp64", 1e-9), ("fp32", 1e-3)):
        e2 = pkg.BatchedMobiEnvironment(2, 4, 40, 100, "group", precision=prec, seed=77, env_offset=5)
        e2.reset()
        for call in range(2):                                             # the call number is part of the Philox counter
            got = _np(e2.coverage_map()).astype(np.float64)
            bs = e2.get_state()["bs_xy"]
            for e in range(2):
                want = orc.sinr_in_area(cfg, bs[e], by_bs=orc.philox_area_fading(cfg, 77, 5 + e, call))
                assert float(np.abs(got[e] - want).max()) < tol, (prec, call, e)
    single = pkg.MobiEnvironment(4, 40, 100, fading="none")
    single.reset()
    m = single.channel.GetSinrInArea(single.bsLoc)
    assert m.shape == (100, 100) and m.dtype == np.float64 and float(np.abs(m[0]).max()) == 0.0
    assert float(np.abs(m - orc.sinr_in_area(cfg, single.bsLoc)).max()) < 1e-9


@pytest.mark.parametrize("nBS,nUE,G,groups", [(1, 1, 4, [1]), (2, 3, 6, [2, 1]), (5, 33, 12, [11, 11, 11]), (27, 64, 40, [16] * 4)])
@pytest.mark.parametrize("precision", ["fp64", "fp32_guarded"])
def test_edge_sizes_match_oracle(pkg, nBS, nUE, G, groups, precision):
    """Smallest legal sizes, ragged groups, a BS count that leaves lanes of the 4-BSs-per-lane mapping empty and the
    largest BS count whose joint action still fits int64 (5**27 < 2**63): float64 and guarded-fp32 kernels vs the oracle,
    every decision and the observation exact (reward: 1e-9 in float64, the north star's 1e-5 relative or 2e-6 absolute in
    guarded fp32), with both action encodings (joint int64 and per-BS digits) giving the same step."""
    from oracle import mobi_oracle as orc
    seed, E, T = 17, 3, 12
    layout = [[1 + (b * 7) % (G - 1), 1 + (b * 3) % (G - 1)] for b in range(nBS)]
    kw = dict(seed=seed, precision=precision, group_sizes=groups, init_bs_xy=layout)
    a = pkg.BatchedMobiEnvironment(E, nBS, nUE, G, "group", **kw)
    b = pkg.BatchedMobiEnvironment(E, nBS, nUE, G, "group", **kw)
    rtol, atol = (1e-9, 1e-9) if precision == "fp64" else (1e-5, 2e-6)
    cfg = orc.default_cfg(nBS, nUE, G, len(groups))
    oenvs = [orc.OracleEnv(cfg, group_sizes=groups, init_bs_xy=layout, seed=seed, env_id=e) for e in range(E)]
    want = np.stack([o.reset() for o in oenvs])
    assert np.array_equal(_np(a.reset()).astype(np.float64), want)
    b.reset()
    rs = np.random.RandomState(3)
    for t in range(T):
        digits = rs.randint(0, 5, size=(E, nBS)).astype(np.uint8)
        joint = np.array([int(sum(int(d) * 5 ** (nBS - 1 - k) for k, d in enumerate(row))) for row in digits], dtype=np.int64)
        oa, ra, da, ia = a.step(digits)
        ob, rb, db, ib = b.step(joint)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(ia["serving"], ib["serving"]), t
        for e in range(E):
            s, r, d, oi = oenvs[e].step(digits[e].astype(np.int32))
            assert np.array_equal(_np(oa[e]).astype(np.float64), s), (t, e)
            assert np.array_equal(_np(ia["serving"][e]), oenvs[e].current_BS), (t, e)
            assert int(ia["n_out"][e]) == oi["n_out"] and int(ia["n_ho"][e]) == oi["n_ho"], (t, e)
            assert abs(float(ra[e]) - r) <= max(rtol * abs(r), atol), (t, e)
    assert a.check() == 0 and b.check() == 0


def test_joint_action_overflow_and_bad_sizes_are_rejected(pkg):
    """5**nBS overflows int64 beyond 27 BSs (mobile_env.py:104): joint actions are then rejected by validation (every
    int64 decodes to at most 27 digits, so any value is a legal action for nBS > 27 only through per-BS digits);
    impossible sizes fail at creation with ValueError."""
    env = pkg.BatchedMobiEnvironment(2, 28, 56, 60, "group", seed=1, group_sizes=[14] * 4,
                                     init_bs_xy=[[2 + 2 * b, 2 + (b * 5) % 50] for b in range(28)])
    env.reset()
    env.step(np.array([2 ** 62, 5], dtype=np.int64))                   # decodes into 28 base-5 digits: legal
    assert env.check() == 0
    env.step(np.array([-1, 5], dtype=np.int64))
    with pytest.raises(ValueError):
        env.check()
    for bad in (dict(nBS=33), dict(nBS=0), dict(nUE=0), dict(grid_n=3)):
        args = dict(nBS=4, nUE=40, grid_n=100)
        args.update(bad)
        with pytest.raises(ValueError):
            pkg.BatchedMobiEnvironment(2, args["nBS"], args["nUE"], args["grid_n"], "group")
    with pytest.raises(ValueError):
        pkg.BatchedMobiEnvironment(2, 4, 40, 100, "group", group_sizes=[10, 10, 10])      # does not sum to nUE
    with pytest.raises(ValueError):
        pkg.BatchedMobiEnvironment(1, 32, 64, 9000, "group", obs="none")                   # (nBS+1) G^2 >= 2^31


def test_long_moves_per_env_traces_and_incremental_masked_reset(pkg):
    """Less-travelled options against the oracle (float64 kernels): N_ACT = 9 (the long moves 5-8 of BS_move,
    ue_mobility.py:239-253), one trace PER env (uavenv_set_trace per_env), and incremental observations across a
    masked reset (the reset envs are rewritten in full, the others only patched)."""
    from oracle import mobi_oracle as orc
    E, T, seed = 4, 25, 66
    cfg = orc.default_cfg(n_act=9)
    # -- N_ACT = 9, group mode, incremental observation, masked reset half way
    env = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", precision="fp64", seed=seed, obs="f32_incremental", n_act=9)
    assert env.action_space_dim == 9 ** 4
    oenvs = [orc.OracleEnv(cfg, seed=seed, env_id=e) for e in range(E)]
    want = np.stack([o.reset() for o in oenvs])
    assert np.array_equal(_np(env.reset()).astype(np.float64), want)
    rs = np.random.RandomState(8)
    for t in range(T):
        digits = rs.randint(0, 9, size=(E, 4)).astype(np.uint8)
        obs, r, d, info = env.step(digits)
        for e in range(E):
            want[e] = oenvs[e].step(digits[e].astype(np.int32))[0]
            assert np.array_equal(_np(info["bs_xy"][e]), oenvs[e].bs_xy), (t, e)
        assert np.array_equal(_np(obs).astype(np.float64), want), t
        if t == 12:
            mask = np.array([1, 0, 0, 1], dtype=np.uint8)
            obs = env.reset(env_mask=mask)
            for e in range(E):
                if mask[e]:
                    want[e] = oenvs[e].reset()
            assert np.array_equal(_np(obs).astype(np.float64), want)
    assert env.check() == 0
    # -- one trace per env
    cfg5 = orc.default_cfg()
    traces = np.stack([orc.make_trace(cfg5, 100 + e, e, T + 1) for e in range(E)], axis=1)        # [T+1, E, 40, 2]
    tenv = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "read_trace", trace=traces, trace_per_env=True, precision="fp64",
                                      seed=seed)
    torcs = [orc.OracleEnv(cfg5, mobility=orc.MOB_TRACE, seed=seed, env_id=e, trace=traces[:, e], warmup_ticks=0) for e in range(E)]
    want = np.stack([o.reset() for o in torcs])
    assert np.array_equal(_np(tenv.reset()).astype(np.float64), want)
    for t in range(T):
        act = rs.randint(0, 625, size=E)
        obs, r, d, info = tenv.step(act)
        for e in range(E):
            s, rr, dd, oi = torcs[e].step(int(act[e]))
            assert np.array_equal(_np(obs[e]).astype(np.float64), s), (t, e)
            assert np.array_equal(_np(info["ue_xy"][e]), traces[t, e]), (t, e)
            assert int(info["n_out"][e]) == oi["n_out"] and abs(float(r[e]) - rr) < 1e-9
    assert tenv.check() == 0
