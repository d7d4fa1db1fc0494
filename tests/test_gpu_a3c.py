"""GPU tests of the actor-critic learner built on the env step path: the sparse first layer,
the hand-written backward, the fused RMSProp step, the n-step targets and the trainer loop, each against a plain
PyTorch restatement of the reference's TF1 graph (main.py:64-78,143-156,217-227,300-301) on the DENSE observation.
TensorFlow is absent, so these restatements are the learner's oracle: parity unpinned by the reference.
Tolerances: float32 arithmetic vs a float64 reference -- 1e-5 absolute on probabilities / values, 1e-4 relative
(of the largest gradient entry) on gradients."""
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


def _dense(idx, n_s):
    d = torch.zeros((idx.shape[0], n_s), dtype=torch.float64, device=idx.device)
    d.scatter_add_(1, idx.long(), torch.ones(idx.shape, dtype=torch.float64, device=idx.device))
    return d


def _ref_params(net):
    """float64 leaf copies of the net's parameters in the reference's per-net layout"""
    H, p = net.h, net.p
    names = dict(la=p["W1"][:, :H], la_b=p["b1"][:H], la2=p["Wa2"], la2_b=p["ba2"], ap=p["Wa3"], ap_b=p["ba3"],
                 lc=p["W1"][:, H:], lc_b=p["b1"][H:], lc2=p["Wc2"], lc2_b=p["bc2"], v=p["Wc3"], v_b=p["bc3"])
    return {k: t.detach().double().clone().requires_grad_(True) for k, t in names.items()}


def _ref_forward(P, s):
    r6 = lambda x: torch.clamp(x, 0, 6)  # noqa: E731
    l_a = r6(r6(s @ P["la"] + P["la_b"]) @ P["la2"] + P["la2_b"])
    a_prob = torch.softmax(l_a @ P["ap"] + P["ap_b"], dim=1)                    # main.py:147-149
    l_c = r6(r6(s @ P["lc"] + P["lc_b"]) @ P["lc2"] + P["lc2_b"])
    v = l_c @ P["v"] + P["v_b"]                                                  # main.py:151-153
    return a_prob, v


def _rand_idx(M, K, n_s, seed, dup=True):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, n_s, (M, K), generator=g, dtype=torch.int32)
    if dup:
        idx[:, 1] = idx[:, 0]                                                   # two UEs on one cell: count 2
    return idx.cuda()


def test_obs_idx_is_the_sparse_form_of_the_observation(pkg):
    env = pkg.BatchedMobiEnvironment(16, 4, 40, 100, "group", seed=3)
    obs = env.reset()
    assert torch.equal(_dense(env.obs_idx, 50000).float().view_as(obs), obs)
    for t in range(12):
        obs, r, d, info = env.step(np.random.RandomState(t).randint(0, 625, size=16))
        assert torch.equal(_dense(info["obs_idx"], 50000).float().view_as(obs), obs), t
    env2 = pkg.BatchedMobiEnvironment(16, 4, 40, 100, "group", seed=3, obs="none")  # no dense observation at all
    env3 = pkg.BatchedMobiEnvironment(16, 4, 40, 100, "group", seed=3)
    env2.reset()
    env3.reset()
    assert env2.obs is None and torch.equal(env2.obs_idx, env3.obs_idx)


def test_forward_matches_dense_reference(pkg):
    from drl_uav_cellularnet_b200.a3c import ACNet
    net = ACNet(50000, 625, "cuda:0")
    with torch.no_grad():
        net.p["b1"].normal_(0, 0.5)
        net.p["ba2"].normal_(0, 0.5)
        net.p["bc3"].fill_(0.3)
    idx = _rand_idx(64, 44, 50000, 1)
    prob, v, _ = net.forward(idx)
    P = _ref_params(net)
    rp, rv = _ref_forward(P, _dense(idx, 50000))
    assert float((prob.double() - rp.detach()).abs().max()) < 1e-5
    assert float((v.double() - rv.detach().squeeze(1)).abs().max()) < 1e-4
    assert torch.equal(net.greedy_action(idx), prob.argmax(1))
    a = net.choose_action(idx, seed=7)
    assert a.shape == (64,) and int(a.min()) >= 0 and int(a.max()) < 625
    a2 = net.choose_action(idx, seed=7)                       # the call counter is part of the key: fresh draws
    assert not torch.equal(a, a2)
    # the draws follow the probabilities: with one action's logit pushed up every row picks it
    with torch.no_grad():
        net.p["ba3"][17] += 50.0
    assert bool((net.choose_action(idx, seed=1) == 17).all())


def test_gradients_match_autograd_of_the_reference_losses(pkg):
    from drl_uav_cellularnet_b200.a3c import ACNet, ENTROPY_BETA
    net = ACNet(50000, 625, "cuda:0")
    M = 96
    idx = _rand_idx(M, 44, 50000, 2)
    g = torch.Generator().manual_seed(5)
    a_his = torch.randint(0, 625, (M,), generator=g).cuda()
    v_target = torch.randn(M, generator=g).cuda()
    a_loss, c_loss = net.accumulate_grads(idx, a_his, v_target)
    # reference: the TF graph of main.py:64-78 on the dense observation, float64, autograd
    P = _ref_params(net)
    a_prob, v = _ref_forward(P, _dense(idx, 50000))
    td = v_target.double().unsqueeze(1) - v
    rc_loss = (td ** 2).mean()
    log_prob = (torch.log(a_prob + 1e-5) * torch.nn.functional.one_hot(a_his, 625)).sum(1, keepdim=True)
    entropy = -(a_prob * torch.log(a_prob + 1e-5)).sum(1, keepdim=True)
    ra_loss = (-(ENTROPY_BETA * entropy + log_prob * td.detach())).mean()
    ga = torch.autograd.grad(ra_loss, [P[k] for k in ("la", "la_b", "la2", "la2_b", "ap", "ap_b")])
    gc = torch.autograd.grad(rc_loss, [P[k] for k in ("lc", "lc_b", "lc2", "lc2_b", "v", "v_b")])
    assert abs(float(a_loss) - float(ra_loss.detach())) < 1e-5
    assert abs(float(c_loss) - float(rc_loss.detach())) < 1e-5 * max(1, float(rc_loss.detach()))
    H, G = net.h, net.g
    mine = [G["W1"][:, :H], G["b1"][:H], G["Wa2"], G["ba2"], G["Wa3"], G["ba3"],
            G["W1"][:, H:], G["b1"][H:], G["Wc2"], G["bc2"], G["Wc3"], G["bc3"]]
    for m, r in zip(mine, list(ga) + list(gc)):
        scale = float(r.abs().max())
        assert scale > 0
        assert float((m.double() - r).abs().max()) <= 1e-4 * scale, (m.shape, scale)
    # accumulation: a second call adds the same gradient again
    before = net.grad.clone()
    net.accumulate_grads(idx, a_his, v_target)
    assert torch.allclose(net.grad, 2 * before, rtol=1e-4, atol=1e-7)


def test_rmsprop_matches_tf1_update_rule(pkg):
    from drl_uav_cellularnet_b200.a3c import ACNet, RMS_DECAY, RMS_EPS
    net = ACNet(1000, 25, "cuda:0", hidden=8)
    rs = np.random.RandomState(0)
    p0 = net.flat.cpu().numpy().astype(np.float64)
    ms = np.ones_like(p0)
    p = p0.copy()
    for it in range(3):
        gnp = rs.normal(0, 1e-2, size=p.shape).astype(np.float32)
        net.grad.copy_(torch.from_numpy(gnp))
        net.apply_grads(1e-4, world_size=2)
        gg = gnp.astype(np.float64) / 2                                  # averaged over 2 ranks
        ms = RMS_DECAY * ms + (1 - RMS_DECAY) * gg * gg                  # tf.train.RMSPropOptimizer, momentum 0
        p = p - 1e-4 * gg / np.sqrt(ms + RMS_EPS)
        assert float(net.grad.abs().max()) == 0.0                        # zeroed for the next accumulation
    assert np.max(np.abs(net.flat.cpu().numpy() - p)) < 1e-6
    assert np.max(np.abs(net.ms.cpu().numpy() - ms)) < 1e-6


def test_n_step_targets_match_worker_loop(pkg):
    from drl_uav_cellularnet_b200.a3c import GAMMA, n_step_targets
    T, E = 10, 7
    rs = np.random.RandomState(4)
    r = rs.normal(size=(T, E))
    done = rs.rand(T, E) < 0.15
    vb = rs.normal(size=E)
    got = n_step_targets(torch.from_numpy(r).cuda(), torch.from_numpy(done).cuda(), torch.from_numpy(vb).cuda()).cpu().numpy()
    for e in range(E):
        # main.py:212-238 flushes the buffer at `done` with v_s_ = 0 and at the end of the rollout with v(s')
        t0 = 0
        for t in range(T):
            if done[t, e] or t == T - 1:
                v_s_ = 0.0 if done[t, e] else vb[e]
                tgt = []
                for rr in r[t0:t + 1, e][::-1]:
                    v_s_ = rr + GAMMA * v_s_
                    tgt.append(v_s_)
                tgt.reverse()
                assert np.allclose(got[t0:t + 1, e], tgt, rtol=0, atol=1e-12), (e, t0, t)
                t0 = t + 1
    # the float32 kernel (uavnet_nstep_targets) against the same recursion
    got32 = n_step_targets(torch.from_numpy(r).float().cuda(), torch.from_numpy(done).cuda(), torch.from_numpy(vb).float().cuda())
    assert got32.dtype == torch.float32 and np.allclose(got32.cpu().numpy(), got, rtol=0, atol=1e-5)


def test_rollout_states_are_written_in_place_by_the_env(pkg):
    """The trainer's rollout storage: slot t holds the sparse observation the policy saw at step t (written by the env
    itself through bind_obs_idx, masked resets included), slot T the bootstrap state = slot 0 of the next rollout."""
    from drl_uav_cellularnet_b200.a3c import A3CTrainer, ACNet
    E = 32
    env = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=5, obs="none", max_step=7)
    ref = pkg.BatchedMobiEnvironment(E, 4, 40, 100, "group", seed=5, obs="none", max_step=7)
    net = ACNet(env.observation_space_dim, env.action_space_dim, env.device)
    tr = A3CTrainer(env, net, seed=3)
    ref.reset()
    for it in range(2):
        first = tr.buf_idx[0].clone()
        tr.rollout()
        assert torch.equal(first, ref.obs_idx), "slot 0 is not the state the rollout started from"
        for t in range(tr.T):                                      # replay the sampled actions on an independent env
            assert torch.equal(tr.buf_idx[t] if t else first, ref.obs_idx), (it, t)
            _, r, done, _ = ref.step(tr.buf_a[t])
            assert torch.equal(done, tr.buf_done[t]) and torch.allclose(r.float(), tr.buf_r[t])
            ref.reset(env_mask=ref.done_u8)
        assert torch.equal(tr.buf_idx[tr.T], ref.obs_idx) and torch.equal(tr.buf_idx[0], ref.obs_idx)
    assert tr.buf_done.any(), "max_step 7 < 20 steps: some episode must have ended"
    assert env.check() == 0 and ref.check() == 0


def test_actor_npz_round_trip(pkg, tmp_path):
    from drl_uav_cellularnet_b200.a3c import ACNet
    a, b = ACNet(500, 25, "cuda:0", hidden=8), ACNet(500, 25, "cuda:0", hidden=8, seed=9)
    path = str(tmp_path / "Global_A_PARA.npz")
    a.save_actor_npz(path)
    arr = np.load(path, allow_pickle=True)["arr_0"]                       # main_test.py:15
    assert [x.shape for x in arr] == [(500, 8), (8,), (8, 8), (8,), (8, 25), (25,)]
    b.load_actor_npz(path)
    idx = _rand_idx(5, 6, 500, 3)
    assert torch.equal(a.forward(idx, "actor")[0], b.forward(idx, "actor")[0])


def test_grouped_rollout_on_streams_equals_the_single_handle_rollout(pkg):
    """Env handles with consecutive global ranges, each on its own stream: same rollout, same update as one handle
    (all draws are keyed by the global env id) -- eager and replayed as one CUDA graph."""
    from drl_uav_cellularnet_b200.a3c import A3CTrainer, ACNet
    E, G = 48, 3
    mk = lambda n, off: pkg.BatchedMobiEnvironment(n, 4, 40, 100, "group", seed=9, obs="none", env_offset=off, max_step=13)  # noqa: E731
    one = A3CTrainer(mk(E, 0), ACNet(50000, 625, "cuda:0"), seed=4)
    grp = A3CTrainer([mk(E // G, g * (E // G)) for g in range(G)], ACNet(50000, 625, "cuda:0"), seed=4)
    assert torch.equal(one.net.flat, grp.net.flat)
    for it in range(3):
        grp.net.flat.copy_(one.net.flat)          # gradient REDs arrive in any order: re-align the last bits
        grp.net.ms.copy_(one.net.ms)
        la, lc = one.train_iteration()
        ga, gc = grp.train_iteration()
        torch.cuda.synchronize()
        for name in ("buf_idx", "buf_a", "buf_r", "buf_done", "buf_h1", "buf_h2a", "buf_prob", "buf_vt"):
            assert torch.equal(getattr(one, name), getattr(grp, name)), (it, name)
        assert abs(float(la) - float(ga)) < 1e-6 and abs(float(lc) - float(gc)) < 1e-6
    assert torch.allclose(one.net.flat, grp.net.flat, rtol=0, atol=1e-6)      # gradient REDs arrive in any order
    with pytest.raises(ValueError):
        A3CTrainer([mk(8, 0), mk(8, 9)], grp.net)                             # ranges must be consecutive
    grp.capture(warmup=1)
    a, c = grp.train_iteration_graph()
    torch.cuda.synchronize()
    assert torch.isfinite(a) and torch.isfinite(c)
    for e in grp.envs:
        assert e.check() == 0


def test_trainer_iterations_run_and_learn_signal(pkg):
    """A few synchronous A3C iterations on 64 envs: finite losses, parameters move, gradients are consumed, episodes
    restart at MAXSTEP, and the critic loss on a fixed batch drops when the same batch is replayed (sanity of the
    sign conventions of the hand-written backward + RMSProp)."""
    from drl_uav_cellularnet_b200.a3c import A3CTrainer, ACNet
    env = pkg.BatchedMobiEnvironment(64, 4, 40, 100, "group", seed=1, obs="none", max_step=25)
    net = ACNet(env.observation_space_dim, env.action_space_dim, env.device)
    tr = A3CTrainer(env, net, seed=2)
    p0 = net.flat.clone()
    saw_done = False
    for it in range(4):
        a_loss, c_loss = tr.train_iteration()
        assert torch.isfinite(a_loss) and torch.isfinite(c_loss)
        saw_done = saw_done or bool(tr.buf_done.any())
    assert float((net.flat - p0).abs().max()) > 0 and float(net.grad.abs().max()) == 0.0
    assert int(env.step_n.max()) == 15 and saw_done                            # 40 steps, MAXSTEP 25: episodes restarted
    M = tr.T * tr.E
    idx, a, vt = tr.buf_idx[:tr.T].reshape(M, -1), tr.buf_a.view(M), torch.zeros(M, device=env.device)
    first = last = None
    for it in range(60):
        _, c_loss = net.accumulate_grads(idx, a, vt)
        net.apply_grads(1e-4)
        first = float(c_loss) if first is None else first
        last = float(c_loss)
    assert last < first
    assert env.check() == 0


def test_evaluation_driver_writes_the_reference_files(pkg, tmp_path):
    """main_test.py: load Global_A_PARA.npz, replay a trace with greedy actions, write the nine .npy files."""
    from oracle import mobi_oracle as orc
    from drl_uav_cellularnet_b200.a3c import ACNet
    from drl_uav_cellularnet_b200.evaluate import FILES, load_ac_net, run_test
    cfg = orc.default_cfg()
    trace = orc.make_trace(cfg, 3, 0, 40)
    src = ACNet(50000, 625, "cuda:0", seed=11)
    npz = str(tmp_path / "Global_A_PARA.npz")
    src.save_actor_npz(npz)
    net = load_ac_net(npz, 50000, 625, "cuda:0")
    out = run_test(net, trace, str(tmp_path / "test"), max_step=30, seed=4)
    shapes = dict(reward=(31,), decomposed_reward=(31, 2), sinr=(31, 40), time=(31,), outage_fraction=(31,),
                  ue_location=(31, 40, 2), bs_location=(31, 4, 3), action=(31, 4), sinr_area=(2, 100, 100))
    for k in FILES:
        a = np.load(str(tmp_path / "test" / (k + ".npy")))
        assert a.shape == shapes[k], (k, a.shape)
        assert np.array_equal(a, out[k])
    # the recorded episode is what stepping the env with the same greedy policy gives
    env = pkg.BatchedMobiEnvironment(1, 4, 40, 100, "read_trace", trace=trace, precision="fp64", obs="none", seed=4)
    env.reset()
    for t in range(31):
        _, r, _, info = env.step(net.greedy_action(env.obs_idx))
        assert float(r[0]) == out["reward"][t]
        assert np.array_equal(info["ue_xy"][0].cpu().numpy(), out["ue_location"][t]) and np.array_equal(out["ue_location"][t], trace[t])
        assert np.allclose(out["reward"][t], max(out["decomposed_reward"][t].sum(), -1.0), rtol=0, atol=1e-12)
    # ... and what the ORACLE gives (config[0]: main_test.py's loop on the C restatement of the reference): the oracle env
    # replays the same trace with the same Philox fading; at every step the greedy action is recomputed in float64 from the
    # oracle's own dense state (argmax of the MLP of main.py:147-149, main_test.py:68) and must be the action the driver took
    # (unless the top-2 probabilities are within fp32 reach of each other), and the driver's recorded reward / SINR /
    # outage / locations / per-BS actions must be the oracle's.
    P = {k: v.double().cpu().numpy() for k, v in (("W1", net.p["W1"][:, :net.h]), ("b1", net.p["b1"][:net.h]), ("W2", net.p["Wa2"]),
                                                   ("b2", net.p["ba2"]), ("W3", net.p["Wa3"]), ("b3", net.p["ba3"]))}
    oenv = orc.OracleEnv(cfg, mobility=orc.MOB_TRACE, fading=orc.FADE_PHILOX, seed=4, env_id=0, trace=trace)
    state = oenv.reset()
    for t in range(31):
        x = state.reshape(-1)                                             # np.ravel(s), main_test.py:54,100: [plane, x, y] C order
        nz = np.nonzero(x)[0]
        h = np.clip(x[nz] @ P["W1"][nz] + P["b1"], 0, 6)
        h = np.clip(h @ P["W2"] + P["b2"], 0, 6)
        logits = h @ P["W3"] + P["b3"]
        top = np.sort(logits)[-2:]
        taken = int(out["action"][t] @ np.array([125, 25, 5, 1]))
        if top[1] - top[0] > 1e-4:
            assert int(np.argmax(logits)) == taken, t
        state, r, d, oi = oenv.step(taken)
        assert abs(out["reward"][t] - r) <= 1e-9 * max(1.0, abs(r)), t
        assert np.array_equal(out["ue_location"][t], oenv.ue_xy) and np.array_equal(out["bs_location"][t][:, :2], oenv.bs_xy), t
        assert np.max(np.abs(out["sinr"][t] - oenv.current_BS_sinr)) < 1e-9, t
        assert abs(out["outage_fraction"][t] - oi["n_out"] / 40.0) < 1e-15, t
        assert out["action"][t].tolist() == [float(v) for v in oi["digits"]], t


def test_reset_launches_are_skipped_only_when_no_episode_can_end(pkg):
    """The trainer launches the masked reset only at steps where an episode can end (host-side bound on step_n) -- eagerly and
    through its two CUDA graphs -- and the envs still restart exactly when they are done: with MAXSTEP = 25 and rollouts of
    10 steps, episodes end inside every third rollout; step counters, rewards and finished-episode returns equal those of a
    trainer that launches the reset after every step."""
    from drl_uav_cellularnet_b200.a3c import A3CTrainer, ACNet

    def run(mode):
        env = pkg.BatchedMobiEnvironment(32, 4, 40, 100, "group", seed=9, obs="none", max_step=25)
        net = ACNet(env.observation_space_dim, env.action_space_dim, "cuda:0", hidden=16, precision="fp32")
        tr = A3CTrainer(env, net, seed=3)
        if mode == "graph":
            tr.capture(warmup=1)
        rows, launches0 = [], env.launch_count
        for it in range(9):
            if mode == "always":
                tr.train_iteration(resets=True)
            elif mode == "graph":
                tr.train_iteration_graph()
            else:
                tr.train_iteration()
            torch.cuda.synchronize()
            rows.append((env.step_n.clone(), tr.buf_r.clone(), tr.buf_done.clone(), tr.ep_finished.clone(), net.flat.clone()))
        return rows, env.launch_count - launches0

    ref, n_ref = run("always")
    for mode in ("auto", "graph"):
        got, n = run(mode)
        skip = 1 if mode == "graph" else 0                              # the graph run made one eager warm-up iteration first
        for it in range(9 - skip):
            a, b = got[it], ref[it + skip]
            assert torch.equal(a[0], b[0]), (mode, it)                    # step counters: resets happened when due
            assert torch.equal(a[2], b[2]), (mode, it)                    # ... and the done flags of the rollout
            assert torch.equal(torch.isnan(a[3]), torch.isnan(b[3])), (mode, it)      # which envs have finished an episode
            if it == 0 and skip == 0:
                # same rollout and same update (the gradient sums are ordered by atomics: equal up to fp32 rounding; later
                # iterations may sample different actions from parameters that differ in the last bit)
                assert torch.equal(a[1], b[1]), mode
                assert torch.allclose(a[4], b[4], rtol=0, atol=1e-6), mode
        if mode == "auto":
            assert n < n_ref                                             # and fewer launches were made


def test_p2p_push_world_size_1_equals_rmsprop(pkg):
    """The peer-memory push (uavnet_p2p_push) with a single rank is the plain RMSProp step: same parameters bit for
    bit, gradients zeroed, buffers living in IPC-shareable allocations wrapped as torch tensors."""
    from drl_uav_cellularnet_b200.a3c import ACNet
    a, b = ACNet(2000, 25, "cuda:0", hidden=8), ACNet(2000, 25, "cuda:0", hidden=8)
    b.enable_p2p()
    assert torch.equal(a.flat, b.flat)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(3):
        grad = torch.randn(a.n_flat, device="cuda", generator=g) * 1e-2
        a.grad.copy_(grad)
        b.grad.copy_(grad)
        a.apply_grads(1e-4)
        b.apply_grads(1e-4)
        assert torch.equal(a.flat, b.flat) and torch.equal(a.ms, b.ms), it
        assert float(b.grad.abs().max()) == 0.0
    idx = _rand_idx(4, 6, 2000, 0)
    assert torch.equal(a.forward(idx)[0], b.forward(idx)[0])           # the views follow the new buffers
    assert b.p2p_status() == (3, False)
    b.close_p2p()
    assert torch.equal(a.flat, b.flat)


def test_p2p_push_two_ranks_equals_nccl_path():
    """Needs two GPUs (skipped on the single-GPU box): torchrun profiles/p2p_check.py -- the fused peer-memory push
    gives bit-identical parameters to NCCL all-reduce + RMSProp at 2 ranks, on every rank."""
    import json
    import os
    import subprocess
    import sys
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "profiles", "p2p_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["max_abs_param_diff_vs_nccl_path"] == 0.0
    assert d["param_sums_per_rank"][0] == d["param_sums_per_rank"][1]
    assert not d["flag_wait_gave_up"] and d["pushes"] > 30
    assert d["ms_slot_sums_after_close_per_rank"][0] == d["ms_slot_sums_after_close_per_rank"][1]


def test_p2p_push_two_processes_on_one_gpu():
    """The peer-memory push between two PROCESSES that share cuda:0 (runs on the single-GPU box): rendezvous and the
    reference all-reduce through gloo, the push itself through CUDA IPC mappings of the other process's gradient /
    parameter / flag buffers -- the flag protocol, the fixed summation order and the gathered optimiser slots are the
    same code as between GPUs.  Parameters bit-identical to all-reduce + uavnet_rmsprop on both ranks, no flag wait
    gave up, the RMSProp slots of both ranks agree after close_p2p()."""
    import json
    import os
    import subprocess
    import sys
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for split in (False, True):                       # one push per update / the two-part push of the trainer
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29579", os.path.join(root, "profiles", "p2p_check.py"), "--same-gpu", "--n-s", "20000", "--hidden", "64",
               "--iters", "5"] + (["--split"] if split else [])
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        d = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        assert d["world"] == 2 and d["max_abs_param_diff_vs_nccl_path"] == 0.0 and d["split"] == split
        assert d["param_sums_per_rank"][0] == d["param_sums_per_rank"][1]
        assert not d["flag_wait_gave_up"] and d["pushes"] >= 9
        assert d["ms_slot_sums_after_close_per_rank"][0] == d["ms_slot_sums_after_close_per_rank"][1]
        assert d["ms_slot_max_abs_diff_vs_nccl_path"] == 0.0          # every element's RMSProp slot came back from its owner


@pytest.mark.parametrize("M,K,R,H,passes", [(300, 44, 5000, 400, 2), (64, 7, 50, 16, 1), (2048, 44, 50000, 400, 4), (10, 3, 4, 8, 2)])
def test_sparse_bwd_gather_equals_scatter_and_float64(pkg, M, K, R, H, passes):
    """First-layer weight gradient, gather side (counting sort of the row indices + one plain sum per row) against the
    scatter side (float REDs) and a float64 index_add: same result up to fp32 summation order, duplicates inside a sample
    counted twice, rows nobody touches left alone (the buffer accumulates)."""
    import ctypes as C
    from drl_uav_cellularnet_b200 import _native as N
    L = N.lib()
    g = torch.Generator().manual_seed(M + K)
    idx = torch.randint(0, R, (M, K), generator=g, dtype=torch.int32)
    idx[:, 1] = idx[:, 0]                                                     # a duplicate in every sample
    idx = idx.cuda()
    dpre = torch.randn(M, H, generator=g).cuda()
    dpre[dpre.abs() < 0.5] = 0.0                                              # relu6' zeros
    base = torch.randn(R, H, generator=g).cuda()
    ref = base.double()
    ref.index_add_(0, idx.reshape(-1).long(), dpre.double().repeat_interleave(K, dim=0))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    a, b = base.clone(), base.clone()
    assert L.uavnet_sparse_bwd(C.c_void_p(idx.data_ptr()), M, K, R, C.c_void_p(dpre.data_ptr()), H, C.c_void_p(a.data_ptr()), st) == 0
    nb = L.uavnet_sparse_bwd_gather_workspace(M, K, R)
    ws = torch.empty(nb // 4, dtype=torch.int32, device="cuda")
    assert L.uavnet_sparse_bwd_gather(C.c_void_p(idx.data_ptr()), M, K, R, C.c_void_p(dpre.data_ptr()), H, C.c_void_p(b.data_ptr()),
                                      C.c_void_p(ws.data_ptr()), passes, st) == 0
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert float((a.double() - ref).abs().max()) < 1e-5 * scale
    assert float((b.double() - ref).abs().max()) < 1e-5 * scale
    untouched = torch.ones(R, dtype=torch.bool, device="cuda")
    untouched[idx.reshape(-1).long()] = False
    assert torch.equal(b[untouched], base[untouched])


def test_softmax_sample_kernel_matches_inverse_cdf(pkg):
    """uavnet_softmax_sample: probabilities = torch.softmax (1e-6), and the action is exactly the inverse-CDF pick for the
    Philox uniform of (seed, row, counter) -- recomputed on the host with the oracle's Philox -- and the empirical
    action frequencies over many draws follow the probabilities."""
    from oracle import mobi_oracle as orc
    from drl_uav_cellularnet_b200.a3c import ACNet
    net = ACNet(2000, 625, "cuda:0", hidden=16)
    g = torch.Generator(device="cuda").manual_seed(3)
    M, seed, row0 = 300, 99, 1000
    h2a = torch.rand((M, 16), device="cuda", generator=g) * 6
    ctr = torch.tensor([5], dtype=torch.int32, device="cuda")
    prob, act = net.sample_head(h2a, seed, row0, ctr, 2)
    ref = torch.softmax(torch.addmm(net.p["ba3"], h2a, net.p["Wa3"]), dim=1)
    assert float((prob - ref).abs().max()) < 1e-6
    p64 = prob.double().cpu().numpy()
    for m in range(0, M, 7):
        u, _ = orc.philox_uniform2(seed, row0 + m, 0, 7, 12)               # counter 5 + 2, DOM_SAMPLE
        cdf = np.cumsum(prob[m].cpu().numpy().astype(np.float32), dtype=np.float32)
        want = int(np.searchsorted(cdf, np.float32(u), side="right"))
        got = int(act[m])
        assert abs(got - min(want, 624)) <= 1, (m, got, want)              # float32 scan order may shift a boundary pick
        assert p64[m, got] > 0
    # frequencies: one fixed distribution, 40 000 draws through the counter
    hh = h2a[:1].expand(40000, 16).contiguous()
    _, a2 = net.sample_head(hh, seed, 0, ctr, 0)
    freq = torch.bincount(a2, minlength=625).double().cpu().numpy() / 40000
    p0 = p64[0]
    assert np.abs(freq - p0).max() < 5 * np.sqrt(p0.max() / 40000) + 1e-3
    _, a3 = net.sample_head(hh, seed, 0, ctr, 0)
    assert torch.equal(a2, a3)                                             # same counters, same draws
    ctr += 1
    _, a4 = net.sample_head(hh, seed, 0, ctr, 0)
    assert not torch.equal(a2, a4)


def test_small_net_against_the_oracles_known_answers(pkg):
    """ACNet(600, 25, hidden=40) against tests/golden/acnet_oracle_vectors.npz (oracle/acnet_oracle.py, float64):
    probabilities / values within 1e-5, losses within 1e-5, every gradient within 1e-4 of its largest entry."""
    import os
    from drl_uav_cellularnet_b200.a3c import ACNet
    w = np.load(os.path.join(os.path.dirname(__file__), "golden", "acnet_oracle_vectors.npz"))
    n_s, n_a, H = int(w["n_s"]), int(w["n_a"]), int(w["hidden"])
    net = ACNet(n_s, n_a, "cuda:0", hidden=H)
    f32 = lambda k: torch.from_numpy(w[k].astype(np.float32)).cuda()  # noqa: E731
    with torch.no_grad():
        net.p["W1"][:, :H].copy_(f32("p_la")); net.p["W1"][:, H:].copy_(f32("p_lc"))
        net.p["b1"][:H].copy_(f32("p_la_b")); net.p["b1"][H:].copy_(f32("p_lc_b"))
        for mine, theirs in (("Wa2", "la2"), ("ba2", "la2_b"), ("Wa3", "ap"), ("ba3", "ap_b"), ("Wc2", "lc2"), ("bc2", "lc2_b"),
                             ("Wc3", "v"), ("bc3", "v_b")):
            net.p[mine].copy_(f32("p_" + theirs))
    idx = torch.from_numpy(w["idx"]).cuda()
    prob, v, _ = net.forward(idx)
    assert float((prob.double().cpu() - torch.from_numpy(w["a_prob"])).abs().max()) < 1e-5
    assert float((v.double().cpu() - torch.from_numpy(w["v"])).abs().max()) < 1e-4
    a_loss, c_loss = net.accumulate_grads(idx, torch.from_numpy(w["a_his"]).cuda(), f32("v_target"))
    assert abs(float(a_loss) - float(w["a_loss"])) < 1e-5 and abs(float(c_loss) - float(w["c_loss"])) < 1e-5 * max(1.0, float(w["c_loss"]))
    G = net.g
    pairs = [(G["W1"][:, :H], "la"), (G["b1"][:H], "la_b"), (G["Wa2"], "la2"), (G["ba2"], "la2_b"), (G["Wa3"], "ap"), (G["ba3"], "ap_b"),
             (G["W1"][:, H:], "lc"), (G["b1"][H:], "lc_b"), (G["Wc2"], "lc2"), (G["bc2"], "lc2_b"), (G["Wc3"], "v"), (G["bc3"], "v_b")]
    for mine, name in pairs:
        ref = torch.from_numpy(w["g_" + name]).reshape(mine.shape)
        scale = float(ref.abs().max()) + 1e-12
        assert float((mine.double().cpu() - ref).abs().max()) <= 1e-4 * scale, name
