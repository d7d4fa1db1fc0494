"""oracle/acnet_oracle.py (the float64 restatement of the reference's actor-critic graph, main.py:64-80,143-156,
217-227,300-301) pinned on the CPU: its hand-written gradients against torch.autograd of the same graph, its targets
against the batched n_step_targets of the product's host code, RMSProp against the written-out TF1 rule.  TensorFlow is
absent and unpinned by the reference, so this is the strongest pin available for rows f1/f2 (see the oracle's header)."""
import numpy as np
import pytest
import torch

from oracle import acnet_oracle as orc


def _torch_losses(p, s, a_his, v_target, beta):
    """the TF graph of main.py:64-78 written with torch ops on the dense observation"""
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    x = torch.tensor(s, dtype=torch.float64)
    relu6 = lambda y: torch.clamp(y, 0.0, 6.0)  # noqa: E731
    l_a = relu6(x @ t["la"] + t["la_b"])
    l_a2 = relu6(l_a @ t["la2"] + t["la2_b"])
    a_prob = torch.softmax(l_a2 @ t["ap"] + t["ap_b"], dim=1)
    l_c = relu6(x @ t["lc"] + t["lc_b"])
    l_c2 = relu6(l_c @ t["lc2"] + t["lc2_b"])
    v = l_c2 @ t["v"] + t["v_b"]
    td = torch.tensor(v_target, dtype=torch.float64)[:, None] - v
    c_loss = (td ** 2).mean()
    onehot = torch.nn.functional.one_hot(torch.tensor(a_his), a_prob.shape[1]).double()
    log_prob = (torch.log(a_prob + 1e-5) * onehot).sum(1, keepdim=True)
    exp_v = log_prob * td.detach()
    entropy = -(a_prob * torch.log(a_prob + 1e-5)).sum(1, keepdim=True)
    a_loss = (-(beta * entropy + exp_v)).mean()
    ga = torch.autograd.grad(a_loss, [t[k] for k in orc.ACTOR], retain_graph=True)
    gc = torch.autograd.grad(c_loss, [t[k] for k in orc.CRITIC])
    grads = {k: g.numpy() for k, g in zip(orc.ACTOR + orc.CRITIC, ga + gc)}
    return float(a_loss.detach()), float(c_loss.detach()), a_prob.detach().numpy(), v.detach().numpy()[:, 0], grads


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_forward_losses_and_gradients_match_autograd(seed):
    rs = np.random.RandomState(seed)
    n_s, n_a, H, M, K = 90, 7, 12, 33, 9
    p = orc.init_params(n_s, n_a, hidden=H, seed=seed)
    for k in p:                                                   # biases away from zero, some units pushed past 6
        if k.endswith("_b"):
            p[k] = rs.normal(0.0, 0.5, size=p[k].shape)
    p["la"] *= 20.0
    idx = rs.randint(0, n_s, size=(M, K))
    s = orc.dense_from_idx(idx, n_s)
    assert s.sum() == M * K and s.max() >= 2                      # duplicates count twice
    a_his = rs.randint(0, n_a, size=M)
    v_target = rs.normal(size=M)
    a_prob, v, cache = orc.forward(p, s)
    assert (cache["h1a"] == 6.0).any() and (cache["h1a"] == 0.0).any(), "the test must exercise both relu6 kinks"
    a_loss, c_loss, g = orc.losses_and_grads(p, s, a_his, v_target)
    ra, rc, rp, rv, rg = _torch_losses(p, s, a_his, v_target, orc.ENTROPY_BETA)
    assert np.allclose(a_prob, rp, rtol=0, atol=1e-13) and np.allclose(v, rv, rtol=0, atol=1e-12)
    assert abs(a_loss - ra) < 1e-13 and abs(c_loss - rc) < 1e-12
    for k in orc.ACTOR + orc.CRITIC:
        assert g[k].shape == rg[k].shape, k
        assert np.allclose(g[k], rg[k], rtol=1e-10, atol=1e-13), (k, np.abs(g[k] - rg[k]).max())


def test_worker_targets_equal_the_batched_host_recursion():
    from drl_uav_cellularnet_b200.a3c import n_step_targets
    rs = np.random.RandomState(3)
    T, E = 10, 9
    r = rs.normal(size=(T, E))
    done = rs.rand(T, E) < 0.2
    done[:, 0] = False
    done[T - 1, 1] = True
    vb = rs.normal(size=E)
    got = n_step_targets(torch.from_numpy(r), torch.from_numpy(done), torch.from_numpy(vb)).numpy()
    for e in range(E):
        assert np.allclose(got[:, e], orc.worker_targets(r[:, e], done[:, e], vb[e]), rtol=0, atol=1e-12), e


def test_rmsprop_is_the_tf1_rule():
    rs = np.random.RandomState(4)
    p, ms = rs.normal(size=50), np.ones(50)
    p0 = p.copy()
    g1, g2 = rs.normal(size=50) * 1e-2, rs.normal(size=50) * 1e-2
    orc.rmsprop_step(p, g1, ms, lr=1e-4)
    m1 = 0.9 * 1.0 + 0.1 * g1 * g1
    assert np.allclose(ms, m1) and np.allclose(p, p0 - 1e-4 * g1 / np.sqrt(m1 + 1e-10))
    orc.rmsprop_step(p, g2, ms, lr=1e-4)
    m2 = 0.9 * m1 + 0.1 * g2 * g2
    assert np.allclose(ms, m2)
    assert np.allclose(p, p0 - 1e-4 * g1 / np.sqrt(m1 + 1e-10) - 1e-4 * g2 / np.sqrt(m2 + 1e-10))


def test_committed_vectors_are_the_oracles_answers():
    """tests/golden/acnet_oracle_vectors.npz (oracle/make_acnet_vectors.py): oracle-generated known answers for a small
    net, frozen so that kernel tests compare against numbers that do not move; regenerating must reproduce them."""
    import os
    from oracle import make_acnet_vectors as mk
    path = os.path.join(os.path.dirname(__file__), "golden", "acnet_oracle_vectors.npz")
    got, want = mk.build(), np.load(path)
    assert set(got) == set(want.files)
    for k in want.files:
        assert np.allclose(np.asarray(got[k], dtype=np.float64), want[k].astype(np.float64), rtol=1e-12, atol=1e-14), k
    p = {k[2:]: want[k] for k in want.files if k.startswith("p_")}
    s = orc.dense_from_idx(want["idx"], int(want["n_s"]))
    _, _, cache = orc.forward(p, s)
    assert (cache["h1a"] == 6.0).any() and (cache["h1a"] == 0.0).any() and (cache["h2a"] == 0.0).any()


def test_gradients_agree_with_central_differences():
    """independent of autograd: a few coordinates of every parameter by central differences of the oracle's own losses
    (away from the relu6 kinks the losses are smooth; step 1e-6, agreement to 1e-6 relative)"""
    rs = np.random.RandomState(7)
    n_s, n_a, H, M, K = 40, 5, 6, 17, 5
    p = orc.init_params(n_s, n_a, hidden=H, seed=7)
    for k in p:
        if k.endswith("_b"):
            p[k] = rs.normal(0.0, 0.3, size=p[k].shape)
    s = orc.dense_from_idx(rs.randint(0, n_s, size=(M, K)), n_s)
    a_his, v_target = rs.randint(0, n_a, size=M), rs.normal(size=M)
    _, _, g = orc.losses_and_grads(p, s, a_his, v_target)
    eps = 1e-6
    for name in orc.ACTOR + orc.CRITIC:
        which = 0 if name in orc.ACTOR else 1
        flat = p[name].reshape(-1)
        for i in rs.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[i]
            flat[i] = old + eps
            up = orc.losses_and_grads(p, s, a_his, v_target)[which]
            flat[i] = old - eps
            dn = orc.losses_and_grads(p, s, a_his, v_target)[which]
            flat[i] = old
            fd = (up - dn) / (2 * eps)
            an = g[name].reshape(-1)[i]
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 1e-9, (name, i, fd, an)
