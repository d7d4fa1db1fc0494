"""TEST INFRASTRUCTURE ONLY -- a float64 numpy restatement of the reference's actor-critic graph and of its update
(rows f1/f2 of SURVEY.md section 8).  Only tests/ may import it; the product path never does.

What it restates (file:line into the reference repo):
  main.py:143-156   _build_net, netType 'MLP': actor s -> relu6(200) -> relu6(200) -> softmax(N_A),
                    critic s -> relu6(200) -> relu6(200) -> 1, tf.layers.dense = x @ kernel + bias
  main.py:64-78     td = v_target - v;  c_loss = mean(td^2)
                    log_prob = sum(log(a_prob + 1e-5) * one_hot(a_his));  exp_v = log_prob * stop_gradient(td)
                    entropy = -sum(a_prob * log(a_prob + 1e-5));  a_loss = mean(-(ENTROPY_BETA * entropy + exp_v))
  main.py:79-80     a_grads = d a_loss / d a_params, c_grads = d c_loss / d c_params (written out by hand below)
  main.py:217-227   v_s_ = 0 if done else v(s'); walking the buffer backwards v_s_ = r + GAMMA * v_s_
  main.py:300-301   tf.train.RMSPropOptimizer(lr): ms = decay ms + (1 - decay) g^2; var -= lr g / sqrt(ms + eps),
                    TF 1.x defaults decay 0.9, momentum 0, epsilon 1e-10, the ms slot starts at ones

Parity status: TensorFlow is absent from the build image and its version is not pinned by the reference
(README.md:7), so this restatement is **unpinned by reference-owned artefacts**; tests/test_acnet_oracle.py pins the
hand-written gradients to torch.autograd (float64, CPU) of the same graph.
"""
from __future__ import annotations

import numpy as np

ENTROPY_BETA = 0.001
GAMMA = 0.9
RMS_DECAY, RMS_EPS = 0.9, 1e-10

ACTOR = ("la", "la_b", "la2", "la2_b", "ap", "ap_b")
CRITIC = ("lc", "lc_b", "lc2", "lc2_b", "v", "v_b")


def relu6(x):
    return np.clip(x, 0.0, 6.0)


def relu6_grad(y):
    """derivative of relu6 expressed through its output (0 at both kinks, like tf.nn.relu6's gradient)"""
    return ((y > 0.0) & (y < 6.0)).astype(np.float64)


def init_params(n_s, n_a, hidden=200, seed=0):
    """tf.random_normal_initializer(0., .1) kernels, zero biases (main.py:145-153)"""
    rs = np.random.RandomState(seed)
    shapes = {"la": (n_s, hidden), "la2": (hidden, hidden), "ap": (hidden, n_a),
              "lc": (n_s, hidden), "lc2": (hidden, hidden), "v": (hidden, 1)}
    p = {k: rs.normal(0.0, 0.1, size=s) for k, s in shapes.items()}
    for k, s in shapes.items():
        p[k + "_b"] = np.zeros(s[1])
    return p


def dense_from_idx(idx, n_s):
    """count vector of the observation from its non-zero cells (duplicates count twice): [M, K] int -> [M, n_s]"""
    s = np.zeros((idx.shape[0], n_s))
    for m in range(idx.shape[0]):
        np.add.at(s[m], idx[m], 1.0)
    return s


def forward(p, s):
    """main.py:143-156 -> (a_prob [M, N_A], v [M], cache)"""
    h1a = relu6(s @ p["la"] + p["la_b"])
    h2a = relu6(h1a @ p["la2"] + p["la2_b"])
    z = h2a @ p["ap"] + p["ap_b"]
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    a_prob = e / e.sum(axis=1, keepdims=True)
    h1c = relu6(s @ p["lc"] + p["lc_b"])
    h2c = relu6(h1c @ p["lc2"] + p["lc2_b"])
    v = (h2c @ p["v"] + p["v_b"])[:, 0]
    return a_prob, v, dict(s=s, h1a=h1a, h2a=h2a, h1c=h1c, h2c=h2c, a_prob=a_prob, v=v)


def losses_and_grads(p, s, a_his, v_target, beta=ENTROPY_BETA):
    """main.py:64-80 -> (a_loss, c_loss, grads dict keyed like p)"""
    a_prob, v, c = forward(p, s)
    M, n_a = a_prob.shape
    td = v_target - v
    c_loss = np.mean(td * td)
    lp = np.log(a_prob + 1e-5)
    log_prob = lp[np.arange(M), a_his]
    entropy = -(a_prob * lp).sum(axis=1)
    a_loss = np.mean(-(beta * entropy + log_prob * td))               # td enters as a constant (tf.stop_gradient)
    g = {}
    # ---- critic ----
    dv = (-2.0 / M) * td                                               # d c_loss / d v
    g["v"] = c["h2c"].T @ dv[:, None]
    g["v_b"] = np.array([dv.sum()])
    d2c = (dv[:, None] * p["v"].T) * relu6_grad(c["h2c"])
    g["lc2"] = c["h1c"].T @ d2c
    g["lc2_b"] = d2c.sum(axis=0)
    d1c = (d2c @ p["lc2"].T) * relu6_grad(c["h1c"])
    g["lc"] = s.T @ d1c
    g["lc_b"] = d1c.sum(axis=0)
    # ---- actor: d a_loss / d prob, then through the softmax ----
    gp = (beta / M) * (lp + a_prob / (a_prob + 1e-5))
    gp[np.arange(M), a_his] -= td / (M * (a_prob[np.arange(M), a_his] + 1e-5))
    dz = a_prob * (gp - (a_prob * gp).sum(axis=1, keepdims=True))
    g["ap"] = c["h2a"].T @ dz
    g["ap_b"] = dz.sum(axis=0)
    d2a = (dz @ p["ap"].T) * relu6_grad(c["h2a"])
    g["la2"] = c["h1a"].T @ d2a
    g["la2_b"] = d2a.sum(axis=0)
    d1a = (d2a @ p["la2"].T) * relu6_grad(c["h1a"])
    g["la"] = s.T @ d1a
    g["la_b"] = d1a.sum(axis=0)
    return a_loss, c_loss, g


def rmsprop_step(param, grad, ms, lr=1e-4, decay=RMS_DECAY, eps=RMS_EPS):
    """one tf.train.RMSPropOptimizer (TF 1.x, momentum 0) step in place -> (param, ms)"""
    ms *= decay
    ms += (1.0 - decay) * grad * grad
    param -= lr * grad / np.sqrt(ms + eps)
    return param, ms


def worker_targets(rewards, dones, v_boot, gamma=GAMMA):
    """buffer_v_target of Worker.work (main.py:212-238) for one env: the buffer is flushed at `done` with v_s_ = 0 and at
    the end of the rollout with v_s_ = v(s'); rewards/dones [T] -> targets [T]"""
    T = len(rewards)
    out = np.zeros(T)
    t0 = 0
    for t in range(T):
        if dones[t] or t == T - 1:
            v_s_ = 0.0 if dones[t] else float(v_boot)
            for i in range(t, t0 - 1, -1):
                v_s_ = rewards[i] + gamma * v_s_
                out[i] = v_s_
            t0 = t + 1
    return out
