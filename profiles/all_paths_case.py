#!/usr/bin/env python
"""Small end-to-end case that runs every kernel path of libuavenv once (for the -DUAVENV_BOUNDS_CHECK build, and for
compute-sanitizer where that tool is open) --
constructor, reset, steps with the TMA zero stream, the plain-store fallback (odd grid), incremental observations, the
4-BSs-per-lane channel pass, trace replay with injected fading (fp64), masked reset, host-buffer step, sparse MLP layer
forward / backward and the RMSProp step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drl_uav_cellularnet_b200 import BatchedMobiEnvironment  # noqa: E402
from drl_uav_cellularnet_b200.a3c import A3CTrainer, ACNet  # noqa: E402

rs = np.random.RandomState(0)
for kw in (dict(nBS=4, nUE=40, G=100, obs="f32"), dict(nBS=4, nUE=40, G=25, obs="f32"),
           dict(nBS=4, nUE=40, G=100, obs="f32_incremental"), dict(nBS=12, nUE=100, G=64, obs="f32"),
           dict(nBS=4, nUE=40, G=100, obs="f32", precision="fp64")):
    nBS, nUE, G = kw.pop("nBS"), kw.pop("nUE"), kw.pop("G")
    gs = [nUE // 4] * 4
    env = BatchedMobiEnvironment(6, nBS, nUE, G, "group", seed=1, group_sizes=gs,
                                 init_bs_xy=[[2 + (G - 4) * b // nBS, 2 + (G - 4) * ((b * 5) % nBS) // nBS] for b in range(nBS)], **kw)
    env.reset()
    for t in range(4):
        env.step(rs.randint(0, 5, size=(6, nBS)).astype(np.uint8))
    env.reset(env_mask=np.array([1, 0, 1, 0, 0, 1], dtype=np.uint8))
    env.step(rs.randint(0, 5, size=(6, nBS)).astype(np.uint8))
    assert env.check() == 0
    torch.cuda.synchronize()
trace = rs.randint(0, 100, size=(8, 40, 2))
env = BatchedMobiEnvironment(3, 4, 40, 100, "read_trace", trace=trace, fading="injected", precision="fp64")
env.ctor_pass(fading=rs.normal(0, 2, size=(3, 40, 4)))
env.reset(fading=rs.normal(0, 2, size=(3, 40, 4)))
for t in range(5):
    env.step(rs.randint(0, 625, size=3), fading=rs.normal(0, 2, size=(3, 40, 4)))
assert env.check() == 0
env = BatchedMobiEnvironment(16, 4, 40, 100, "group", seed=2)
env.reset()
act, rew, done = torch.zeros(16, dtype=torch.int64).pin_memory(), torch.zeros(16, dtype=torch.float64).pin_memory(), \
    torch.zeros(16, dtype=torch.uint8).pin_memory()
for t in range(3):
    env.step_host(act, rew, done)
env2 = BatchedMobiEnvironment(16, 4, 40, 100, "group", seed=3, obs="none", max_step=12)
net = ACNet(env2.observation_space_dim, env2.action_space_dim, env2.device)
tr = A3CTrainer(env2, net, seed=4)
for it in range(2):
    tr.train_iteration()
torch.cuda.synchronize()
print("all-paths case ok")
