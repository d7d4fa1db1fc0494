#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Known-answer vectors of the learner's oracle (oracle/acnet_oracle.py) on a small net:
    python oracle/make_acnet_vectors.py      ->  tests/golden/acnet_oracle_vectors.npz
These are ORACLE-generated, not recorded from the reference: TensorFlow is absent from the build image and unpinned by
the reference (README.md:7), so rows f1/f2 have no reference-owned vectors.  The oracle itself is pinned to
torch.autograd on the CPU (tests/test_acnet_oracle.py); the vectors freeze its answers so that a GPU test can compare the
kernels (ACNet with n_s = 600, n_a = 25, hidden = 40) against numbers that do not move."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import acnet_oracle as orc  # noqa: E402

N_S, N_A, HIDDEN, M, K, SEED = 600, 25, 40, 96, 11, 20261019


def build():
    rs = np.random.RandomState(SEED)
    p = orc.init_params(N_S, N_A, hidden=HIDDEN, seed=SEED)
    for k in p:
        if k.endswith("_b"):
            p[k] = rs.normal(0.0, 0.3, size=p[k].shape)
    p["la"] *= 6.0                                    # some first-layer units saturate at 6
    idx = rs.randint(0, N_S, size=(M, K)).astype(np.int32)
    a_his = rs.randint(0, N_A, size=M).astype(np.int64)
    v_target = rs.normal(size=M)
    s = orc.dense_from_idx(idx, N_S)
    a_prob, v, _ = orc.forward(p, s)
    a_loss, c_loss, g = orc.losses_and_grads(p, s, a_his, v_target)
    out = dict(n_s=N_S, n_a=N_A, hidden=HIDDEN, idx=idx, a_his=a_his, v_target=v_target, a_prob=a_prob, v=v,
               a_loss=a_loss, c_loss=c_loss)
    out.update({"p_" + k: x for k, x in p.items()})
    out.update({"g_" + k: x for k, x in g.items()})
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "acnet_oracle_vectors.npz")
    np.savez_compressed(path, **build())
    print(path, os.path.getsize(path), "bytes")
