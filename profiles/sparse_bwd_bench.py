#!/usr/bin/env python
"""First-layer weight gradient of a rollout batch (81 920 samples x 44 indices into 50 000 rows x 400 columns): the scatter
version (float REDs) against the gather version (counting sort + per-row sums) for several column-pass counts.
    python profiles/sparse_bwd_bench.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drl_uav_cellularnet_b200 import BatchedMobiEnvironment  # noqa: E402
from drl_uav_cellularnet_b200 import _native as N  # noqa: E402

L = N.lib()
M, K, R, H = 81920, 44, 50000, 400
# indices of a real rollout: 8192 envs x 10 steps
env = BatchedMobiEnvironment(8192, 4, 40, 100, "group", seed=3, obs="none")
env.reset()
rows = []
g = torch.Generator(device="cuda").manual_seed(1)
for t in range(10):
    env.step(torch.randint(0, 625, (8192,), device="cuda", generator=g))
    rows.append(env.obs_idx.clone())
idx = torch.cat(rows).contiguous()
dpre = torch.randn(M, H, device="cuda")
dpre[torch.rand(M, H, device="cuda") < 0.5] = 0.0
dW = torch.zeros(R, H, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ws = torch.empty(L.uavnet_sparse_bwd_gather_workspace(M, K, R) // 4, dtype=torch.int32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


out = {"distinct_rows_touched": int(torch.unique(idx).numel()), "pairs": M * K,
       "scatter_us": timed(lambda: L.uavnet_sparse_bwd(p(idx), M, K, R, p(dpre), H, p(dW), st))}
for passes in (1, 2, 4, 5, 10, 25):
    out["gather_us_passes_%d" % passes] = timed(lambda: L.uavnet_sparse_bwd_gather(p(idx), M, K, R, p(dpre), H, p(dW), p(ws), passes, st))
print(json.dumps(out))
