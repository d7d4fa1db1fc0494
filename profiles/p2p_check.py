#!/usr/bin/env python
"""Two or more ranks (torchrun, one process per GPU): the gradient push as ONE peer-memory kernel per rank
(uavnet_p2p_rmsprop: reduce-scatter by NVLink peer loads + RMSProp + all-gather by peer stores) against the baseline
NCCL all-reduce + uavnet_rmsprop -- same parameters bit for bit at 2 ranks, and the device time of both."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from drl_uav_cellularnet_b200 import dist as udist  # noqa: E402
from drl_uav_cellularnet_b200.a3c import ACNet  # noqa: E402

rank, world, local = udist.world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
udist.init("nccl", dev)
a, b = ACNet(50000, 625, dev), ACNet(50000, 625, dev)
b.enable_p2p()
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(4):
    grad = torch.randn(a.n_flat, device=dev, generator=g) * 1e-2       # a different gradient on every rank
    a.grad.copy_(grad)
    b.grad.copy_(grad)
    if world > 1:
        dist.all_reduce(a.grad)
    a.apply_grads(1e-4, world)
    b.apply_grads(1e-4)
    torch.cuda.synchronize()
    worst = max(worst, float((a.flat - b.flat).abs().max()))
    assert float(b.grad.abs().max()) == 0.0
same = [None] * world
dist.all_gather_object(same, float(b.flat.double().sum()))             # every rank holds the same parameters


def timed(fn, n=30):
    for _ in range(3):
        fn()
    udist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def nccl_push():
    if world > 1:
        dist.all_reduce(a.grad)
    a.apply_grads(1e-4, world)


ms_nccl = timed(nccl_push)
ms_p2p = timed(lambda: b.apply_grads(1e-4))
ms_nccl, ms_p2p = udist.max_over_ranks([ms_nccl, ms_p2p], dev)
if rank == 0:
    print(json.dumps({"world": world, "max_abs_param_diff_vs_nccl_path": worst, "param_sums_per_rank": same,
                      "ms_nccl_allreduce_plus_rmsprop": ms_nccl, "ms_p2p_fused": ms_p2p, "bytes": a.n_flat * 4}))
b.close_p2p()
dist.destroy_process_group()
