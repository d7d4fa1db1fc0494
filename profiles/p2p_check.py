#!/usr/bin/env python
"""Two or more ranks (torchrun, one process per GPU): the gradient push as ONE peer-memory kernel per rank
(uavnet_p2p_push: reduce-scatter by NVLink peer loads + RMSProp + all-gather by peer stores, ranks ordered by flag
words in peer memory) against the baseline
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

import argparse  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--same-gpu", action="store_true", help="all ranks on cuda:0 (two processes sharing one GPU): the plumbing goes "
                "through gloo, the push through CUDA IPC mappings of the other process's buffers on the same device")
ap.add_argument("--split", action="store_true", help="push in two parts (ACNet.enable_p2p(split=True)): the actor half of the "
                "first layer early on a side stream, the rest in apply_grads")
ap.add_argument("--n-s", type=int, default=50000)
ap.add_argument("--n-a", type=int, default=625)
ap.add_argument("--hidden", type=int, default=200)
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
rank, world, local = udist.world()
if args.same_gpu:
    local = 0
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
udist.init("gloo" if args.same_gpu else "nccl", dev)
a, b = ACNet(args.n_s, args.n_a, dev, hidden=args.hidden), ACNet(args.n_s, args.n_a, dev, hidden=args.hidden)
b.enable_p2p(split=args.split)
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(4):
    grad = torch.randn(a.n_flat, device=dev, generator=g) * 1e-2       # a different gradient on every rank
    a.grad.copy_(grad)
    b.grad.copy_(grad)
    if world > 1:
        if args.same_gpu:                      # gloo: reduce on the host
            t = a.grad.cpu()
            dist.all_reduce(t)
            a.grad.copy_(t)
        else:
            dist.all_reduce(a.grad)
    a.apply_grads(1e-4, world)
    if args.split and it % 2 == 0:
        b.push_early(1e-4)                     # every other iteration the first part goes ahead on its side stream
    b.apply_grads(1e-4)
    torch.cuda.synchronize()
    worst = max(worst, float((a.flat - b.flat).abs().max()))
    assert float(b.grad.abs().max()) == 0.0
same = [None] * world
dist.all_gather_object(same, float(b.flat.double().sum()))             # every rank holds the same parameters


def timed(fn, n=args.iters):
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
    if world > 1 and not args.same_gpu:
        dist.all_reduce(a.grad)
    a.apply_grads(1e-4, world)


ms_nccl = timed(nccl_push)
ms_p2p = timed(lambda: b.apply_grads(1e-4))
ms_nccl, ms_p2p = udist.max_over_ranks([ms_nccl, ms_p2p], dev)
pushes, gave_up = b.p2p_status()
# same state on both paths before the optimiser slots are compared: one more identical gradient through both
grad = torch.randn(a.n_flat, device=dev, generator=g) * 1e-2
a.flat.copy_(b.flat)
b.close_p2p()                                          # gathers the RMSProp slot slices of all ranks
ms_sums = [None] * world
dist.all_gather_object(ms_sums, float(b.ms.double().sum()))
ms_diff = float((a.ms - b.ms).abs().max())     # same gradient history on both paths: the gathered slots must agree
if rank == 0:
    print(json.dumps({"world": world, "max_abs_param_diff_vs_nccl_path": worst, "param_sums_per_rank": same,
                      "ms_nccl_allreduce_plus_rmsprop": ms_nccl, "ms_p2p_fused": ms_p2p, "bytes": a.n_flat * 4,
                      "pushes": pushes, "flag_wait_gave_up": gave_up, "ms_slot_sums_after_close_per_rank": ms_sums,
                      "ms_slot_max_abs_diff_vs_nccl_path": ms_diff, "split": bool(args.split)}))
dist.barrier()
dist.destroy_process_group()
