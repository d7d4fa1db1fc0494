#!/usr/bin/env python
"""config[2]-style measurement of the learner on top of the env step path: synchronous A3C iterations (rollout of 10
steps + one update) over E envs per GPU, device-timed; prints one JSON line with env-steps/s and the time split.
    python profiles/bench_a3c.py [--envs 8192] [--iters 20]         (torchrun for N > 1: NCCL gradient all-reduce)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drl_uav_cellularnet_b200 import BatchedMobiEnvironment  # noqa: E402
from drl_uav_cellularnet_b200 import dist as udist  # noqa: E402
from drl_uav_cellularnet_b200.a3c import A3CTrainer, ACNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=8192)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--tf32", action="store_true", help="dense layers with one tcgen05 TF32 MMA per k-step (default: 3xTF32, fp32-class accuracy)")
ap.add_argument("--p2p", action="store_true", help="gradient push as one peer-memory kernel (uavnet_p2p_rmsprop) instead of NCCL all-reduce + RMSProp")
ap.add_argument("--groups", type=int, default=1, help="env handles per GPU, each rolling out on its own stream")
ap.add_argument("--graph", action="store_true", help="also time the iteration replayed as one CUDA graph")
args = ap.parse_args()
rank, world, local = udist.world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
udist.init("nccl", dev)
assert args.envs % args.groups == 0
eg = args.envs // args.groups
envs = [BatchedMobiEnvironment(eg, 4, 40, 100, "group", seed=2026, obs="none", env_offset=rank * args.envs + g * eg, device=local)
        for g in range(args.groups)]
env = envs[0]
net = ACNet(env.observation_space_dim, env.action_space_dim, dev, precision="tf32" if args.tf32 else "fp32")
if args.p2p:
    net.enable_p2p()
tr = A3CTrainer(envs if args.groups > 1 else env, net, seed=100 + rank)


def timed(fn, n):
    udist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    udist.barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(args.warmup):
    tr.train_iteration()
ms_iter = timed(tr.train_iteration, args.iters)
ms_roll = timed(tr.rollout, args.iters)
vt = tr.rollout()
ms_upd = timed(lambda: tr.update(vt), args.iters)
ms_env = timed(lambda: env.step(tr.buf_a[0][:env.n_envs]), 50)
ms_graph = float("nan")
if args.graph:
    tr.capture()
    for _ in range(2):
        tr.train_iteration_graph()
    ms_graph = timed(tr.train_iteration_graph, args.iters)
ms_iter, ms_roll, ms_upd, ms_env, ms_graph = udist.max_over_ranks([ms_iter, ms_roll, ms_upd, ms_env, ms_graph], dev)
if rank == 0:
    steps = args.envs * world * tr.T
    print(json.dumps({"metric": "A3C env-steps/sec (rollout + update)", "value": steps / (ms_iter * 1e-3), "n_gpus": world,
                      "envs_per_gpu": args.envs, "rollout_steps": tr.T, "ms_per_iteration": ms_iter, "ms_rollout": ms_roll,
                      "ms_update": ms_upd, "ms_env_step_no_obs": ms_env, "ms_per_iteration_graph": ms_graph,
                      "value_graph": steps / (ms_graph * 1e-3) if args.graph else None, "tf32": bool(args.tf32), "groups": args.groups, "p2p_push": bool(args.p2p), "params": net.n_params,
                      "allreduce_bytes": net.n_flat * 4 if world > 1 else 0}))
if world > 1:
    # no process-group teardown: with NCCL collectives captured in the iteration's CUDA graph (and peer mappings open)
    # destroy_process_group() did not return on 2 B200s; the numbers are out, leave without it
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)
