#!/usr/bin/env python
"""Device time of uavnet_gemm on the learner's shapes (CUDA events, L2 flushed between launches is NOT done: the
operands of one launch exceed L2 for the update shapes).  python profiles/gemm_bench.py [--prec tf32|fp32] [--only i]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drl_uav_cellularnet_b200 import dense  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--prec", default="tf32")
ap.add_argument("--only", type=int, default=-1)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = "cuda"
M, E, H, A = 81920, 8192, 200, 625
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(s, device=dev, generator=g)  # noqa: E731
h1, h1s = r(M, 2 * H), r(E, 2 * H)
W2, W3 = r(H, H), torch.zeros((H, 628), device=dev)[:, :A]
b2, b3 = r(H), r(A)
h2, h2s, dzf = r(M, H), r(E, H), torch.zeros((M, 628), device=dev)
dz = dzf[:, :A]
dz.copy_(r(M, A))
gW2, gW3, gb2, gb3 = torch.zeros((H, H), device=dev), torch.zeros((H, 628), device=dev)[:, :A], torch.zeros(H, device=dev), torch.zeros(A, device=dev)
o200, o200s, o625s, o400 = torch.empty((M, H), device=dev), torch.empty((E, H), device=dev), torch.empty((E, A), device=dev), torch.empty((M, 2 * H), device=dev)
cases = [
    ("fwd 8192x200x200 bias relu6", lambda: dense.gemm(h1s[:, :H], W2, o200s, bias=b2, relu6=True, precision=args.prec), 2.0 * E * H * H),
    ("fwd 8192x625x200 bias", lambda: dense.gemm(h2s, W3, o625s, bias=b3, precision=args.prec), 2.0 * E * A * H),
    ("fwd 81920x200x200 bias relu6", lambda: dense.gemm(h1[:, H:], W2, o200, bias=b2, relu6=True, precision=args.prec), 2.0 * M * H * H),
    ("dgrad 81920x200x625 mask", lambda: dense.gemm(dz, W3, o200, b_trans=True, mask_src=h2, precision=args.prec), 2.0 * M * A * H),
    ("dgrad 81920x200x200 mask", lambda: dense.gemm(h2, W2, o400[:, :H], b_trans=True, mask_src=h1[:, :H], precision=args.prec), 2.0 * M * H * H),
    ("wgrad 200x625x81920 +colsum", lambda: dense.gemm(h2, dz, gW3, a_trans=True, accumulate=True, colsum=gb3, precision=args.prec), 2.0 * M * A * H),
    ("wgrad 200x200x81920 +colsum", lambda: dense.gemm(h1[:, :H], h2, gW2, a_trans=True, accumulate=True, colsum=gb2, precision=args.prec), 2.0 * M * H * H),
    ("fwd 8192x200x200 W^T bias relu6 (all TMA)", lambda: dense.gemm(h1s[:, :H], W2, o200s, b_trans=True, bias=b2, relu6=True, precision=args.prec), 2.0 * E * H * H),
    ("dgrad 81920x200x200 mask + out_colsum", lambda: dense.gemm(h2, W2, o400[:, :H], b_trans=True, mask_src=h1[:, :H], out_colsum=gb2, precision=args.prec), 2.0 * M * H * H),
    ("colsum 400 of 81920", lambda: dense.gemm(None, h1, colsum=torch.zeros(2 * H, device=dev), accumulate=True, precision=args.prec), 2.0 * M * 2 * H),
]
out = []
for i, (name, fn, flops) in enumerate(cases):
    if args.only >= 0 and i != args.only:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.reps
    out.append({"case": name, "us": round(us, 1), "tflops": round(flops / us / 1e6, 1)})
    print(out[-1], flush=True)
print(json.dumps({"prec": args.prec, "flag": dense.check(), "cases": out}))
