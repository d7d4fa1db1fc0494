#!/usr/bin/env python
"""Structured-operand probes of uavnet_gemm (development aid): prints what the tensor core read."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drl_uav_cellularnet_b200 import dense  # noqa: E402

torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
dev = "cuda"


def probe(name, at, bt, M=128, N=16, K=32, prec="tf32"):
    Al = torch.zeros((M, K), device=dev)
    for i in range(min(M, K)):
        Al[i, i] = 1.0
    Bl = (torch.arange(K, device=dev)[:, None] * 100.0 + torch.arange(N, device=dev)[None, :] + 1.0)
    A = Al.t().contiguous() if at else Al
    B = Bl.t().contiguous() if bt else Bl
    D = dense.gemm(A, B, a_trans=bool(at), b_trans=bool(bt), precision=prec)
    torch.cuda.synchronize()
    ref = Al @ Bl
    print("=== %s (a_trans=%d b_trans=%d) M=%d N=%d K=%d %s: max|D|=%.1f nonzero=%d err=%.3g" %
          (name, at, bt, M, N, K, prec, float(D.abs().max()), int((D != 0).sum()), float((D - ref).abs().max())))
    print(D[:10, :N if N < 17 else 16])


probe("NT", 0, 1)
probe("NN", 0, 0)
probe("TN", 1, 0)
probe("TT", 1, 1)
for (M, N, K) in [(128, 64, 96), (128, 64, 128), (128, 64, 640), (128, 112, 640), (128, 200, 640), (128, 200, 625), (300, 200, 625), (128, 128, 640), (128, 96, 640)]:
    for prec in ("tf32", "fp32"):
        A = torch.randn(M, K, device=dev)
        B = torch.randn(N, K, device=dev)
        D = dense.gemm(A, B, b_trans=True, precision=prec)
        ref = A.double() @ B.double().t()
        e = (D.double() - ref).abs()
        bad_cols = (e.max(0).values > 0.05).nonzero().flatten().tolist()
        bad_rows = (e.max(1).values > 0.05).nonzero().flatten().tolist()
        print("NT %dx%dx%d %s err=%.3g bad cols %s%s bad rows %s%s" % (M, N, K, prec, float(e.max() / ref.abs().max()), bad_cols[:6], "..." if len(bad_cols) > 6 else "",
                                                           bad_rows[:6], "..." if len(bad_rows) > 6 else ""), len(bad_cols), len(bad_rows))
print("flag", dense.check())
