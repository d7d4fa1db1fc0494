#!/usr/bin/env python
"""uavnet_gemm against a float64 product of the same operands, case by case (prints, does not assert):
    python profiles/gemm_check.py
(Round 1 used this with UAVNET_GEMM_DBG = 0..15 to swap the LBO / SBO fields of the MN-major descriptors: every variant
returned zeros, see profiles/r1/NOTES.md section 8; the knob no longer exists.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_trunc(x):
    import torch
    return (x.view(torch.int32) & -8192).view(torch.float32)


def run_cases():
    import torch
    from drl_uav_cellularnet_b200 import dense
    torch.manual_seed(0)
    dev = "cuda"
    out = []
    cases = [  # name, M, N, K, a_trans, b_trans
        ("NT 128x16x32", 128, 16, 32, 0, 1), ("NN 128x16x32", 128, 16, 32, 0, 0), ("TN 128x16x32", 128, 16, 32, 1, 0),
        ("TT 128x16x32", 128, 16, 32, 1, 1),
        ("NT 128x64x64", 128, 64, 64, 0, 1), ("NN 128x64x64", 128, 64, 64, 0, 0), ("TN 128x64x64", 128, 64, 64, 1, 0),
        ("NN 300x200x200", 300, 200, 200, 0, 0), ("NN 300x625x200", 300, 625, 200, 0, 0),
        ("NT 300x200x625", 300, 200, 625, 0, 1), ("TN 200x625x5000", 200, 625, 5000, 1, 0),
        ("NN 8192x200x200", 8192, 200, 200, 0, 0),
    ]
    for prec in ("tf32", "fp32"):
        for name, M, N, K, at, bt in cases:
            A = torch.randn((K, M) if at else (M, K), device=dev)
            B = torch.randn((N, K) if bt else (K, N), device=dev)
            try:
                D = dense.gemm(A, B, a_trans=bool(at), b_trans=bool(bt), precision=prec)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                out.append({"case": name, "prec": prec, "error": repr(e)})
                break
            Ad, Bd = (A.t() if at else A).double(), (B.t() if bt else B).double()
            ref = Ad @ Bd
            At, Bt = tf32_trunc(A), tf32_trunc(B)
            ref_t = (At.t() if at else At).double() @ (Bt.t() if bt else Bt).double()
            scale = float(ref.abs().max())
            out.append({"case": name, "prec": prec, "err_vs_f64": float((D.double() - ref).abs().max()) / scale,
                        "err_vs_truncated_inputs": float((D.double() - ref_t).abs().max()) / scale,
                        "nan": bool(torch.isnan(D).any())})
    out.append({"gemm_check_flag": dense.check()})
    return out


if __name__ == "__main__":
    if "--child" in sys.argv:
        print(json.dumps(run_cases()))
    else:
        sweep = [0]
        for dbg in sweep:
            env = dict(os.environ, UAVNET_GEMM_DBG=str(dbg))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=600)
            print("=== UAVNET_GEMM_DBG=%d rc=%d" % (dbg, r.returncode))
            good = False
            try:
                rows = json.loads(r.stdout.strip().splitlines()[-1])
                for row in rows:
                    print("   ", row)
                good = all(row.get("err_vs_f64", 0.0) < 1e-2 and "error" not in row for row in rows)
            except Exception:  # noqa: BLE001
                print(r.stdout[-2000:], r.stderr[-3000:])
            if good and dbg == 0 and "--dbg-sweep" not in sys.argv:
                break                       # the production descriptors are right: no sweep needed
