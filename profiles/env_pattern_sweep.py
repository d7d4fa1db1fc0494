#!/usr/bin/env python
"""Sweep of the isolated "zeros by TMA, counts by RED one chunk later" pattern (uavenv_diag_fill_env)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drl_uav_cellularnet_b200 import _native as N  # noqa: E402

L = N.diag_lib()          # libuavenv_diag.so (include/uavenv_diag.h)
nbytes = 4096 * 5 * 100 * 100 * 4
buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=30, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


for grid in (148, 296, 444, 592, 1184, 4096):
    for tile in (16384, 32768, 65536):
        for flags, n_red in ((0, 0), (1, 44), (2, 0), (3, 44)):
            def f():
                rc = L.uavenv_diag_fill_env(C.c_void_p(buf.data_ptr()), nbytes, 200000, grid, tile, flags, n_red, st)
                assert rc == 0, rc
            gbs = nbytes / timeit(f) / 1e9
            print({"grid": grid, "tile": tile, "flags": flags, "n_red": n_red, "GBs": round(gbs, 1)}, flush=True)
