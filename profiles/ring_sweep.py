#!/usr/bin/env python
"""Sweep of the isolated store-warp pattern (uavenv_diag_fill_ring): which (grid, ring, tile) shapes of a persistent
TMA tile ring reach the write-only ceiling.  Buffer = one config[1] observation batch (819.2 MB)."""
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


rows = []
for grid in (148, 296, 444):
    for ring, tile in ((2, 16384), (3, 16384), (4, 16384), (2, 32768), (3, 32768)):
        for flags in (1, 5, 13):
            if (ring * tile + 2048) * (grid // 148) > 225 * 1024 or ring * tile > 200 * 1024:
                continue
            def f():
                rc = L.uavenv_diag_fill_ring(C.c_void_p(buf.data_ptr()), nbytes, 200000, grid, ring, tile, flags, st)
                assert rc == 0, rc
            gbs = nbytes / timeit(f) / 1e9
            rows.append({"grid": grid, "ring": ring, "tile": tile, "flags": flags, "GBs": round(gbs, 1)})
            print(rows[-1], flush=True)
def g():
    assert L.uavenv_diag_fill(C.c_void_p(buf.data_ptr()), nbytes, 200000, 1, st) == 0
print(json.dumps({"bulk_all_at_once_GBs": nbytes / timeit(g) / 1e9, "sweep": rows}))
