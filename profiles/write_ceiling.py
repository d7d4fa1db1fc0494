#!/usr/bin/env python
"""Write-only HBM ceiling of the box, measured with the store mechanisms the step kernel can use
(uavenv_diag_fill: st.global.v4 / cp.async.bulk from a zero tile) next to cudaMemset and torch.zero_.
Buffer = one config[1] observation batch (4096 x 5 x 100 x 100 float32 = 819.2 MB > L2).  CUDA-event timed."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drl_uav_cellularnet_b200 import _native as N  # noqa: E402

L = N.diag_lib()          # libuavenv_diag.so (include/uavenv_diag.h)
dev = torch.device("cuda", 0)
nbytes = 4096 * 5 * 100 * 100 * 4
buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=50, warm=5):
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


out = {"bytes": nbytes}
out["torch_zero_GBs"] = nbytes / timeit(lambda: buf.zero_()) / 1e9
for mode, name in ((0, "stg_v4"), (1, "bulk_tma")):
    for per_cta in (200000, 400000, 1600000, 16384 * 32):
        def f():
            rc = L.uavenv_diag_fill(C.c_void_p(buf.data_ptr()), nbytes, per_cta, mode, st)
            assert rc == 0, rc
        out["%s_per_cta_%d_GBs" % (name, per_cta)] = nbytes / timeit(f) / 1e9
a = torch.empty(nbytes // 2, dtype=torch.uint8, device=dev)
b = torch.empty(nbytes // 2, dtype=torch.uint8, device=dev)
out["torch_copy_rw_GBs"] = nbytes / timeit(lambda: b.copy_(a)) / 1e9
print(json.dumps(out))
