"""Builds libuavenv.so (the C-ABI library of include/uavenv.h) in-tree with nvcc for sm_100a.

    python -m drl_uav_cellularnet_b200.build [--force]

The .so is git-ignored (history stays source-only) but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libuavenv.so")
SOURCES = ["uavenv.cu", "uavnet.cu"]
HEADERS = ["env_kernels.cuh", "philox.cuh", "tc_gemm.cuh", os.path.join(ROOT, "include", "uavenv.h"), os.path.join(ROOT, "include", "uavnet.h")]
# diagnostics of the observation stream (include/uavenv_diag.h): a library of its own, not part of the product C-ABI
DIAG_SO = os.path.join(PKG, "libuavenv_diag.so")
DIAG_SOURCES = ["uavenv_diag.cu"]
DIAG_HEADERS = ["env_kernels.cuh", "philox.cuh", os.path.join(ROOT, "include", "uavenv_diag.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libuavenv.so")


def up_to_date(so: str = SO, sources=SOURCES, headers=HEADERS) -> bool:
    if not os.path.isfile(so):
        return False
    t = os.path.getmtime(so)
    deps = [os.path.join(CSRC, s) for s in sources] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in headers]
    return all(os.path.getmtime(d) <= t for d in deps)


def build_diag(force: bool = False) -> str:
    """libuavenv_diag.so (include/uavenv_diag.h): the zero-fill kernels behind profiles/write_ceiling.py etc."""
    if not force and up_to_date(DIAG_SO, DIAG_SOURCES, DIAG_HEADERS):
        return DIAG_SO
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", DIAG_SO] + [os.path.join(CSRC, s) for s in DIAG_SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return DIAG_SO


def build(force: bool = False, verbose: bool = False, defines=(), out: str = SO) -> str:
    """defines / out: tuning variants (e.g. defines=["UAVENV_NT_SMALL=256"], out=".../variants/x.so") selected at run
    time with the UAVENV_SO environment variable; the default build takes neither."""
    if not force and out == SO and up_to_date():
        return SO
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_diag(force="--force" in sys.argv))
