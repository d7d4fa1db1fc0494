"""Host wrapper of ``uavnet_gemm`` (include/uavnet.h): the dense layers of the reference's MLPs (main.py:148-149,
152-153) and their gradients on the tcgen05 tensor cores.  Operands are 2-D float32 CUDA tensors whose last dimension
is contiguous (slices of wider activations are fine: the row stride is the leading dimension)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native as N

PRECISIONS = {"tf32": N.GEMM_TF32, "fp32": N.GEMM_3XTF32, "3xtf32": N.GEMM_3XTF32}


def _chk(t: torch.Tensor, name: str):
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
        raise ValueError("%s: need a 2-D float32 CUDA tensor with a contiguous last dimension" % name)
    return t


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def gemm(A: Optional[torch.Tensor], B: torch.Tensor, out: Optional[torch.Tensor] = None, *, a_trans: bool = False,
         b_trans: bool = False, bias: Optional[torch.Tensor] = None, relu6: bool = False,
         mask_src: Optional[torch.Tensor] = None, accumulate: bool = False, split_k: int = 0,
         colsum: Optional[torch.Tensor] = None, out_colsum: Optional[torch.Tensor] = None, dot_w: Optional[torch.Tensor] = None,
         dot_b: Optional[torch.Tensor] = None, dot_out: Optional[torch.Tensor] = None, precision: str = "fp32",
         want_out: bool = True) -> Optional[torch.Tensor]:
    """out[M,N] (+)= op(A) . op(B) with the fused epilogue of uavnet_gemm.
    a_trans: A is given as [K,M] (the product is A^T . B); b_trans: B is given as [N,K] (the product is A . B^T).
    A = None (with colsum): column sums of op(B) only."""
    B = _chk(B, "B")
    if b_trans:
        n, k = B.shape
    else:
        k, n = B.shape
    if A is None:
        m = 0
        lda = 0
    else:
        A = _chk(A, "A")
        m, ka = (A.shape[1], A.shape[0]) if a_trans else (A.shape[0], A.shape[1])
        if ka != k:
            raise ValueError("reduction lengths differ: %d vs %d" % (ka, k))
        lda = A.stride(0)
    if out is None and want_out and m > 0:
        if accumulate:
            raise ValueError("accumulate needs an output tensor")
        out = torch.empty((m, n), dtype=torch.float32, device=B.device)
    d = N.GemmDesc()
    d.A, d.lda, d.a_trans = _p(A), lda, int(a_trans)
    d.B, d.ldb, d.b_trans = _p(B), B.stride(0), int(b_trans)
    if out is not None:
        _chk(out, "out")
        if tuple(out.shape) != (m, n):
            raise ValueError("out has shape %s, expected %s" % (tuple(out.shape), (m, n)))
        d.D, d.ldd = _p(out), out.stride(0)
    d.M, d.N, d.K = m, n, k
    d.bias, d.relu6 = _p(bias), int(relu6)
    if mask_src is not None:
        _chk(mask_src, "mask_src")
        d.mask_src, d.ld_mask = _p(mask_src), mask_src.stride(0)
    d.accumulate, d.split_k = int(accumulate), int(split_k)
    d.colsum, d.out_colsum, d.dot_w, d.dot_b, d.dot_out = _p(colsum), _p(out_colsum), _p(dot_w), _p(dot_b), _p(dot_out)
    d.precision = PRECISIONS[precision]
    stream = C.c_void_p(torch.cuda.current_stream(B.device).cuda_stream)
    rc = N.lib().uavnet_gemm(C.byref(d), stream)
    if rc:
        raise RuntimeError("uavnet_gemm failed (%d)" % rc)
    return out


def check() -> int:
    """Synchronises the device; non-zero if any uavnet_gemm launch gave up on a barrier."""
    return int(N.lib().uavnet_gemm_check())
