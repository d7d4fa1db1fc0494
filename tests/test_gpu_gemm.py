"""uavnet_gemm (include/uavnet.h): the dense layers of the reference's MLPs (main.py:148-149,152-153) and their
gradients on the tcgen05 tensor cores, against float64 products of the same operands.

Tolerances (written here, relative to the largest |entry| of the exact result):
  precision="fp32" (3xTF32, three MMAs per k-step):  5e-5  (fp32-class; the tensor core's accumulator truncates)
  precision="tf32" (operands cut to 10 mantissa bits): 2e-3 against the exact product, and 5e-5 against the product
      of operands truncated to TF32 the way the hardware reads them (this one pins the data path, not the rounding).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {"fp32": 5e-5, "tf32": 2e-3}


@pytest.fixture(scope="module")
def dense():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from drl_uav_cellularnet_b200 import dense as d
    yield d
    assert d.check() == 0, "a uavnet_gemm launch gave up on a barrier"


def tf32_trunc(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)


def rel_err(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("at,bt", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (1, 1, 1), (37, 5, 9), (300, 200, 200), (300, 625, 200), (257, 200, 625),
                                   (200, 625, 3000), (129, 257, 70)])
def test_plain_products_all_operand_layouts(dense, prec, at, bt, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + at * 2 + bt)
    A = torch.randn((K, M) if at else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if bt else (K, N), device="cuda", generator=g)
    D = dense.gemm(A, B, a_trans=bool(at), b_trans=bool(bt), precision=prec)
    ref = (A.t() if at else A).double() @ (B.t() if bt else B).double()
    assert D.shape == ref.shape
    assert rel_err(D, ref) < TOL[prec]
    if prec == "tf32":
        At, Bt = tf32_trunc(A), tf32_trunc(B)
        ref_t = (At.t() if at else At).double() @ (Bt.t() if bt else Bt).double()
        assert rel_err(D, ref_t) < 5e-5


def test_forward_layer_bias_relu6_on_a_slice_of_wider_activations(dense):
    """h2 = relu6(h1[:, H:] @ W2 + b2) (main.py:152) with the value head fused: v = h2 @ w3 + b3 (main.py:153)"""
    g = torch.Generator(device="cuda").manual_seed(1)
    M, H = 1000, 200
    h1 = torch.rand((M, 2 * H), device="cuda", generator=g) * 6
    W2 = torch.randn((H, H), device="cuda", generator=g) * 0.1
    b2 = torch.randn(H, device="cuda", generator=g)
    w3 = torch.randn((H, 1), device="cuda", generator=g) * 0.1
    b3 = torch.full((1,), 0.3, device="cuda")
    v = torch.empty(M, device="cuda")
    h2 = dense.gemm(h1[:, H:], W2, bias=b2, relu6=True, dot_w=w3.view(-1), dot_b=b3, dot_out=v)
    ref = torch.clamp(h1[:, H:].double() @ W2.double() + b2.double(), 0, 6)
    assert float((h2.double() - ref).abs().max()) < 5e-5
    assert ((h2 >= 0) & (h2 <= 6)).all()
    ref_v = (ref @ w3.double()).squeeze(1) + 0.3
    assert float((v.double() - ref_v).abs().max()) < 5e-5
    # the value alone (no D)
    v2 = torch.empty(M, device="cuda")
    dense.gemm(h1[:, H:], W2, bias=b2, relu6=True, dot_w=w3.view(-1), dot_b=b3, dot_out=v2, want_out=False)
    assert torch.equal(v, v2)


def test_logits_with_625_columns_and_unaligned_weights(dense):
    """logits = h2 @ Wa3 + ba3 (main.py:149): 625-float rows are not 16-byte multiples -> scalar staging path"""
    g = torch.Generator(device="cuda").manual_seed(2)
    M, H, A = 500, 200, 625
    h2 = torch.rand((M, H), device="cuda", generator=g) * 6
    W = torch.randn((H, A), device="cuda", generator=g) * 0.1
    b = torch.randn(A, device="cuda", generator=g)
    z = dense.gemm(h2, W, bias=b)
    ref = h2.double() @ W.double() + b.double()
    assert float((z.double() - ref).abs().max()) < 1e-4     # logits up to ~10: 1e-5 relative
    # the same weights stored with a padded leading dimension (vector staging path): identical results
    Wp = torch.zeros((H, 640), device="cuda")
    Wp[:, :A] = W
    z2 = dense.gemm(h2, Wp[:, :A], bias=b)
    assert torch.equal(z, z2)
    # an operand whose base pointer is only 4-byte aligned
    flat = torch.zeros(H * A + 1, device="cuda")
    flat[1:] = W.reshape(-1)
    z3 = dense.gemm(h2, flat[1:].view(H, A), bias=b)
    assert torch.equal(z, z3)


def test_data_gradient_with_relu6_mask(dense):
    """dpre = (dy @ W^T) * relu6'(h) (backward of main.py:148/152), written into a slice of a wider buffer"""
    g = torch.Generator(device="cuda").manual_seed(3)
    M, H, A = 700, 200, 625
    dy = torch.randn((M, A), device="cuda", generator=g)
    W = torch.randn((H, A), device="cuda", generator=g) * 0.1
    h = torch.rand((M, 2 * H), device="cuda", generator=g) * 8 - 1          # some units outside (0, 6)
    h[5, 3] = 0.0
    h[6, 4] = 6.0
    out = torch.full((M, 2 * H), 7.0, device="cuda")
    dense.gemm(dy, W, out[:, :H], b_trans=True, mask_src=h[:, :H])
    hm = h[:, :H]
    ref = (dy.double() @ W.double().t()) * ((hm > 0) & (hm < 6))
    assert rel_err(out[:, :H], ref) < TOL["fp32"]
    assert float(out[5, 3]) == 0.0 and float(out[6, 4]) == 0.0
    assert (out[:, H:] == 7.0).all(), "columns outside the output slice were touched"
    # the bias gradient of the layer below = column sums of what was stored, from the same epilogue
    gb = torch.ones(H, device="cuda")
    out2 = torch.empty((M, H), device="cuda")
    dense.gemm(dy, W, out2, b_trans=True, mask_src=h[:, :H], out_colsum=gb)
    assert torch.equal(out2, out[:, :H])
    assert rel_err(gb, 1.0 + ref.sum(0)) < TOL["fp32"]


def test_rank1_masked_product(dense):
    """uavnet_rank1_mask: (dv (x) wc3) * relu6'(h2c), the value head's data gradient (main.py:153)"""
    import ctypes as C
    from drl_uav_cellularnet_b200 import _native as N
    g = torch.Generator(device="cuda").manual_seed(8)
    M, H = 5000, 200
    dv = torch.randn(M, device="cuda", generator=g)
    w = torch.randn(H, device="cuda", generator=g)
    h = torch.rand((M, H), device="cuda", generator=g) * 8 - 1
    out = torch.empty((M, H), device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    assert N.lib().uavnet_rank1_mask(vp(dv), vp(w), vp(h), M, H, vp(out), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(out, (dv[:, None] * w[None, :]) * ((h > 0) & (h < 6)))
    assert N.lib().uavnet_rank1_mask(vp(dv), vp(w), vp(h), M, 201, vp(out), None) == -1


@pytest.mark.parametrize("split_k", [0, 1, 7])
def test_weight_gradient_accumulates_with_split_k_and_bias_gradient(dense, split_k):
    """gW += x^T @ dy and gb += column sums of dy from the same pass (a row of ones appended to x^T)"""
    g = torch.Generator(device="cuda").manual_seed(4)
    M, H, A = 4000, 200, 625
    x = torch.rand((M, 2 * H), device="cuda", generator=g) * 6
    dy = torch.randn((M, A), device="cuda", generator=g) * 1e-2
    gW = torch.randn((H, A), device="cuda", generator=g)
    gb = torch.randn(A, device="cuda", generator=g)
    gW0, gb0 = gW.clone(), gb.clone()
    dense.gemm(x[:, H:], dy, gW, a_trans=True, accumulate=True, split_k=split_k, colsum=gb)
    ref_w = gW0.double() + x[:, H:].double().t() @ dy.double()
    ref_b = gb0.double() + dy.double().sum(0)
    assert rel_err(gW, ref_w) < TOL["fp32"]
    assert rel_err(gb, ref_b) < TOL["fp32"]


def test_column_sums_only_and_single_column_products(dense):
    g = torch.Generator(device="cuda").manual_seed(5)
    M, H = 3000, 400
    d1 = torch.randn((M, H), device="cuda", generator=g)
    gb = torch.zeros(H, device="cuda")
    dense.gemm(None, d1, colsum=gb, accumulate=True)
    assert rel_err(gb, d1.double().sum(0)) < TOL["fp32"]
    # gWc3 += h2c^T @ dv  (N = 1) with its bias gradient
    h2c = torch.rand((M, 200), device="cuda", generator=g) * 6
    dv = torch.randn((M, 1), device="cuda", generator=g)
    gw, gb3 = torch.zeros((200, 1), device="cuda"), torch.zeros(1, device="cuda")
    dense.gemm(h2c, dv, gw, a_trans=True, accumulate=True, colsum=gb3)
    assert rel_err(gw, h2c.double().t() @ dv.double()) < TOL["fp32"]
    assert abs(float(gb3) - float(dv.double().sum())) < 1e-3
    # rank-1 data gradient: dpre2c = (dv (x) wc3) * mask   (K = 1)
    wc3 = torch.randn((200, 1), device="cuda", generator=g)
    out = dense.gemm(dv, wc3, b_trans=True, mask_src=h2c)
    ref = (dv.double() @ wc3.double().t()) * ((h2c > 0) & (h2c < 6))
    assert rel_err(out, ref) < TOL["fp32"]


def test_full_update_batch_shapes(dense):
    """the shapes of one A3C update (81 920 samples): deterministic across calls for the non-atomic products"""
    g = torch.Generator(device="cuda").manual_seed(6)
    M = 81920
    dz = torch.randn((M, 625), device="cuda", generator=g) * 1e-3
    Wa3 = torch.randn((200, 625), device="cuda", generator=g) * 0.1
    h2a = torch.rand((M, 200), device="cuda", generator=g) * 6
    a = dense.gemm(dz, Wa3, b_trans=True, mask_src=h2a, precision="tf32")
    b = dense.gemm(dz, Wa3, b_trans=True, mask_src=h2a, precision="tf32")
    assert torch.equal(a, b)
    rows = torch.randint(0, M, (512,), device="cuda", generator=g)
    ref = (dz[rows].double() @ Wa3.double().t()) * ((h2a[rows] > 0) & (h2a[rows] < 6))
    assert rel_err(a[rows], ref) < TOL["tf32"]
    gW = torch.zeros((200, 625), device="cuda")
    gb = torch.zeros(625, device="cuda")
    dense.gemm(h2a, dz, gW, a_trans=True, accumulate=True, colsum=gb)
    assert rel_err(gW, h2a.double().t() @ dz.double()) < TOL["fp32"]
    assert rel_err(gb, dz.double().sum(0)) < TOL["fp32"]


def test_rejections(dense):
    A = torch.zeros((8, 8), device="cuda")
    with pytest.raises(ValueError):
        dense.gemm(A, torch.zeros((9, 8), device="cuda"))                       # reduction lengths differ
    with pytest.raises(ValueError):
        dense.gemm(A.double(), A)                                               # dtype
    with pytest.raises(ValueError):
        dense.gemm(A.t(), A)                                                    # last dimension not contiguous
    with pytest.raises(RuntimeError):
        dense.gemm(A, A, torch.zeros((8, 8), device="cuda"), accumulate=True, bias=torch.zeros(8, device="cuda"))
    with pytest.raises(RuntimeError):
        dense.gemm(A, A, split_k=4)                                             # split-K needs accumulate
    with pytest.raises(ValueError):
        dense.gemm(A, A, torch.zeros((4, 8), device="cuda"))                    # wrong output shape
    assert np.isfinite(float(dense.gemm(A, A).sum()))
