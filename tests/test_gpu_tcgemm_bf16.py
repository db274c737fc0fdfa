"""GPU parity of the bf16 tcgen05 GEMM (include/tc_gemm.h: tc_gemm_bf16, tcgen05.mma kind::f16) and of the bf16 Linear route built on
it.  Operands are exactly representable in bf16, accumulation is fp32 in TMEM, so against an fp64 product of the SAME bf16 operands the
error is fp32-accumulation level (plus one bf16 rounding when the result is stored as bf16)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _operand(rows, red, mn_major, gen):
    if mn_major:
        st = torch.randn(red, rows, generator=gen).to(torch.bfloat16).cuda()
        return st, st.double().t(), rows
    st = torch.randn(rows, red, generator=gen).to(torch.bfloat16).cuda()
    return st, st.double(), red


SHAPES = [(128, 128, 64), (256, 256, 128), (304, 200, 104), (1000, 384, 384), (520, 1024, 392), (64, 8, 16), (2048, 1024, 1024),
          (20000, 384, 384), (20080, 192, 96), (40000, 1024, 384)]          # the last three fill the machine with CTA pairs


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("M,N,R", SHAPES)
@pytest.mark.parametrize("out", [torch.bfloat16, torch.float32])
def test_bf16_gemm_matches_fp64(a_mn, b_mn, M, N, R, out):
    from transoar_b200.linear import gemm_bf16
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major bf16 operands need a leading dimension that is a multiple of 8 (16-byte TMA strides)")
    gen = torch.Generator().manual_seed(M * 7 + N * 3 + R + a_mn * 2 + b_mn)
    A, Ad, lda = _operand(M, R, a_mn, gen)
    B, Bd, ldb = _operand(N, R, b_mn, gen)
    bias = torch.randn(N, generator=gen).cuda()
    D = torch.full((M, N), float("nan"), device="cuda", dtype=out)
    gemm_bf16(A, a_mn, lda, B, b_mn, ldb, D, M, N, R, bias=bias, relu=True)
    torch.cuda.synchronize()
    want = (Ad @ Bd.t() + bias.double()).clamp_min(0)
    scale = float(want.abs().max())
    tol = (2 ** -8 if out == torch.bfloat16 else 1e-5 * math.sqrt(R)) * scale
    assert float((D.double() - want).abs().max()) <= tol


@pytest.mark.parametrize("M,N,R", [(4096, 384, 2000), (384, 1024, 234000), (1024, 384, 117008)])
def test_bf16_weight_gradient_gemm_split_k_accumulates_in_fp32(M, N, R):
    """dW = dY^T X: both operands MN-major, split-K over a long reduction, partial tiles added into fp32 with vector reductions."""
    from transoar_b200.linear import gemm_bf16
    gen = torch.Generator().manual_seed(R)
    A, Ad, lda = _operand(M, R, 1, gen)
    B, Bd, ldb = _operand(N, R, 1, gen)
    D = torch.zeros(M, N, device="cuda")
    gemm_bf16(A, 1, lda, B, 1, ldb, D, M, N, R, accumulate=True, split_k=0)
    torch.cuda.synchronize()
    want = Ad @ Bd.t()
    assert float((D.double() - want).abs().max()) <= 1e-4 * float(want.abs().max())      # fp32 accumulation over splits, exact bf16 products


def test_bf16_linear_route_under_autocast_matches_cublas_bf16():
    """TCLinear inside torch.autocast(bfloat16): forward and the three gradients against F.linear under the same autocast region
    (cuBLAS bf16) -- both are bf16 products with fp32 accumulation; they may differ by one bf16 rounding of the result."""
    import torch.nn.functional as F
    from transoar_b200 import _lib
    from transoar_b200.linear import TCLinear
    torch.manual_seed(0)
    lin = TCLinear(384, 1024).cuda()
    x = torch.randn(2, 3000, 384, device="cuda", requires_grad=True)
    g = torch.randn(2, 3000, 1024, device="cuda")
    n0 = _lib.lib().msda3d_launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = lin(x)
    assert y.dtype == torch.bfloat16
    y.backward(g.to(torch.bfloat16))
    assert _lib.lib().msda3d_launch_count() == n0 + 3, "forward + two gradient GEMMs on the tcgen05 kernel"
    got = (y.float(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = lin.weight.grad = lin.bias.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = F.linear(x, lin.weight, lin.bias)
    y2.backward(g.to(torch.bfloat16))
    want = (y2.float(), x.grad, lin.weight.grad, lin.bias.grad)
    for a, b, tol in zip(got, want, (2 ** -7, 2 ** -7, 2 ** -7, 2 ** -7)):
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()), (a.shape, float((a - b).abs().max()), float(b.abs().max()))
