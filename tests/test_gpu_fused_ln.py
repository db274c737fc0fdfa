"""GPU parity of y = LayerNorm(a + dropout(b)) (include/fused_ln.h) against the ATen composition in fp64, plus the statistical
and consistency properties of its hash-based dropout."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(2, 1000, 384), (3, 7, 48), (1, 1, 4), (5, 96), (2, 33, 1024), (4, 100, 640), (1, 3000, 132)])
def test_no_dropout_matches_layer_norm_of_the_sum(shape):
    from transoar_b200 import _lib
    from transoar_b200.fused_ln import add_dropout_layer_norm
    g = torch.Generator().manual_seed(sum(shape))
    C = shape[-1]
    norm = nn.LayerNorm(C).to(DEV)
    with torch.no_grad():
        norm.weight.copy_(torch.rand(C, generator=g) + 0.5)
        norm.bias.copy_(torch.randn(C, generator=g) * 0.3)
    a = (torch.randn(*shape, generator=g) * 2 + 1).to(DEV).requires_grad_(True)
    b = torch.randn(*shape, generator=g).to(DEV).requires_grad_(True)
    dy = torch.randn(*shape, generator=g).to(DEV)
    n0 = _lib.lib().msda3d_launch_count()
    y = add_dropout_layer_norm(a, b, norm, 0.1, training=False)
    y.backward(dy)
    assert _lib.lib().msda3d_launch_count() - n0 == 3
    ad, bd = a.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    wd, bid = norm.weight.detach().double().requires_grad_(True), norm.bias.detach().double().requires_grad_(True)
    yd = F.layer_norm(ad + bd, (C,), wd, bid, norm.eps)
    yd.backward(dy.double())
    assert _rel(y, yd) < 1e-5
    assert _rel(a.grad, ad.grad) < 1e-4 and _rel(b.grad, bd.grad) < 1e-4
    assert _rel(norm.weight.grad, wd.grad) < 1e-4 and _rel(norm.bias.grad, bid.grad) < 1e-4


def test_dropout_mask_is_consistent_between_forward_and_backward_and_has_the_right_rate():
    """With gamma = 1, beta = 0 and a = 0 the forward is LN(dropout(b)); the backward's db must be zero exactly where the forward
    dropped, and equal to da / (1 - p) elsewhere.  The keep rate must be 1 - p within sampling noise, the seed must matter."""
    from transoar_b200.fused_ln import AddDropoutLayerNormFunction
    rows, C, p = 4096, 384, 0.1
    g = torch.Generator().manual_seed(0)
    w, bias = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    a = torch.zeros(rows, C, device=DEV).requires_grad_(True)
    b = (torch.rand(rows, C, generator=g) + 1).to(DEV).requires_grad_(True)          # strictly positive: dropped <=> z == 0
    y = AddDropoutLayerNormFunction.apply(a, b, w, bias, 1e-5, p, 1234)
    (zsaved,) = [t for t in y.grad_fn.saved_tensors if t.shape == (rows, C)][:1]
    dropped = zsaved == 0
    rate = float(dropped.float().mean())
    assert abs(rate - p) < 4 * (p * (1 - p) / (rows * C)) ** 0.5 + 1e-4, rate
    assert torch.allclose(zsaved[~dropped], b.detach()[~dropped] / (1 - p), rtol=1e-6)
    y.backward(torch.randn(rows, C, generator=g).to(DEV))
    assert bool((b.grad[dropped] == 0).all())
    assert torch.allclose(b.grad[~dropped], a.grad[~dropped] / (1 - p), rtol=1e-6, atol=0)
    # per-column and per-row keep rates are flat (no structure in the hash)
    assert float(dropped.float().mean(0).max()) < p + 0.03 and float(dropped.float().mean(1).max()) < p + 0.08
    y2 = AddDropoutLayerNormFunction.apply(a, b, w, bias, 1e-5, p, 1235)
    z2 = [t for t in y2.grad_fn.saved_tensors if t.shape == (rows, C)][0]
    assert 0.15 < float(((z2 == 0) != dropped).float().mean()) < 0.21                # two independent masks differ at 2p(1-p) = 0.18


def test_training_mode_is_reproducible_under_manual_seed_and_falls_back_off_device():
    from transoar_b200.fused_ln import add_dropout_layer_norm
    norm = nn.LayerNorm(64).to(DEV)
    a, b = torch.randn(10, 64, device=DEV), torch.randn(10, 64, device=DEV)
    torch.manual_seed(5)
    y1 = add_dropout_layer_norm(a, b, norm, 0.3, True)
    torch.manual_seed(5)
    y2 = add_dropout_layer_norm(a, b, norm, 0.3, True)
    assert torch.equal(y1, y2) and not torch.equal(y1, add_dropout_layer_norm(a, b, norm, 0.3, True))
    ncpu = nn.LayerNorm(64)
    out = add_dropout_layer_norm(a.cpu(), b.cpu(), ncpu, 0.3, False)                      # CPU tensors: the plain composition
    assert torch.allclose(out, ncpu(a.cpu() + b.cpu()))


@pytest.mark.parametrize("shape", [(2, 1000, 384), (3, 7, 48), (2, 33, 1024)])
@pytest.mark.parametrize("autocast", [False, True])
def test_bf16_branch_with_fp32_residual_stream(shape, autocast):
    """The bf16 route: a (residual stream) fp32, b (branch, output of a bf16 GEMM) bf16 -- the dtypes around norm(x + dropout(y)) under
    torch.autocast(bfloat16).  Output and da stay fp32, db is bf16; same kernel count as the fp32 case."""
    import contextlib
    from transoar_b200 import _lib
    from transoar_b200.fused_ln import add_dropout_layer_norm
    g = torch.Generator().manual_seed(sum(shape) + 5)
    C = shape[-1]
    norm = nn.LayerNorm(C).to(DEV)
    a = (torch.randn(*shape, generator=g) * 2 + 1).to(DEV).requires_grad_(True)
    b = torch.randn(*shape, generator=g).to(DEV).to(torch.bfloat16).requires_grad_(True)
    dy = torch.randn(*shape, generator=g).to(DEV)
    n0 = _lib.lib().msda3d_launch_count()
    with (torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()):
        y = add_dropout_layer_norm(a, b, norm, 0.1, training=False)
    y.backward(dy)
    assert _lib.lib().msda3d_launch_count() - n0 == 3
    assert y.dtype == torch.float32 and a.grad.dtype == torch.float32 and b.grad.dtype == torch.bfloat16
    ad, bd = a.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    wd, bid = norm.weight.detach().double().requires_grad_(True), norm.bias.detach().double().requires_grad_(True)
    yd = F.layer_norm(ad + bd, (C,), wd, bid, norm.eps)
    yd.backward(dy.double())
    assert _rel(y, yd) < 1e-5 and _rel(a.grad, ad.grad) < 1e-4
    assert _rel(b.grad, bd.grad) < 1e-2                                              # one bf16 rounding
    assert _rel(norm.weight.grad, wd.grad) < 1e-4 and _rel(norm.bias.grad, bid.grad) < 1e-4


@pytest.mark.parametrize("shape,dtype", [((2, 6, 7, 9, 48), torch.float32), ((3, 1000, 384), torch.float32), ((2, 5, 5, 8, 768), torch.bfloat16),
                                         ((4, 100, 96), torch.bfloat16)])
def test_plain_layer_norm_on_the_warp_per_row_kernel(shape, dtype):
    from transoar_b200 import _lib
    from transoar_b200.fused_ln import layer_norm
    g = torch.Generator().manual_seed(sum(shape))
    C = shape[-1]
    norm = nn.LayerNorm(C).to(DEV)
    with torch.no_grad():
        norm.weight.copy_(torch.rand(C, generator=g) + 0.5)
        norm.bias.copy_(torch.randn(C, generator=g) * 0.3)
    x = (torch.randn(*shape, generator=g) * 2 + 1).to(DEV).to(dtype).requires_grad_(True)
    dy = torch.randn(*shape, generator=g).to(DEV)
    n0 = _lib.lib().msda3d_launch_count()
    y = layer_norm(x, norm)
    y.backward(dy)
    assert _lib.lib().msda3d_launch_count() - n0 == 3 and y.dtype == torch.float32 and x.grad.dtype == dtype
    xd = x.detach().double().requires_grad_(True)
    wd, bd = norm.weight.detach().double().requires_grad_(True), norm.bias.detach().double().requires_grad_(True)
    yd = F.layer_norm(xd, (C,), wd, bd, norm.eps)
    yd.backward(dy.double())
    assert _rel(y, yd) < 1e-5 and _rel(x.grad, xd.grad) < (1e-4 if dtype == torch.float32 else 1e-2)
    assert _rel(norm.weight.grad, wd.grad) < 1e-4 and _rel(norm.bias.grad, bd.grad) < 1e-4
