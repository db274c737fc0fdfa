"""GPU parity of the tcgen05 implicit-GEMM 3x3x3 convolution (include/conv3d_tc.h) against torch's conv3d in fp64 on the same
inputs (TF32 tolerance), forward and input gradient, ragged shapes included."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _err(a, b):
    return float((a.double() - b.double()).abs().max())


@pytest.mark.parametrize("shape", [(1, 3, 16, 8), (2, 5, 32, 24), (1, 4, 20, 13), (1, 2, 7, 5), (1, 9, 48, 40)])
@pytest.mark.parametrize("c", [24, 8, 16])
def test_forward_and_input_gradient_match_conv3d(shape, c):
    from transoar_b200.conv3d_tc import conv3d_k3
    N, D, H, W = shape
    g = torch.Generator().manual_seed(N + D + H + W + c)
    x = torch.randn(N, c, D, H, W, generator=g).to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    w = (torch.randn(c, c, 3, 3, 3, generator=g) / math.sqrt(27 * c)).to(DEV).requires_grad_(True)
    dy = torch.randn(N, c, D, H, W, generator=g).to(DEV)
    y = conv3d_k3(x, w)
    y.backward(dy)
    assert y.is_contiguous(memory_format=torch.channels_last_3d) and x.grad.shape == x.shape
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yd = F.conv3d(xd, wd, None, 1, 1)
    yd.backward(dy.double())
    tol = 4e-3 * math.sqrt(27 * c) * 4.5 * 4.5 / math.sqrt(27 * c)            # TF32: ~2^-10 relative per product, |x|, |dy| up to ~4.5
    assert _err(y, yd) < tol, (_err(y, yd), tol)
    assert _err(x.grad, xd.grad) < tol
    assert _err(w.grad, wd.grad) < 4e-3 * math.sqrt(N * D * H * W) * 4.5 * 4.5 / 9


def test_exact_on_tf32_representable_inputs():
    """Small integers are exact in TF32 and their sums exact in fp32: the result must be the exact integer convolution (torch's fp64
    GPU convolution itself is off by ~1e-13 from the integers, hence the round())."""
    from transoar_b200.conv3d_tc import conv3d_k3
    g = torch.Generator().manual_seed(1)
    x = torch.randint(-4, 5, (1, 24, 6, 20, 11), generator=g).float().to(DEV).contiguous(memory_format=torch.channels_last_3d)
    w = torch.randint(-3, 4, (24, 24, 3, 3, 3), generator=g).float().to(DEV)
    assert torch.equal(conv3d_k3(x, w).double(), F.conv3d(x.double(), w.double(), None, 1, 1).round())


@pytest.mark.parametrize("shape,ci,co", [((2, 6, 20, 11), 24, 24), ((1, 3, 16, 8), 8, 16), ((1, 5, 33, 17), 16, 8), ((3, 2, 7, 5), 24, 24),
                                         ((1, 40, 48, 40), 24, 24)])
def test_weight_gradient_exact_on_integers(shape, ci, co):
    """The weight gradient on its own entry point: small integers -> every product and partial sum is exact, so dweight must equal the
    fp64 conv3d_weight result exactly -- a wrong tap, a wrong halo row or a dropped tile shows as an integer difference.  The last
    shape has more tiles than CTAs (the persistent loop) and rows that cross the 16 x 8 tile raggedly."""
    import ctypes
    from transoar_b200 import _lib
    N, D, H, W = shape
    g = torch.Generator().manual_seed(D * H + W)
    x = torch.randint(-2, 3, (N, ci, D, H, W), generator=g).float().to(DEV).contiguous(memory_format=torch.channels_last_3d)
    dy = torch.randint(-2, 3, (N, co, D, H, W), generator=g).float().to(DEV).contiguous(memory_format=torch.channels_last_3d)
    dw = torch.full((co, ci, 3, 3, 3), float("nan"), device=DEV)
    ws = torch.empty(_lib.lib().conv3d_tc_wgrad_workspace_floats(), device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = _lib.lib().conv3d_tc_k3_wgrad(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), p(x), p(dy), N, D, H, W, ci, co, p(dw), p(ws))
    assert rc == 0
    ref = torch.nn.grad.conv3d_weight(x.double(), (co, ci, 3, 3, 3), dy.double(), stride=1, padding=1).round()
    assert torch.equal(dw.double(), ref)


def test_weight_gradient_rejects_unsupported_channels():
    from transoar_b200 import _lib
    assert _lib.lib().conv3d_tc_k3_wgrad(None, None, None, 1, 4, 4, 4, 24, 24, None, None) < 0
    x = torch.zeros(64 * 40 * 4 * 4 * 4, device=DEV)
    import ctypes
    p = ctypes.c_void_p(x.data_ptr())
    assert _lib.lib().conv3d_tc_k3_wgrad(None, p, p, 1, 4, 4, 4, 40, 24, p, p) < 0
    assert _lib.lib().conv3d_tc_k3_wgrad(None, p, p, 1, 4, 4, 4, 24, 6, p, p) < 0


@pytest.mark.parametrize("row0", [0, 1, 2, 3, 5, 12])
def test_mn_major_overlapping_slab_probe(row0):
    """The operand form a tensor-core weight gradient over channels-last volumes needs (see conv3d_tc_kernels.cuh): MN-major, 128-byte
    swizzle with 32-byte atoms, four slabs one row apart (= the kw taps), start address at any row."""
    import ctypes
    from transoar_b200 import _lib
    g = torch.Generator().manual_seed(2 + row0)
    X = torch.randint(-4, 5, (24, 32), generator=g).float().to(DEV)
    Y = torch.randint(-4, 5, (8, 32), generator=g).float().to(DEV)
    D = torch.zeros(128, 32, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = _lib.lib().conv3d_tc_debug_mn_probe(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), p(X), p(Y), p(D), row0)
    torch.cuda.synchronize()
    assert rc == 0
    want = torch.cat([X[row0 + j:row0 + j + 8].t() @ Y for j in range(4)], 0)          # [4 * 32, 32]
    assert torch.equal(D, want)
