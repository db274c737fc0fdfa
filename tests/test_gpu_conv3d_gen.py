"""GPU parity of the general tcgen05 3x3x3 convolution (include/conv3d_gen.h) and of the 1x1 / transposed-convolution GEMM routes
(transoar_b200/conv3d_gen.py) against torch's fp64 convolutions on the same inputs.

Small integers are exact in TF32 and their sums exact in fp32, so the C-ABI entry points must return the exact integer result: a wrong
tap, parity class, halo voxel, channel chunk or a dropped tile shows as an integer difference.  Random-data tests bound the TF32 error."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last_3d)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _taps(w):
    """[CO, CI, 3, 3, 3] -> the tap-major [27, CO, CI] tensor the C ABI takes (include/conv3d_gen.h)."""
    return w.permute(2, 3, 4, 0, 1).reshape(27, w.shape[0], w.shape[1]).contiguous()


def _ints(g, lo, hi, shape):
    return torch.randint(lo, hi + 1, shape, generator=g).float().to(DEV)


# (N, D, H, W), CI, CO: the model's channel pairs, ragged volumes, odd extents (stride 2 parity classes of unequal size),
# one case with more tiles than SMs (persistent loop), one with several column tiles (CO = 384) and one with ragged channel chunks (40 / 72)
CASES = [((1, 4, 8, 16), 24, 48), ((1, 3, 37, 21), 24, 48), ((1, 2, 33, 10), 96, 40), ((2, 5, 7, 9), 48, 48), ((1, 6, 10, 12), 48, 96), ((1, 3, 5, 8), 96, 384), ((1, 5, 5, 8), 384, 192),
         ((1, 3, 6, 7), 40, 72), ((1, 20, 40, 64), 48, 48), ((2, 2, 3, 2), 768, 768),
         # edges: the smallest channel counts, a single voxel row / column, more samples than the model uses, a volume of one 16 x 16 tile pair
         ((3, 2, 2, 2), 4, 4), ((1, 2, 2, 17), 8, 12), ((4, 3, 16, 16), 24, 24), ((1, 2, 40, 8), 100, 36)]


@pytest.fixture(params=["auto", "tap", "halo"])
def path(request):
    """Both forward / input-gradient kernels on every case (include/conv3d_gen.h: conv3d_gen_set_path)."""
    from transoar_b200 import _lib
    _lib.lib().conv3d_gen_set_path({"auto": 0, "tap": 1, "halo": 2}[request.param])
    yield request.param
    _lib.lib().conv3d_gen_set_path(0)


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("shape,ci,co", CASES)
def test_forward_exact_on_integers(shape, ci, co, stride, path):
    from transoar_b200 import _lib
    N, D, H, W = shape
    g = torch.Generator().manual_seed(D * H + W + ci)
    x = _cl(_ints(g, -3, 3, (N, ci, D, H, W)))
    w = _cl(_ints(g, -2, 2, (co, ci, 3, 3, 3)))
    b = _ints(g, -5, 5, (co,))
    od, oh, ow = ((v + stride - 1) // stride for v in (D, H, W))
    y = _cl(torch.full((N, co, od, oh, ow), float("nan"), device=DEV))
    rc = _lib.lib().conv3d_gen_forward(_stream(), _p(x), _p(_taps(w)), _p(b), N, D, H, W, ci, co, stride, _p(y))
    assert rc == 0
    ref = F.conv3d(x.double(), w.double(), b.double(), stride, 1).round()
    assert y.shape == ref.shape and torch.equal(y.double(), ref)


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("shape,ci,co", CASES)
def test_input_gradient_exact_on_integers(shape, ci, co, stride, path):
    from transoar_b200 import _lib
    N, D, H, W = shape
    g = torch.Generator().manual_seed(D * H + W + co)
    od, oh, ow = ((v + stride - 1) // stride for v in (D, H, W))
    dy = _cl(_ints(g, -3, 3, (N, co, od, oh, ow)))
    w = _cl(_ints(g, -2, 2, (co, ci, 3, 3, 3)))
    dx = _cl(torch.full((N, ci, D, H, W), float("nan"), device=DEV))
    rc = _lib.lib().conv3d_gen_dgrad(_stream(), _p(dy), _p(_taps(w)), N, D, H, W, ci, co, stride, _p(dx))
    assert rc == 0
    ref = torch.nn.grad.conv3d_input((N, ci, D, H, W), w.double(), dy.double(), stride=stride, padding=1).round()
    assert torch.equal(dx.double(), ref)


@pytest.mark.parametrize("shape,ci,co", [c for c in CASES if c[1] <= 64])
def test_folded_stride2_input_gradient_exact_on_integers(shape, ci, co):
    """The class-folded stride-2 input gradient (narrow layers) through its own entry point and the python weight folding."""
    from transoar_b200 import _lib
    from transoar_b200.conv3d_gen import fold_stride2_weights
    N, D, H, W = shape
    g = torch.Generator().manual_seed(D * H + W + co + 1)
    od, oh, ow = ((v + 1) // 2 for v in (D, H, W))
    dy = _cl(_ints(g, -3, 3, (N, co, od, oh, ow)))
    w = _cl(_ints(g, -2, 2, (co, ci, 3, 3, 3)))
    dx = _cl(torch.full((N, ci, D, H, W), float("nan"), device=DEV))
    rc = _lib.lib().conv3d_gen_dgrad_s2_folded(_stream(), _p(dy), _p(fold_stride2_weights(_taps(w))), N, D, H, W, ci, co, _p(dx))
    assert rc == 0
    ref = torch.nn.grad.conv3d_input((N, ci, D, H, W), w.double(), dy.double(), stride=2, padding=1).round()
    assert torch.equal(dx.double(), ref)


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("shape,ci,co", CASES)
def test_weight_gradient_exact_on_integers(shape, ci, co, stride):
    from transoar_b200 import _lib
    N, D, H, W = shape
    g = torch.Generator().manual_seed(D * H + W + ci + co)
    od, oh, ow = ((v + stride - 1) // stride for v in (D, H, W))
    x = _cl(_ints(g, -2, 2, (N, ci, D, H, W)))
    dy = _cl(_ints(g, -2, 2, (N, co, od, oh, ow)))
    dw = _cl(torch.full((co, ci, 3, 3, 3), float("nan"), device=DEV))
    rc = _lib.lib().conv3d_gen_wgrad(_stream(), _p(x), _p(dy), N, D, H, W, ci, co, stride, _p(dw))
    assert rc == 0
    ref = torch.nn.grad.conv3d_weight(x.double(), (co, ci, 3, 3, 3), dy.double(), stride=stride, padding=1).round()
    assert torch.equal(dw.double(), ref)


@pytest.mark.parametrize("stride,ci,co,bias", [(1, 96, 384, True), (2, 24, 48, False), (1, 48, 48, False), (2, 192, 384, False)])
def test_autograd_function_matches_fp64(stride, ci, co, bias):
    from transoar_b200.conv3d_gen import conv3d_k3_gen
    g = torch.Generator().manual_seed(ci + co)
    x = _cl(torch.randn(2, ci, 6, 9, 12, generator=g).to(DEV)).requires_grad_(True)
    w = _cl((torch.randn(co, ci, 3, 3, 3, generator=g) / math.sqrt(27 * ci)).to(DEV)).requires_grad_(True)
    b = torch.randn(co, generator=g).to(DEV).requires_grad_(True) if bias else None
    y = conv3d_k3_gen(x, w, b, stride)
    dy = torch.randn(y.shape, generator=g).to(DEV)
    y.backward(dy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = F.conv3d(xd, wd, bd, stride, 1)
    yd.backward(dy.double())
    rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())
    assert y.is_contiguous(memory_format=torch.channels_last_3d)
    assert rel(y, yd) < 3e-3 and rel(x.grad, xd.grad) < 3e-3 and rel(w.grad, wd.grad) < 3e-3
    if bias:
        assert rel(b.grad, bd.grad) < 1e-5


def test_rejects_bad_arguments():
    from transoar_b200 import _lib
    lib = _lib.lib()
    x = torch.zeros(1 << 16, device=DEV)
    p = _p(x)
    assert lib.conv3d_gen_supported(24, 48, 2) == 1 and lib.conv3d_gen_supported(24, 48, 3) == 0 and lib.conv3d_gen_supported(6, 48, 1) == 0
    assert lib.conv3d_gen_forward(None, None, p, None, 1, 4, 4, 4, 8, 8, 1, p) == -1
    assert lib.conv3d_gen_forward(None, p, p, None, 1, 4, 4, 4, 6, 8, 1, p) == -1             # CI % 4
    assert lib.conv3d_gen_forward(None, p, p, None, 1, 4, 4, 4, 8, 8, 3, p) == -1             # stride
    assert lib.conv3d_gen_dgrad(None, p, p, 1, 1, 4, 4, 8, 8, 2, p) == -1                      # stride 2 needs >= 2 voxels per axis
    assert lib.conv3d_gen_wgrad(None, ctypes.c_void_p(x.data_ptr() + 4), p, 1, 4, 4, 4, 8, 8, 1, p) == -3


@pytest.mark.parametrize("ci,co", [(96, 96), (768, 384), (192, 192)])
def test_lateral_1x1_matches_fp64(ci, co):
    from transoar_b200.conv3d_gen import conv3d_1x1
    g = torch.Generator().manual_seed(ci)
    x = _cl(torch.randn(2, ci, 5, 6, 8, generator=g).to(DEV)).requires_grad_(True)
    w = (torch.randn(co, ci, 1, 1, 1, generator=g) / math.sqrt(ci)).to(DEV).requires_grad_(True)
    b = torch.randn(co, generator=g).to(DEV).requires_grad_(True)
    y = conv3d_1x1(x, w, b)
    dy = torch.randn(y.shape, generator=g).to(DEV)
    y.backward(dy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yd = F.conv3d(xd, wd, bd)
    yd.backward(dy.double())
    rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())
    assert y.is_contiguous(memory_format=torch.channels_last_3d)
    assert rel(y, yd) < 3e-3 and rel(x.grad, xd.grad) < 3e-3 and rel(w.grad, wd.grad) < 3e-3 and rel(b.grad, bd.grad) < 1e-4


@pytest.mark.parametrize("ci,co,skip", [(384, 384, True), (384, 192, True), (192, 96, False)])
def test_transposed_k2s2_matches_fp64(ci, co, skip):
    from transoar_b200.conv3d_gen import conv_transpose3d_k2s2
    g = torch.Generator().manual_seed(ci + co)
    x = _cl(torch.randn(2, ci, 3, 5, 4, generator=g).to(DEV)).requires_grad_(True)
    w = (torch.randn(ci, co, 2, 2, 2, generator=g) / math.sqrt(ci)).to(DEV).requires_grad_(True)
    b = torch.randn(co, generator=g).to(DEV).requires_grad_(True)
    s = _cl(torch.randn(2, co, 6, 10, 8, generator=g).to(DEV)).requires_grad_(True) if skip else None
    y = conv_transpose3d_k2s2(x, w, b, s)
    dy = torch.randn(y.shape, generator=g).to(DEV)
    y.backward(dy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yd = F.conv_transpose3d(xd, wd, bd, stride=2)
    if skip:
        yd = yd + s.detach().double()
    yd.backward(dy.double())
    rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())
    assert y.is_contiguous(memory_format=torch.channels_last_3d)
    assert rel(y, yd) < 3e-3 and rel(x.grad, xd.grad) < 3e-3 and rel(w.grad, wd.grad) < 3e-3 and rel(b.grad, bd.grad) < 1e-4
    if skip:
        assert torch.equal(s.grad, dy)
