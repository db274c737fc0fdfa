"""GPU parity of the fused InstanceNorm3d + ReLU (include/instnorm.h) against torch's instance_norm + relu -- the very ops
the reference's EncoderCnnBlock runs (encoder_blocks.py:28-46) -- in fp64 on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

from transoar_b200.instnorm import instance_norm_relu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(2, 24, 16, 16, 32), (1, 3, 5, 5, 8), (2, 5, 3, 7, 11), (1, 768, 5, 5, 8), (3, 2, 40, 41, 9), (2, 3, 2, 3, 5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_backward_match_torch_instance_norm_relu(shape, dtype):
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 2 + 3).to(DEV)              # large mean / std ratio stresses the variance
    w = (torch.rand(shape[1], generator=g) + 0.5).to(DEV)
    b = (torch.randn(shape[1], generator=g) * 0.3).to(DEV)
    dy = torch.randn(*shape, generator=g).to(DEV)
    xs = x.to(dtype).requires_grad_(True)
    ws, bs = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = instance_norm_relu(xs, ws, bs, 1e-5)
    y.backward(dy.to(dtype))
    xr = xs.detach().double().requires_grad_(True)                       # oracle on the same (rounded) inputs, fp64
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.relu(F.instance_norm(xr, weight=wr, bias=br, eps=1e-5))
    yr.backward(dy.to(dtype).double())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert y.dtype == dtype and _rel(y, yr) < tol
    assert _rel(xs.grad, xr.grad) < (2e-4 if dtype == torch.float32 else 2e-2)
    assert _rel(ws.grad, wr.grad) < (1e-4 if dtype == torch.float32 else 1e-2) and _rel(bs.grad, br.grad) < (1e-4 if dtype == torch.float32 else 1e-2)


def test_visceral_first_stage_shape_and_statistics():
    """24 instances of 160x160x256 voxels (629 MB fp32): mean 0 / var 1 per instance before the affine + ReLU."""
    x = torch.randn(1, 24, 160, 160, 256, device=DEV) * 5 + 100
    w, b = torch.ones(24, device=DEV), torch.zeros(24, device=DEV)
    y = instance_norm_relu(x, w, b)
    ref = F.relu(F.instance_norm(x, weight=w, bias=b, eps=1e-5))
    assert _rel(y, ref) < 1e-4
    assert abs(float(y.mean()) - 0.3989) < 2e-3                          # E[relu(N(0,1))] = 1/sqrt(2 pi)


def test_cpu_tensor_raises():
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        instance_norm_relu(torch.zeros(1, 2, 2, 2, 2), torch.ones(2), torch.zeros(2))


@pytest.mark.parametrize("shape", [(2, 24, 16, 16, 32), (1, 48, 9, 7, 13), (2, 96, 5, 6, 7), (1, 768, 5, 5, 8), (3, 384, 2, 3, 5), (2, 192, 1, 1, 3),
                                   (1, 24, 40, 41, 67)])
def test_channels_last_layout_matches_torch_and_keeps_the_layout(shape):
    """NDHWC kernels (include/instnorm.h, *_ndhwc): same numbers as the NCDHW path, output and gradient stay channels-last."""
    g = torch.Generator().manual_seed(sum(shape) + 1)
    x = (torch.randn(*shape, generator=g) * 2 + 3).to(DEV)
    w = (torch.rand(shape[1], generator=g) + 0.5).to(DEV)
    b = (torch.randn(shape[1], generator=g) * 0.3).to(DEV)
    dy = torch.randn(*shape, generator=g).to(DEV)
    xs = x.contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    ws, bs = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    from transoar_b200 import _lib
    n0 = _lib.lib().msda3d_launch_count()
    y = instance_norm_relu(xs, ws, bs, 1e-5)
    y.backward(dy)
    assert _lib.lib().msda3d_launch_count() - n0 == 7
    assert y.is_contiguous(memory_format=torch.channels_last_3d) and xs.grad.is_contiguous(memory_format=torch.channels_last_3d)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.relu(F.instance_norm(xr, weight=wr, bias=br, eps=1e-5))
    yr.backward(dy.double())
    assert _rel(y, yr) < 1e-5 and _rel(xs.grad, xr.grad) < 2e-4
    assert _rel(ws.grad, wr.grad) < 1e-4 and _rel(bs.grad, br.grad) < 1e-4


def test_channels_last_first_stage_statistics_at_full_size():
    x = (torch.randn(1, 24, 160, 160, 256, device=DEV) * 5 + 100).contiguous(memory_format=torch.channels_last_3d)
    w, b = torch.ones(24, device=DEV), torch.zeros(24, device=DEV)
    y = instance_norm_relu(x, w, b)
    ref = F.relu(F.instance_norm(x, weight=w, bias=b, eps=1e-5))
    assert _rel(y, ref) < 1e-4


@pytest.mark.parametrize("shape", [(2, 24, 16, 16, 32), (1, 48, 9, 7, 13), (2, 96, 5, 6, 7), (1, 768, 5, 5, 8), (1, 24, 40, 41, 67)])
def test_channels_last_bf16_storage(shape):
    """NDHWC kernels with bf16 storage (the encoder under the bf16 autocast route): fp32 arithmetic on bf16-rounded inputs, bf16 outputs."""
    g = torch.Generator().manual_seed(sum(shape) + 2)
    x = (torch.randn(*shape, generator=g) * 2 + 3).to(DEV).to(torch.bfloat16)
    w = (torch.rand(shape[1], generator=g) + 0.5).to(DEV)
    b = (torch.randn(shape[1], generator=g) * 0.3).to(DEV)
    dy = torch.randn(*shape, generator=g).to(DEV).to(torch.bfloat16)
    xs = x.contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    ws, bs = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    from transoar_b200 import _lib
    n0 = _lib.lib().msda3d_launch_count()
    y = instance_norm_relu(xs, ws, bs, 1e-5)
    y.backward(dy.contiguous(memory_format=torch.channels_last_3d))
    assert _lib.lib().msda3d_launch_count() - n0 == 7
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last_3d)
    assert xs.grad.dtype == torch.bfloat16 and xs.grad.is_contiguous(memory_format=torch.channels_last_3d)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.relu(F.instance_norm(xr, weight=wr, bias=br, eps=1e-5))
    yr.backward(dy.double())
    assert _rel(y, yr) < 1e-2 and _rel(xs.grad, xr.grad) < 2e-2
    assert _rel(ws.grad, wr.grad) < 1e-2 and _rel(bs.grad, br.grad) < 1e-2
