"""GPU parity of the encoder's first convolution (include/stem_conv.h) against torch's conv3d in fp64 on the same inputs, and
its use inside the channels-last backbone."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(2, 8, 9, 37), (1, 5, 6, 300), (1, 3, 3, 3), (2, 16, 16, 256), (1, 1, 1, 1), (1, 4, 2, 513)])
@pytest.mark.parametrize("co", [16, 24, 32])
def test_forward_and_weight_gradient_match_conv3d(shape, co):
    from transoar_b200.stem_conv import stem_conv3d
    N, D, H, W = shape
    g = torch.Generator().manual_seed(N + D + H + W + co)
    x = torch.rand(N, 1, D, H, W, generator=g).to(DEV)
    w = (torch.randn(co, 1, 3, 3, 3, generator=g) * 0.2).to(DEV).requires_grad_(True)
    dy = torch.randn(N, co, D, H, W, generator=g).to(DEV)
    y = stem_conv3d(x, w)
    y.backward(dy)
    assert y.shape == (N, co, D, H, W) and y.is_contiguous(memory_format=torch.channels_last_3d)
    wd = w.detach().double().requires_grad_(True)
    yd = F.conv3d(x.double(), wd, None, 1, 1)
    yd.backward(dy.double())
    assert _rel(y, yd) < 1e-5
    assert _rel(w.grad, wd.grad) < 1e-4


def test_full_size_row_sums():
    """160x160x256 x 24 channels (1.26 GB of output per 2 volumes): with an all-ones kernel every output channel is the 3x3x3 box
    sum of the volume; checked against avg_pool3d (a size-independent identity)."""
    from transoar_b200.stem_conv import stem_conv3d
    x = torch.rand(1, 1, 160, 160, 256, device=DEV)
    w = torch.ones(24, 1, 3, 3, 3, device=DEV)
    y = stem_conv3d(x, w)
    box = F.avg_pool3d(x, 3, 1, 1, count_include_pad=True) * 27
    assert _rel(y[:, 0:1], box) < 1e-5 and torch.equal(y[:, 0], y[:, 23])


def test_backbone_uses_the_stem_kernel_when_channels_last():
    from transoar_b200 import _lib
    from transoar_b200.attn_fpn import EncoderCnnBlock
    torch.manual_seed(0)
    blk = EncoderCnnBlock(1, 24, 3, 1).to(DEV)
    x = torch.rand(1, 1, 16, 16, 32, device=DEV)
    with torch.backends.cudnn.flags(allow_tf32=False):
        ref = blk(x)
        (ref.square().sum()).backward()
        gref = [p.grad.clone() for p in blk.parameters()]
        blk.zero_grad()
        blk_cl = blk.to(memory_format=torch.channels_last_3d)
        n0 = _lib.lib().msda3d_launch_count()
        out = blk_cl(x)
        (out.square().sum()).backward()
        assert _lib.lib().msda3d_launch_count() - n0 == 1 + 2 + 2 * 7       # stem fwd + wgrad (2 kernels) + two InstanceNorm pairs
    assert _rel(out, ref) < 1e-4
    for p, g in zip(blk_cl.parameters(), gref):
        assert _rel(p.grad, g) < 2e-3
    # an input that requires grad is not eligible (no input gradient in the stem kernel): the library convolution runs instead
    xg = x.clone().requires_grad_(True)
    blk_cl(xg).sum().backward()
    assert xg.grad is not None
