"""The fused Swin3D window attention (include/win_attn.h) against the composition the reference runs (encoder_blocks.py:259-285:
q k^T * scale + relative position bias (+ shift mask) -> softmax -> @ v) in fp64 on the same inputs: output and all gradients (qkv and
the bias, i.e. through it the relative_position_bias_table), with and without the shifted-window mask, full and truncated windows."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(qkv, bias, mask, scale):
    Bw, n, _, H, hd = qkv.shape
    q, k, v = qkv.permute(2, 0, 3, 1, 4)                                                   # [Bw, H, n, hd]
    attn = (q * scale) @ k.transpose(-2, -1) + bias[None]
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(Bw // nW, nW, H, n, n) + mask[None, :, None]).view(Bw, H, n, n)
    return (attn.softmax(-1) @ v).transpose(1, 2).reshape(Bw, n, H * hd)


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("Bw,n,H,nW", [(8, 125, 3, 0), (12, 125, 6, 4), (6, 100, 12, 3), (4, 27, 24, 2), (2, 128, 3, 0), (5, 1, 3, 0)])
def test_forward_and_gradients_match_fp64(Bw, n, H, nW):
    from transoar_b200 import _lib
    from transoar_b200.swin import WindowAttentionFunction
    g = torch.Generator().manual_seed(Bw * 1000 + n + H)
    qkv = torch.randn(Bw, n, 3, H, 16, generator=g).to(DEV).requires_grad_(True)
    bias = (torch.randn(H, n, n, generator=g) * 0.5).to(DEV).requires_grad_(True)
    mask = None
    if nW:
        labels = torch.randint(0, 3, (nW, n), generator=g)
        mask = torch.zeros(nW, n, n).masked_fill(labels[:, :, None] != labels[:, None, :], -100.0).to(DEV)
    dout = torch.randn(Bw, n, H * 16, generator=g).to(DEV)
    scale = 16 ** -0.5
    n0 = _lib.lib().msda3d_launch_count()
    out = WindowAttentionFunction.apply(qkv, bias, mask, scale)
    out.backward(dout)
    assert _lib.lib().msda3d_launch_count() == n0 + 2
    qd, bd = qkv.detach().double().requires_grad_(True), bias.detach().double().requires_grad_(True)
    want = _reference(qd, bd, None if mask is None else mask.double(), scale)
    want.backward(dout.double())
    assert _rel(out.detach(), want.detach()) < 2e-5
    assert _rel(qkv.grad, qd.grad) < 1e-4 and _rel(bias.grad, bd.grad) < 1e-4


def test_module_uses_the_kernel_and_matches_the_sdpa_route():
    """WindowAttention3D end to end (strict fp32 Linear layers): fused kernel vs the library SDPA route of the same module, incl. the
    gradient that reaches relative_position_bias_table through the gathered bias."""
    from transoar_b200 import _lib, swin
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        m = swin.WindowAttention3D(48, (5, 5, 5), 3, True, None, 0.0, 0.0).to(DEV)
        with torch.no_grad():
            m.relative_position_bias_table.normal_(std=0.5)
        x = torch.randn(16, 125, 48, device=DEV, requires_grad=True)
        labels = torch.randint(0, 4, (8, 125))
        mask = torch.zeros(8, 125, 125).masked_fill(labels[:, :, None] != labels[:, None, :], -100.0).to(DEV)
        g = torch.randn(16, 125, 48, device=DEV)
        n0 = _lib.lib().msda3d_launch_count()
        y = m(x, mask)
        y.backward(g)
        assert _lib.lib().msda3d_launch_count() == n0 + 2
        got = (y.detach().clone(), x.grad.clone(), m.relative_position_bias_table.grad.clone(), m.qkv.weight.grad.clone())
        x.grad = None
        m.zero_grad()
        orig = _lib.lib().win_attn_supported
        try:
            swin._lib.lib().win_attn_supported = lambda *_: 0                             # force the library route
            y2 = m(x, mask)
        finally:
            swin._lib.lib().win_attn_supported = orig
        y2.backward(g)
        want = (y2.detach(), x.grad, m.relative_position_bias_table.grad, m.qkv.weight.grad)
        for a, b in zip(got, want):
            assert float((a - b).abs().max()) <= 2e-4 * float(b.abs().max())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
