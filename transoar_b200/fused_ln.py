"""y = LayerNorm(a + dropout(b)) as one sm_100a kernel each way (include/fused_ln.h).

Stands in for the ``norm(x + dropout(branch))`` triples of the reference's post-norm blocks (DefAttnLayer,
decoder_blocks.py:163-177; FocusedDecoderLayer, focused_decoder.py:166-189).  The dropout mask is a counter-based hash of
(seed, element index): nothing is stored for it, the backward re-evaluates it.  The seed of every call is drawn from torch's CPU
generator, so ``torch.manual_seed`` makes runs reproducible; the keep pattern is of course not the one ATen's Philox stream
would produce (same distribution, different stream).  In eval mode / p = 0 the result is exactly LayerNorm(a + b)."""
import ctypes

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def fused_eligible(a, b, norm):
    C = a.shape[-1]
    # fp32 everywhere, or -- the bf16 route (autocast) -- an fp32 residual stream with a bf16 branch
    return (a.is_cuda and a.dtype == torch.float32 and b.dtype in (torch.float32, torch.bfloat16) and a.shape == b.shape and C % 4 == 0 and C <= 1024
            and norm.elementwise_affine and norm.bias is not None and norm.weight.dtype == torch.float32
            and tuple(norm.normalized_shape) == (C,) and (not torch.is_autocast_enabled() or torch.get_autocast_dtype("cuda") == torch.bfloat16))


class AddDropoutLayerNormFunction(Function):
    @staticmethod
    def forward(ctx, a, b, weight, bias, eps, p, seed):
        if not a.is_cuda:
            raise RuntimeError("add_dropout_layer_norm: Not implemented on the CPU")
        C = a.shape[-1]
        a2, b2 = a.reshape(-1, C).contiguous(), b.reshape(-1, C).contiguous()
        rows = a2.shape[0]
        w, bi = weight.contiguous(), bias.contiguous()
        z, y = torch.empty_like(a2), torch.empty_like(a2)
        mean = torch.empty(rows, dtype=torch.float32, device=a.device)
        rstd = torch.empty_like(mean)
        ctx.bf16_branch = b2.dtype == torch.bfloat16
        fwd = _lib.lib().fused_ln_forward_bf16b if ctx.bf16_branch else _lib.lib().fused_ln_forward
        with torch.cuda.device(a.device):
            rc = fwd(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(a2), _p(b2), _p(w), _p(bi), rows, C,
                     float(eps), float(p), seed, _p(z), _p(y), _p(mean), _p(rstd))
        _lib.check(rc, "fused_ln_forward")
        ctx.save_for_backward(z, w, mean, rstd)
        ctx.p, ctx.seed, ctx.shape = float(p), seed, a.shape
        return y.reshape(a.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        z, w, mean, rstd = ctx.saved_tensors
        rows, C = z.shape
        dy2 = dy.reshape(rows, C).contiguous()
        da = torch.empty_like(z)
        if ctx.bf16_branch:
            db = torch.empty(z.shape, dtype=torch.bfloat16, device=z.device)
        else:
            db = torch.empty_like(z) if ctx.p > 0 else None
        dw, dbias = torch.empty(C, dtype=torch.float32, device=z.device), torch.empty(C, dtype=torch.float32, device=z.device)
        ws = torch.empty(_lib.lib().fused_ln_workspace_floats(C), dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            bwd = _lib.lib().fused_ln_backward_bf16b if ctx.bf16_branch else _lib.lib().fused_ln_backward
            rc = bwd(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(dy2), _p(z), _p(w), _p(mean), _p(rstd),
                     rows, C, ctx.p, ctx.seed, _p(da), _p(db), _p(dw), _p(dbias), _p(ws))
        _lib.check(rc, "fused_ln_backward")
        da = da.reshape(ctx.shape)
        return da, (db.reshape(ctx.shape) if db is not None else da), dw, dbias, None, None, None


class LayerNormFunction(Function):
    """Plain ``LayerNorm(x)`` on the same kernels (branch pointer NULL): one warp per row, which is what small rows need -- ATen's
    layer-norm kernel spends a CTA per row and takes 1.2 ms for the 442 k x 48 token matrix of the first Swin stage (85 MB)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        if not x.is_cuda:
            raise RuntimeError("layer_norm: Not implemented on the CPU")
        C = x.shape[-1]
        x2 = x.reshape(-1, C).float().contiguous()
        rows = x2.shape[0]
        w, bi = weight.float().contiguous(), bias.float().contiguous()
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        with torch.cuda.device(x.device):
            rc = _lib.lib().fused_ln_forward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(x2), None, _p(w), _p(bi), rows, C,
                                             float(eps), 0.0, 0, None, _p(y), _p(mean), _p(rstd))
        _lib.check(rc, "fused_ln_forward")
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.shape, ctx.in_dtype = x.shape, x.dtype
        return y.reshape(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        rows, C = x2.shape
        dy2 = dy.reshape(rows, C).float().contiguous()
        dx = torch.empty_like(x2)
        dw, dbias = torch.empty(C, dtype=torch.float32, device=x2.device), torch.empty(C, dtype=torch.float32, device=x2.device)
        ws = torch.empty(_lib.lib().fused_ln_workspace_floats(C), dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            rc = _lib.lib().fused_ln_backward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(dy2), _p(x2), _p(w), _p(mean), _p(rstd),
                                              rows, C, 0.0, 0, _p(dx), None, _p(dw), _p(dbias), _p(ws))
        _lib.check(rc, "fused_ln_backward")
        return dx.reshape(ctx.shape).to(ctx.in_dtype), dw, dbias, None


def layer_norm(x, norm):
    """``norm(x)`` for an ``nn.LayerNorm`` over the last dimension: the warp-per-row kernel for CUDA tensors (fp32 result, as ATen's
    layer_norm gives under autocast), the module itself otherwise."""
    C = x.shape[-1]
    if not (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and C % 4 == 0 and C <= 1024 and norm.elementwise_affine
            and norm.bias is not None and tuple(norm.normalized_shape) == (C,)):
        return norm(x)
    return LayerNormFunction.apply(x, norm.weight, norm.bias, norm.eps)


def add_dropout_layer_norm(a, b, norm, p, training, seed=None):
    """``norm(a + dropout(b, p, training))`` for an ``nn.LayerNorm`` over the last dimension.  Fused kernel for fp32 CUDA tensors
    (C % 4 == 0, C <= 1024); the plain composition otherwise (CPU tensors, autocast, other shapes)."""
    if not fused_eligible(a, b, norm):
        return norm(a + F.dropout(b, p, training))
    p = float(p) if training else 0.0
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
    return AddDropoutLayerNormFunction.apply(a, b, norm.weight, norm.bias, norm.eps, p, seed)
