"""Fused InstanceNorm3d(affine) + ReLU on the sm_100a kernels (include/instnorm.h).

Two memory layouts: NCDHW (fp32 / bf16) and channels-last NDHWC (fp32; ``torch.channels_last_3d``), picked from the input's
strides; the output has the input's layout, so a channels-last encoder never transposes.

Stands in for the ``nn.InstanceNorm3d(affine=True) -> nn.ReLU(inplace=True)`` pairs of the reference's ``EncoderCnnBlock``
(transoar/models/backbones/encoder_blocks.py:28-46).  fp32 or bf16 activations, fp32 statistics and parameters."""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _channels_last(x):
    """True when x is a 5-D fp32 / bf16 tensor stored NDHWC (and not also NCDHW-contiguous) with a channel count the NDHWC kernels take."""
    C = x.shape[1] if x.dim() == 5 else 0
    return (x.dim() == 5 and x.dtype in (torch.float32, torch.bfloat16) and C % 4 == 0 and C > 0 and 192 % (C // 4) == 0
            and x.is_contiguous(memory_format=torch.channels_last_3d) and not x.is_contiguous())


class InstanceNormReLUFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        if not x.is_cuda:
            raise RuntimeError("instance_norm_relu: Not implemented on the CPU")
        if x.dtype not in _DT:
            x = x.float()
        cl = _channels_last(x)
        if not cl:
            x = x.contiguous()
        B, C = x.shape[:2]
        V = x[0, 0].numel()
        w, b = weight.float().contiguous(), bias.float().contiguous()
        y = torch.empty_like(x)                                  # keeps the input's strides (NDHWC stays NDHWC)
        mean = torch.empty(B * C, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        ctx.channels_last = cl
        if cl:
            ws = torch.empty(_lib.lib().instnorm_ndhwc_workspace_floats(B, C, V), dtype=torch.float32, device=x.device)
            fwd = _lib.lib().instnorm_relu_forward_ndhwc_bf16 if x.dtype == torch.bfloat16 else _lib.lib().instnorm_relu_forward_ndhwc
            with torch.cuda.device(x.device):
                rc = fwd(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(x), _p(w), _p(b),
                         B, C, V, float(eps), _p(y), _p(mean), _p(rstd), _p(ws))
            _lib.check(rc, "instnorm_relu_forward_ndhwc")
            ctx.save_for_backward(x, b, w, mean, rstd)            # the NDHWC backward recomputes the ReLU mask from x: y is not kept
            ctx.param_dtypes = (weight.dtype, bias.dtype)
            return y
        ws = torch.empty(_lib.lib().instnorm_workspace_floats(_DT[x.dtype], B, C, V), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().instnorm_relu_forward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _DT[x.dtype], _p(x), _p(w), _p(b),
                                                  B, C, V, float(eps), _p(y), _p(mean), _p(rstd), _p(ws))
        _lib.check(rc, "instnorm_relu_forward")
        ctx.save_for_backward(x, y, w, mean, rstd)
        ctx.param_dtypes = (weight.dtype, bias.dtype)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, y, w, mean, rstd = ctx.saved_tensors                  # channels-last: the second entry is beta, not y (mask recomputed from x)
        B, C = x.shape[:2]
        V = x[0, 0].numel()
        dy = dy.to(x.dtype)
        dy = dy.contiguous(memory_format=torch.channels_last_3d) if ctx.channels_last else dy.contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty(C, dtype=torch.float32, device=x.device)
        db = torch.empty_like(dw)
        if ctx.channels_last:
            ws = torch.empty(_lib.lib().instnorm_ndhwc_workspace_floats(B, C, V), dtype=torch.float32, device=x.device)
            bwd = _lib.lib().instnorm_relu_backward_ndhwc_bf16 if x.dtype == torch.bfloat16 else _lib.lib().instnorm_relu_backward_ndhwc
            with torch.cuda.device(x.device):
                rc = bwd(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(dy), _p(x), _p(w),
                         _p(y), _p(mean), _p(rstd), B, C, V, _p(dx), _p(dw), _p(db), _p(ws))
            _lib.check(rc, "instnorm_relu_backward_ndhwc")
            return dx, dw.to(ctx.param_dtypes[0]), db.to(ctx.param_dtypes[1]), None
        ws = torch.empty(_lib.lib().instnorm_workspace_floats(_DT[x.dtype], B, C, V), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().instnorm_relu_backward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _DT[x.dtype], _p(dy), _p(x), _p(y),
                                                   _p(w), _p(mean), _p(rstd), B, C, V, _p(dx), _p(dw), _p(db), _p(ws))
        _lib.check(rc, "instnorm_relu_backward")
        return dx, dw.to(ctx.param_dtypes[0]), db.to(ctx.param_dtypes[1]), None


def instance_norm_relu(x, weight, bias, eps=1e-5):
    """relu(instance_norm(x, weight, bias, eps)) for [N, C, *spatial] tensors."""
    return InstanceNormReLUFunction.apply(x, weight, bias, eps)
