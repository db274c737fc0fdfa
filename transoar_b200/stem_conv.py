"""The encoder's first convolution (1 -> start_channels, 3x3x3, stride 1, padding 1, no bias) as a direct sm_100a stencil
(include/stem_conv.h): fp32 FMA arithmetic, channels-last output, weight gradient on the same library.

``stem_conv3d(x, weight)`` equals ``F.conv3d(x, weight, None, 1, 1)`` for ``x`` [N,1,D,H,W] and returns a channels-last
(``torch.channels_last_3d``) tensor.  ``stem_eligible`` says when the kernel applies; the input gradient is not implemented
(the input is the CT volume), so an ``x`` that requires grad is not eligible and takes the library convolution."""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def stem_eligible(conv, x):
    return (x.is_cuda and x.dtype == torch.float32 and not x.requires_grad and x.dim() == 5 and conv.in_channels == 1
            and conv.out_channels in (16, 24, 32) and tuple(conv.kernel_size) == (3, 3, 3) and tuple(conv.stride) == (1, 1, 1)
            and tuple(conv.padding) == (1, 1, 1) and tuple(conv.dilation) == (1, 1, 1) and conv.groups == 1 and conv.bias is None
            and conv.weight.dtype == torch.float32 and not torch.is_autocast_enabled())


class StemConvFunction(Function):
    @staticmethod
    def forward(ctx, x, weight):
        if not x.is_cuda:
            raise RuntimeError("stem_conv3d: Not implemented on the CPU")
        N, _, D, H, W = x.shape
        CO = weight.shape[0]
        x = x.contiguous()
        w = weight.contiguous()                                   # [CO,1,3,3,3]: one input channel, so any memory format is this
        y = torch.empty((N, CO, D, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last_3d)
        with torch.cuda.device(x.device):
            rc = _lib.lib().stem_conv3d_forward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(x), _p(w), N, D, H, W, CO, _p(y))
        _lib.check(rc, "stem_conv3d_forward")
        ctx.save_for_backward(x)
        ctx.wshape = weight.shape
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        N, _, D, H, W = x.shape
        CO = ctx.wshape[0]
        dy = dy.float().contiguous(memory_format=torch.channels_last_3d)
        dw = torch.empty(ctx.wshape, dtype=torch.float32, device=x.device)
        ws = torch.empty(_lib.lib().stem_conv3d_workspace_floats(CO), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().stem_conv3d_wgrad(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(dy), _p(x), N, D, H, W, CO, _p(dw), _p(ws))
        _lib.check(rc, "stem_conv3d_wgrad")
        return None, dw


def stem_conv3d(x, weight):
    return StemConvFunction.apply(x, weight)
