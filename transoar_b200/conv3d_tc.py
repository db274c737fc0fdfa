"""3x3x3 / stride 1 / padding 1 convolution of a channels-last fp32 volume on the tcgen05 implicit-GEMM kernel (include/conv3d_tc.h).

``conv3d_k3(x, weight)`` equals ``F.conv3d(x, weight, None, 1, 1)`` (TF32 multiply, fp32 accumulate) for channels-last ``x``
[N, CI, D, H, W] and returns a channels-last tensor.  The gradient with respect to the input runs on the same kernel (flipped taps,
transposed channels); the weight gradient on its MN-major sibling (``conv3d_tc_k3_wgrad``: the three kw taps as overlapping slabs of one MMA operand).  Used for the second convolution of the encoder's
first stage (24 -> 24 channels at full resolution), where cuDNN's input-gradient kernel is the slow one."""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def conv_tc_eligible(conv, x):
    """fp32 CUDA, channels-last input (NDHWC in memory), TF32 convolutions requested, 3x3x3 / 1 / 1 without bias, supported channel counts."""
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and torch.backends.cudnn.allow_tf32 and not torch.is_autocast_enabled()
            and tuple(conv.kernel_size) == (3, 3, 3) and tuple(conv.stride) == (1, 1, 1) and tuple(conv.padding) == (1, 1, 1)
            and tuple(conv.dilation) == (1, 1, 1) and conv.groups == 1 and conv.bias is None and conv.weight.dtype == torch.float32
            and conv.in_channels == conv.out_channels                     # the input gradient swaps the channel roles: both must qualify
            and bool(_lib.lib().conv3d_tc_supported(conv.in_channels, conv.out_channels))
            and x.is_contiguous(memory_format=torch.channels_last_3d) and not x.is_contiguous())


def _run(x, w_taps, co):
    N, ci, D, H, W = x.shape
    y = torch.empty((N, co, D, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last_3d)
    with torch.cuda.device(x.device):
        rc = _lib.lib().conv3d_tc_k3_forward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(x), _p(w_taps), N, D, H, W, ci, co, _p(y))
    _lib.check(rc, "conv3d_tc_k3_forward")
    return y


class Conv3dK3Function(Function):
    @staticmethod
    def forward(ctx, x, weight):
        if not x.is_cuda:
            raise RuntimeError("conv3d_k3: Not implemented on the CPU")
        x = x.contiguous(memory_format=torch.channels_last_3d)
        co, ci = weight.shape[:2]
        w_taps = weight.permute(2, 3, 4, 0, 1).reshape(27, co, ci).contiguous()                       # [tap][co][ci]
        ctx.save_for_backward(x, weight)
        return _run(x, w_taps, co)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        co, ci = weight.shape[:2]
        dy = dy.contiguous(memory_format=torch.channels_last_3d)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            w_back = weight.flip(2, 3, 4).permute(2, 3, 4, 1, 0).reshape(27, ci, co).contiguous()    # [tap'][ci][co], taps mirrored
            dx = _run(dy, w_back, ci)
        if ctx.needs_input_grad[1]:
            N, _, D, H, W = x.shape
            dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)                      # contiguous [CO, CI, 3, 3, 3]
            ws = torch.empty(_lib.lib().conv3d_tc_wgrad_workspace_floats(), dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                rc = _lib.lib().conv3d_tc_k3_wgrad(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(x), _p(dy), N, D, H, W, ci, co,
                                                   _p(dw), _p(ws))
            _lib.check(rc, "conv3d_tc_k3_wgrad")
            dw = dw.contiguous(memory_format=torch.channels_last_3d) if weight.is_contiguous(memory_format=torch.channels_last_3d) else dw
        return dx, dw


def conv3d_k3(x, weight):
    return Conv3dK3Function.apply(x, weight)
