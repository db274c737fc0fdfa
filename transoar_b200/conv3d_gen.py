"""The AttnFPN backbone's remaining convolutions on this library's tcgen05 kernels -- no cuDNN call left on the path.

* ``conv3d_k3_gen(x, weight, bias, stride)`` = ``F.conv3d(x, weight, bias, stride, 1)`` for 3x3x3 kernels, stride 1 or 2, any channel
  counts divisible by 4 (include/conv3d_gen.h: implicit GEMM, TMA boxes per tap, stride 2 through parity-class tensor maps); input and
  weight gradients on the sibling kernels.  EncoderCnnBlock stages 1-5 (encoder_blocks.py:28-46) and Decoder._out (attn_fpn.py:65-74).
* ``conv3d_1x1(x, weight, bias)`` -- the FPN lateral convolutions (attn_fpn.py:60-62): in NDHWC memory a 1x1x1 convolution IS the
  GEMM [voxels, CI] x [CO, CI]^T of include/tc_gemm.h, bias in the epilogue.
* ``conv_transpose3d_k2s2(x, weight, bias, skip)`` -- the top-down ``ConvTranspose3d(kernel 2, stride 2)`` (attn_fpn.py:76-83) plus the
  lateral skip connection it is added to (attn_fpn.py:113): kernel == stride means every output voxel has exactly one source voxel, so
  it is the GEMM [voxels, CI] x [CI, 8 * CO] followed by a 2x2x2 pixel shuffle.

All fp32 channels-last, TF32 multiply / fp32 accumulate (what ``torch.backends.cudnn.allow_tf32`` = True, the reference's default, makes
cuDNN do).  CUDA only; raises on CPU tensors."""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .linear import colsum, gemm


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last_3d)


def _empty_cl(shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device, memory_format=torch.channels_last_3d)


def _common_ok(conv, x):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and torch.backends.cudnn.allow_tf32 and not torch.is_autocast_enabled()
            and conv.weight.dtype == torch.float32 and conv.groups == 1 and tuple(conv.dilation) == (1, 1, 1)
            and conv.in_channels % 4 == 0 and conv.out_channels % 4 == 0
            and x.is_contiguous(memory_format=torch.channels_last_3d))


def conv_gen_eligible(conv, x):
    """fp32 CUDA channels-last input, TF32 convolutions requested, 3x3x3 / padding 1 / stride 1 or 2, channel counts divisible by 4."""
    return (_common_ok(conv, x) and tuple(conv.kernel_size) == (3, 3, 3) and tuple(conv.padding) == (1, 1, 1)
            and tuple(conv.stride) in ((1, 1, 1), (2, 2, 2)) and conv.padding_mode == "zeros"
            and (conv.stride[0] == 1 or min(x.shape[2:]) >= 2))


def conv_1x1_eligible(conv, x):
    return (_common_ok(conv, x) and tuple(conv.kernel_size) == (1, 1, 1) and tuple(conv.padding) == (0, 0, 0) and tuple(conv.stride) == (1, 1, 1))


def conv_transpose_eligible(conv, x):
    return (_common_ok(conv, x) and tuple(conv.kernel_size) == (2, 2, 2) and tuple(conv.stride) == (2, 2, 2) and tuple(conv.padding) == (0, 0, 0)
            and tuple(conv.output_padding) == (0, 0, 0))


_FOLD_IDX = {}


def fold_stride2_weights(w_taps):
    """[27, CO, CI] tap-major weights -> w_fold [8, 8 * CIP, CO] of include/conv3d_gen.h::conv3d_gen_dgrad_s2_folded: block (delta, class) holds
    the transposed weights of the tap that connects dy[j + delta] with the class's dx[2 j + p] (per axis k = 1 / 2 / 0 for (p, delta) =
    (0, 0) / (1, 0) / (1, 1)), zeros where there is none."""
    _, co, ci = w_taps.shape
    cip = (ci + 31) // 32 * 32
    key = w_taps.device
    if key not in _FOLD_IDX:
        k_of = {(0, 0): 1, (1, 0): 2, (1, 1): 0}
        idx = []
        for delta in range(8):
            for cls in range(8):
                ks = [k_of.get(((cls >> s) & 1, (delta >> s) & 1)) for s in (2, 1, 0)]        # (d, h, w)
                idx.append(27 if None in ks else (ks[0] * 3 + ks[1]) * 3 + ks[2])
        _FOLD_IDX[key] = torch.tensor(idx, dtype=torch.long, device=w_taps.device)
    ext = torch.cat((w_taps, w_taps.new_zeros(1, co, ci)))
    sel = ext.index_select(0, _FOLD_IDX[key]).view(8, 8, co, ci).transpose(2, 3)                # [delta, class, ci, co]
    out = w_taps.new_zeros(8, 8, cip, co)
    out[:, :, :ci] = sel
    return out.view(8, 8 * cip, co)


class Conv3dGenFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        if not x.is_cuda:
            raise RuntimeError("conv3d_k3_gen: Not implemented on the CPU")
        x = _cl(x)
        N, ci, D, H, W = x.shape
        co = weight.shape[0]
        w = weight.permute(2, 3, 4, 0, 1).reshape(27, co, ci).contiguous()          # tap-major [27][CO][CI], shared by forward and input gradient
        od, oh, ow = ((v + stride - 1) // stride for v in (D, H, W))
        y = _empty_cl((N, co, od, oh, ow), x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().conv3d_gen_forward(_stream(), _p(x), _p(w), _p(None if bias is None else bias.contiguous()), N, D, H, W, ci, co, stride, _p(y))
        _lib.check(rc, "conv3d_gen_forward")
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.has_bias = stride, bias is not None
        ctx.w_cl = weight.is_contiguous(memory_format=torch.channels_last_3d) and not weight.is_contiguous()
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        N, ci, D, H, W = x.shape
        co = w.shape[1]
        dy = _cl(dy)
        dx = dw = db = None
        lib = _lib.lib()
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                dx = _empty_cl(x.shape, x.device)
                if ctx.stride == 2 and ci <= 64:                                # narrow stage entries: parity classes folded into N
                    _lib.check(lib.conv3d_gen_dgrad_s2_folded(_stream(), _p(dy), _p(fold_stride2_weights(w)), N, D, H, W, ci, co, _p(dx)),
                               "conv3d_gen_dgrad_s2_folded")
                else:
                    _lib.check(lib.conv3d_gen_dgrad(_stream(), _p(dy), _p(w), N, D, H, W, ci, co, ctx.stride, _p(dx)), "conv3d_gen_dgrad")
            if ctx.needs_input_grad[1]:
                dw = _empty_cl((co, ci, 3, 3, 3), x.device)                      # the kernel writes channels-last weight memory [CO][27][CI]
                _lib.check(lib.conv3d_gen_wgrad(_stream(), _p(x), _p(dy), N, D, H, W, ci, co, ctx.stride, _p(dw)), "conv3d_gen_wgrad")
                if not ctx.w_cl:
                    dw = dw.contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy.permute(0, 2, 3, 4, 1).reshape(-1, co))
        return dx, dw, db, None


def conv3d_k3_gen(x, weight, bias=None, stride=1):
    return Conv3dGenFunction.apply(x, weight, bias, int(stride))


class Conv1x1Function(Function):
    """y[v, co] = sum_ci x[v, ci] w[co, ci] + b[co] over the NDHWC rows of the volume."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        if not x.is_cuda:
            raise RuntimeError("conv3d_1x1: Not implemented on the CPU")
        x = _cl(x)
        N, ci, D, H, W = x.shape
        co = weight.shape[0]
        w2 = weight.reshape(co, ci).contiguous()
        M = N * D * H * W
        y = _empty_cl((N, co, D, H, W), x.device)
        gemm(x, 0, ci, w2, 0, ci, y.permute(0, 2, 3, 4, 1).reshape(M, co), M, co, ci, bias=None if bias is None else bias.contiguous())
        ctx.save_for_backward(x, w2)
        ctx.has_bias, ctx.wshape = bias is not None, weight.shape
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w2 = ctx.saved_tensors
        N, ci, D, H, W = x.shape
        co = w2.shape[0]
        M = N * D * H * W
        dy = _cl(dy)
        dy2 = dy.permute(0, 2, 3, 4, 1).reshape(M, co)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _empty_cl(x.shape, x.device)
            gemm(dy2, 0, co, w2, 1, ci, dx.permute(0, 2, 3, 4, 1).reshape(M, ci), M, ci, co)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(co, ci, dtype=torch.float32, device=x.device)
            gemm(dy2, 1, co, x, 1, ci, dw, co, ci, M, accumulate=True, split_k=0)
            dw = dw.reshape(ctx.wshape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        return dx, dw, db


def conv3d_1x1(x, weight, bias=None):
    return Conv1x1Function.apply(x, weight, bias)


def _split8(t, D, H, W):
    """[N, C, 2D, 2H, 2W] -> the view [N, C, D, 2, H, 2, W, 2] (index 2d + a -> (d, a)); no copy for any strides."""
    return t.view(t.shape[0], t.shape[1], D, 2, H, 2, W, 2)


class ConvTransposeK2S2Function(Function):
    """out = ConvTranspose3d(k=2, s=2)(x) + bias (+ skip).  weight [CI, CO, 2, 2, 2] (torch's transposed-convolution layout)."""

    @staticmethod
    def forward(ctx, x, weight, bias, skip):
        if not x.is_cuda:
            raise RuntimeError("conv_transpose3d_k2s2: Not implemented on the CPU")
        x = _cl(x)
        N, ci, D, H, W = x.shape
        co = weight.shape[1]
        wt = weight.permute(2, 3, 4, 1, 0).reshape(8 * co, ci).contiguous()          # rows (a, b, c, co), K-major in ci
        M = N * D * H * W
        g = torch.empty(M, 8 * co, dtype=torch.float32, device=x.device)
        gemm(x, 0, ci, wt, 0, ci, g, M, 8 * co, ci, bias=None if bias is None else bias.repeat(8).contiguous())
        # pixel shuffle out[n, :, 2d+a, 2h+b, 2w+c] = g[(n,d,h,w), (a,b,c), :] as ONE strided pass that also adds the skip connection
        up = g.view(N, D, H, W, 2, 2, 2, co).permute(0, 7, 1, 4, 2, 5, 3, 6)           # [N, co, D, a, H, b, W, c]
        out = _empty_cl((N, co, 2 * D, 2 * H, 2 * W), x.device)
        if skip is None:
            _split8(out, D, H, W).copy_(up)
        else:
            torch.add(up, _split8(skip, D, H, W), out=_split8(out, D, H, W))
        ctx.save_for_backward(x, wt)
        ctx.has_bias, ctx.has_skip, ctx.wshape = bias is not None, skip is not None, weight.shape
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, wt = ctx.saved_tensors
        N, ci, D, H, W = x.shape
        co = wt.shape[0] // 8
        M = N * D * H * W
        dout = _cl(dout)
        # inverse shuffle: dg[(n,d,h,w), (a,b,c), co] = dout[n, co, 2d+a, 2h+b, 2w+c]
        dg = _split8(dout, D, H, W).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(M, 8 * co)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _empty_cl(x.shape, x.device)
            gemm(dg, 0, 8 * co, wt, 1, ci, dx.permute(0, 2, 3, 4, 1).reshape(M, ci), M, ci, 8 * co)
        if ctx.needs_input_grad[1]:
            dwt = torch.zeros(8 * co, ci, dtype=torch.float32, device=x.device)
            gemm(dg, 1, 8 * co, x, 1, ci, dwt, 8 * co, ci, M, accumulate=True, split_k=0)
            dw = dwt.view(2, 2, 2, co, ci).permute(4, 3, 0, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dout.permute(0, 2, 3, 4, 1).reshape(-1, co))
        return dx, dw, db, (dout if ctx.has_skip and ctx.needs_input_grad[3] else None)


def conv_transpose3d_k2s2(x, weight, bias=None, skip=None):
    return ConvTransposeK2S2Function.apply(x, weight, bias, skip)
