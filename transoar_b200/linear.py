"""Linear layers on the TF32 tcgen05 GEMM (include/tc_gemm.h) -- the dense contractions of the hot path.

``linear(x, weight, bias, relu)`` computes ``relu?(x @ weight.T + bias)`` like ``torch.nn.functional.linear`` (what every
``nn.Linear`` of the reference's MSDeformAttn / DefAttnLayer / FocusedAttn / FocusedDecoderLayer calls), with bias and ReLU
fused into the GEMM epilogue, and its gradients as two more launches of the same kernel (grad_input = grad_output W read with W
in MN-major form, grad_weight = grad_output^T x with both operands MN-major and split-K) -- no operand is transposed in HBM.
fp32 tensors, TF32 multiply, fp32 accumulate: what torch 1.10 (the reference's pin) does for fp32 matmuls on Ampere+.
``TCLinear`` is an ``nn.Linear`` with the same parameter names, so reference checkpoints load.  CUDA only; raises on CPU tensors."""
import ctypes

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(A, a_mn, lda, B, b_mn, ldb, D, M, N, R, bias=None, relu=False, accumulate=False, split_k=1):
    """D[M,N] (+)= sum_r A(m,r) B(n,r) (+bias)(relu) on the current stream; see include/tc_gemm.h for the operand forms."""
    if not (A.is_cuda and B.is_cuda and D.is_cuda):
        raise RuntimeError("tc_gemm: Not implemented on the CPU")
    if A.dtype != torch.float32 or B.dtype != torch.float32 or D.dtype != torch.float32:
        raise RuntimeError("tc_gemm: fp32 tensors only")
    with torch.cuda.device(D.device):
        rc = _lib.lib().tc_gemm_tf32(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(A), int(a_mn), lda, _p(B), int(b_mn), ldb,
                                     _p(D), D.stride(0), _p(bias), M, N, R, int(relu), int(accumulate), split_k)
    _lib.check(rc, "tc_gemm_tf32")
    return D


class LinearFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        if not x.is_cuda:
            raise RuntimeError("tc linear: Not implemented on the CPU")
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.reshape(-1, K).float().contiguous()
        w = weight.float().contiguous()
        M = x2.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        gemm(x2, 0, K, w, 0, K, y, M, N, K, bias=None if bias is None else bias.float().contiguous(), relu=relu)
        ctx.save_for_backward(x2, w, y if relu else None)
        ctx.has_bias, ctx.in_shape = bias is not None, x.shape
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy2 = dy.reshape(M, N).float()
        dy2 = dy2 * (y > 0) if y is not None else dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
            gemm(dy2, 0, N, w, 1, K, dx, M, K, N)                                       # dX = dY W        (W as B(k, n) = W[n, k]: MN-major)
            dx = dx.reshape(ctx.in_shape)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(N, K, dtype=torch.float32, device=dy.device)
            gemm(dy2, 1, N, x2, 1, K, dw, N, K, M, accumulate=True, split_k=0)           # dW = dY^T X      (both MN-major, split-K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db, None


def linear(x, weight, bias=None, relu=False):
    return LinearFunction.apply(x, weight, bias, relu)


def tc_eligible(x, weight):
    """The tcgen05 kernel multiplies in TF32, so it runs exactly where torch itself would use TF32 tensor cores for an fp32
    Linear: CUDA fp32 tensors with ``torch.backends.cuda.matmul.allow_tf32`` on (the default of the reference's torch 1.10;
    off by default in torch >= 1.12).  Feature counts must be multiples of 4 (TMA strides are multiples of 16 bytes)."""
    return (x.is_cuda and torch.backends.cuda.matmul.allow_tf32 and x.dtype == torch.float32 and weight.dtype == torch.float32
            and not torch.is_autocast_enabled() and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0)


class TCLinear(nn.Linear):
    """``nn.Linear`` (same parameters, same state_dict) whose forward / gradient GEMMs run on the tcgen05 kernel whenever TF32
    is the requested precision (see ``tc_eligible``); with strict fp32 requested, under autocast or for the 1- / 6-wide heads
    it is the plain library GEMM, exactly as in the reference.  ``relu=True`` fuses the activation into the GEMM epilogue."""

    def forward(self, x, relu=False):
        if tc_eligible(x, self.weight):
            return linear(x, self.weight, self.bias, relu)
        y = F.linear(x, self.weight, self.bias)
        return F.relu(y) if relu else y
