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


def gemm(A, a_mn, lda, B, b_mn, ldb, D, M, N, R, bias=None, relu=False, accumulate=False, split_k=1, gate=None, gate_scale=1.0,
         p_drop=0.0, seed=0):
    """D[M,N] (+)= sum_r A(m,r) B(n,r) (+bias)(relu)(dropout)(gate) on the current stream; see include/tc_gemm.h."""
    if not (A.is_cuda and B.is_cuda and D.is_cuda):
        raise RuntimeError("tc_gemm: Not implemented on the CPU")
    if A.dtype != torch.float32 or B.dtype != torch.float32 or D.dtype != torch.float32:
        raise RuntimeError("tc_gemm: fp32 tensors only")
    with torch.cuda.device(D.device):
        rc = _lib.lib().tc_gemm_tf32_ex(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(A), int(a_mn), lda, _p(B), int(b_mn), ldb,
                                        _p(D), D.stride(0), _p(bias), M, N, R, int(relu), int(accumulate), split_k,
                                        _p(gate), float(gate_scale), float(p_drop), int(seed))
    _lib.check(rc, "tc_gemm_tf32")
    return D


def gemm_bf16(A, a_mn, lda, B, b_mn, ldb, D, M, N, R, bias=None, relu=False, accumulate=False, split_k=1, gate=None, gate_scale=1.0,
              p_drop=0.0, seed=0):
    """The same GEMM with bf16 operands (tcgen05.mma kind::f16, fp32 accumulation): D is bf16, or fp32 when it accumulates (weight
    gradients).  ``bias`` stays fp32.  include/tc_gemm.h: tc_gemm_bf16."""
    if not (A.is_cuda and B.is_cuda and D.is_cuda):
        raise RuntimeError("tc_gemm: Not implemented on the CPU")
    if A.dtype != torch.bfloat16 or B.dtype != torch.bfloat16 or D.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("tc_gemm_bf16: bf16 operands, bf16 or fp32 result")
    with torch.cuda.device(D.device):
        rc = _lib.lib().tc_gemm_bf16(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(A), int(a_mn), lda, _p(B), int(b_mn), ldb,
                                     _p(D), int(D.dtype == torch.float32), D.stride(0), _p(bias), M, N, R, int(relu), int(accumulate), split_k,
                                     _p(gate), float(gate_scale), float(p_drop), int(seed))
    _lib.check(rc, "tc_gemm_bf16")
    return D


def colsum(x2):
    """Column sums of a contiguous fp32 CUDA matrix [rows, C] (bias gradient); ATen's reduction where the kernel's shape limits do not hold."""
    rows, C = x2.shape
    if not (x2.is_cuda and x2.dtype == torch.float32 and x2.is_contiguous() and C % 4 == 0 and C <= 1024 and rows >= 1024):
        return x2.sum(0)
    out = torch.empty(C, dtype=torch.float32, device=x2.device)
    ws = torch.empty(_lib.lib().tc_colsum_workspace_floats(C), dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        rc = _lib.lib().tc_colsum(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(x2), rows, C, C, _p(out), _p(ws))
    _lib.check(rc, "tc_colsum")
    return out


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
            db = colsum(dy2)
        return dx, dw, db, None


class LinearBf16Function(Function):
    """``LinearFunction`` for the bf16 route (autocast / bf16 activations): x and a bf16 copy of the fp32 master weight go through the
    kind::f16 tcgen05 GEMM, y and grad_input are bf16, grad_weight / grad_bias are accumulated in fp32 straight into master-weight
    precision (autocast's own route produces a bf16 weight gradient and casts it)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        if not x.is_cuda:
            raise RuntimeError("tc linear: Not implemented on the CPU")
        K, N = x.shape[-1], weight.shape[0]
        x2 = x.reshape(-1, K).to(torch.bfloat16).contiguous()
        w = weight.to(torch.bfloat16).contiguous()
        M = x2.shape[0]
        y = torch.empty(M, N, dtype=torch.bfloat16, device=x.device)
        gemm_bf16(x2, 0, K, w, 0, K, y, M, N, K, bias=None if bias is None else bias.float().contiguous(), relu=relu)
        ctx.save_for_backward(x2, w, y if relu else None)
        ctx.has_bias, ctx.in_shape, ctx.in_dtype = bias is not None, x.shape, x.dtype
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy2 = dy.reshape(M, N).to(torch.bfloat16)
        dy2 = dy2 * (y > 0) if y is not None else dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.bfloat16, device=dy.device)
            gemm_bf16(dy2, 0, N, w, 1, K, dx, M, K, N)
            dx = dx.reshape(ctx.in_shape).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(N, K, dtype=torch.float32, device=dy.device)
            gemm_bf16(dy2, 1, N, x2, 1, K, dw, N, K, M, accumulate=True, split_k=0)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0, dtype=torch.float32)
        return dx, dw, db, None


class FFNFunction(Function):
    """y = W2 dropout(relu(W1 x + b1)) + b2 -- the feed-forward block of DefAttnLayer / FocusedDecoderLayer (decoder_blocks.py:172-175,
    focused_decoder.py:186-187) as two GEMMs forward and four backward, with everything elementwise in their epilogues: bias + ReLU +
    dropout in the first GEMM (hash mask, never stored), and the ReLU / dropout gradient in the grad_input GEMM of the second
    Linear, gated by the saved activation h (h > 0 exactly where a unit was active and kept)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, p, seed):
        K, Hd, N = x.shape[-1], w1.shape[0], w2.shape[0]
        x2 = x.reshape(-1, K).contiguous()
        w1c, w2c = w1.contiguous(), w2.contiguous()
        M = x2.shape[0]
        h = torch.empty(M, Hd, dtype=torch.float32, device=x.device)
        gemm(x2, 0, K, w1c, 0, K, h, M, Hd, K, bias=b1.contiguous(), relu=True, p_drop=p, seed=seed)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        gemm(h, 0, Hd, w2c, 0, Hd, y, M, N, Hd, bias=b2.contiguous())
        ctx.save_for_backward(x2, w1c, w2c, h)
        ctx.p, ctx.in_shape = float(p), x.shape
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w1, w2, h = ctx.saved_tensors
        M, K = x2.shape
        Hd, N = w1.shape[0], w2.shape[0]
        dy2 = dy.reshape(M, N).contiguous()
        dh = torch.empty(M, Hd, dtype=torch.float32, device=dy.device)
        gemm(dy2, 0, N, w2, 1, Hd, dh, M, Hd, N, gate=h, gate_scale=1.0 / (1.0 - ctx.p))          # (dY W2) * relu'/dropout gate
        dw2 = torch.zeros(N, Hd, dtype=torch.float32, device=dy.device)
        gemm(dy2, 1, N, h, 1, Hd, dw2, N, Hd, M, accumulate=True, split_k=0)
        dw1 = torch.zeros(Hd, K, dtype=torch.float32, device=dy.device)
        gemm(dh, 1, Hd, x2, 1, K, dw1, Hd, K, M, accumulate=True, split_k=0)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
            gemm(dh, 0, Hd, w1, 1, K, dx, M, K, Hd)
            dx = dx.reshape(ctx.in_shape)
        return dx, dw1, colsum(dh), dw2, colsum(dy2), None, None


def ffn(x, linear1, linear2, p, training):
    """``linear2(dropout(relu(linear1(x)), p))``: the fused two-GEMM form when both layers are tcgen05-eligible, else the composition."""
    if (tc_eligible(x, linear1.weight) and tc_eligible(x, linear2.weight) and linear1.bias is not None and linear2.bias is not None
            and x.shape[-1] == linear1.weight.shape[1]):
        p = float(p) if training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
        return FFNFunction.apply(x, linear1.weight, linear1.bias, linear2.weight, linear2.bias, p, seed)
    return linear2(F.dropout(F.relu(linear1(x)), p, training))


def linear(x, weight, bias=None, relu=False):
    if bf16_eligible(x, weight):
        return LinearBf16Function.apply(x, weight, bias, relu)
    return LinearFunction.apply(x, weight, bias, relu)


def bf16_eligible(x, weight):
    """The bf16 route: CUDA, fp32 master weights, and either bf16 activations or an active bf16 autocast region (what the reference's
    trainer runs the model under -- fp16 there, trainer.py:67-69; bf16 is BASELINE configs[2] / [3]).  Feature counts must be multiples
    of 8 (TMA strides are multiples of 16 bytes)."""
    if not (x.is_cuda and weight.is_cuda and weight.shape[0] % 8 == 0 and weight.shape[1] % 8 == 0):
        return False
    if torch.is_autocast_enabled():
        return torch.get_autocast_dtype("cuda") == torch.bfloat16 and x.dtype in (torch.float32, torch.bfloat16)
    return x.dtype == torch.bfloat16


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
        if tc_eligible(x, self.weight) or bf16_eligible(x, self.weight):
            return linear(x, self.weight, self.bias, relu)
        y = F.linear(x, self.weight, self.bias)
        return F.relu(y) if relu else y
