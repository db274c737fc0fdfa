"""Autograd surface of the op -- mirror of transoar/models/ops/functions/ms_deform_attn_func.py:21-38.

``MSDeformAttnFunction.apply(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
im2col_step)`` returns ``[N, Lq, M*C]``; backward yields ``(grad_value, None, None, grad_sampling_loc,
grad_attn_weight, None)`` and is once-differentiable, exactly like the reference.

The reference's ``ms_deform_attn_core_pytorch`` (the ``use_cuda=False`` debug route, func.py:41-65) is deliberately
NOT provided: this package has no non-CUDA compute path (its restatement lives in oracle/ as a test checker).
"""
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                           attention_weights, im2col_step)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        saved = ctx.saved_tensors
        g_value, g_loc, g_attn = MSDA.ms_deform_attn_backward(*saved, grad_output.contiguous(), ctx.im2col_step)
        return g_value, None, None, g_loc, g_attn, None


class MSDeformAttnFusedFunction(Function):
    """The op with its prologue fused: ``apply(value, spatial_shapes, level_start_index, reference_points,
    sampling_offsets, attention_logits)`` equals ``MSDeformAttnFunction.apply(value, shapes, starts, reference_points[:, :, None, :, None, :]
    + sampling_offsets / (W, H, D), softmax(attention_logits over L*P), 64)`` (ms_deform_attn.py:115-136) without materialising the
    locations or the attention weights; gradients flow to value, the raw offsets and the raw logits."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits):
        reference_points, sampling_offsets, attention_logits = (t.contiguous() for t in (reference_points, sampling_offsets, attention_logits))
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits)
        return MSDA.ms_deform_attn_forward_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        g_value, g_off, g_logit = MSDA.ms_deform_attn_backward_fused(*ctx.saved_tensors, grad_output.contiguous())
        return g_value, None, None, None, g_off, g_logit


class MSDeformAttnMergedFunction(Function):
    """``MSDeformAttnFusedFunction`` with the raw offsets and logits in one tensor ``merged`` [N, Lq, 4*M*L*P] -- the output of a single
    Linear layer over the concatenated sampling_offsets / attention_weights weights; the gradient comes back in the same layout."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, merged, n_levels, n_points):
        reference_points, merged = reference_points.contiguous(), merged.contiguous()
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, merged)
        ctx.lp = (n_levels, n_points)
        return MSDA.ms_deform_attn_forward_merged(value, spatial_shapes, level_start_index, reference_points, merged, n_levels, n_points)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        g_value, g_merged = MSDA.ms_deform_attn_backward_merged(*ctx.saved_tensors, grad_output.contiguous(), *ctx.lp)
        return g_value, None, None, None, g_merged, None, None
