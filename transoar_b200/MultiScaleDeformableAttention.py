"""Stand-in for the reference's compiled extension module ``MultiScaleDeformableAttention``
(transoar/models/ops/setup.py:53, exported by transoar/models/ops/src/vision.cpp:13-16).

Same two entry points, same argument order, same error behaviour as the reference's host wrappers
(transoar/models/ops/src/ms_deform_attn.h:20-61, cuda/ms_deform_attn_cuda.cu:20-80,83-154):
contiguity / device checks raise RuntimeError, CPU tensors raise "Not implemented on the CPU",
``batch % min(batch, im2col_step) == 0`` is enforced although one launch now covers the whole batch.

Extension over the reference: ``value`` may be bf16/fp16 with fp32 ``sampling_loc`` / ``attn_weight``
-- the case the reference's own trainer produces under autocast and its op rejects.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


# Optional per-launch timing (bench.py): when a list is installed, each forward / backward appends
# (kind, start_event, stop_event) recorded on the launching stream.  None (the default) costs nothing.
_EVENT_LOG = None


def set_event_log(log):
    """Install (or with None remove) the list that receives ("fwd"|"bwd", start_event, stop_event, dims) per launch;
    dims = (N, S, M, C, L, Lq, P, bytes per value element)."""
    global _EVENT_LOG
    _EVENT_LOG = log


class _timed:
    def __init__(self, kind, dims=None):
        self.kind, self.dims = kind, dims

    def __enter__(self):
        if _EVENT_LOG is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _EVENT_LOG is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _EVENT_LOG.append((self.kind, self.e0, e1, self.dims))


def _p(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _check(named, im2col_step):
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")           # ms_deform_attn_cuda.cu:28-32
    value = named[0][1]
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")                          # ms_deform_attn.h:38,60
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")                   # ms_deform_attn_cuda.cu:34-38
        if t.device != value.device:
            raise RuntimeError(f"{name} must live on {value.device}")
    if value.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn: unsupported dtype {value.dtype}")    # AT_DISPATCH_FLOATING_TYPES, .cu:64
    batch = value.size(0)
    step = min(batch, int(im2col_step))
    if step <= 0 or batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")     # ms_deform_attn_cuda.cu:52
    return _DTYPES[value.dtype]


def _aux_dtype(value):
    return torch.float64 if value.dtype == torch.float64 else torch.float32


def _dims(value, spatial_shapes, sampling_loc):
    N, S, M, C = value.shape
    L = spatial_shapes.size(0)
    Lq, P = sampling_loc.size(1), sampling_loc.size(4)
    if sampling_loc.shape != (N, Lq, M, L, P, 3):
        raise RuntimeError(f"sampling_loc has shape {tuple(sampling_loc.shape)}, expected {(N, Lq, M, L, P, 3)}")
    return N, S, M, C, L, Lq, P


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """-> output [N, Lq, M*C]   (vision.cpp:14, ms_deform_attn_cuda_forward)."""
    dt = _check([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                 ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)], im2col_step)
    aux = _aux_dtype(value)
    if sampling_loc.dtype != aux or attn_weight.dtype != aux:
        sampling_loc, attn_weight = sampling_loc.to(aux), attn_weight.to(aux)
    N, S, M, C, L, Lq, P = _dims(value, spatial_shapes, sampling_loc)
    out = torch.empty((N, Lq, M * C), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _timed("fwd", (N, S, M, C, L, Lq, P, value.element_size())):
        rc = _lib.lib().msda3d_forward(
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), dt, _p(value), _p(spatial_shapes),
            _p(level_start_index), _p(sampling_loc), _p(attn_weight), N, S, M, C, L, Lq, P, _p(out))
    _lib.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight]   (vision.cpp:15, ms_deform_attn_cuda_backward)."""
    dt = _check([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                 ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)], im2col_step)
    aux = _aux_dtype(value)
    loc_dtype, aw_dtype = sampling_loc.dtype, attn_weight.dtype
    if loc_dtype != aux or aw_dtype != aux:
        sampling_loc, attn_weight = sampling_loc.to(aux), attn_weight.to(aux)
    if grad_output.dtype != value.dtype:
        grad_output = grad_output.to(value.dtype)
    N, S, M, C, L, Lq, P = _dims(value, spatial_shapes, sampling_loc)
    grad_value = torch.empty(value.shape, dtype=aux, device=value.device)       # zero-filled by the library
    grad_loc = torch.empty_like(sampling_loc)
    grad_aw = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device), _timed("bwd", (N, S, M, C, L, Lq, P, value.element_size())):
        rc = _lib.lib().msda3d_backward(
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), dt, _p(grad_output), _p(value), _p(spatial_shapes),
            _p(level_start_index), _p(sampling_loc), _p(attn_weight), N, S, M, C, L, Lq, P,
            _p(grad_value), _p(grad_loc), _p(grad_aw))
    _lib.check(rc, "ms_deform_attn_backward")
    return [grad_value.to(value.dtype), grad_loc.to(loc_dtype), grad_aw.to(aw_dtype)]


# ---------------------------------------------------------------------------------------------------------------
# Fused prologue (include/msda3d.h, msda3d_*_fused): softmax over the unit's L*P logits and ref + offset / (W, H, D) inside the op
# ---------------------------------------------------------------------------------------------------------------
def fused_supported(value, reference_points, n_levels, n_points):
    """fp32 CUDA tensors, vector kernels, L * P <= C / 4 (C = 64, L * P <= 16: the reference's configuration), 3-D reference points."""
    if not (value.is_cuda and value.dtype == torch.float32 and reference_points.dtype == torch.float32 and reference_points.shape[-1] == 3
            and not torch.is_autocast_enabled()):
        return False
    if reference_points.requires_grad and torch.is_grad_enabled():
        # learnable reference points (Deformable-DETR decoders): the fused functions return no gradient for them, the unfused route
        # propagates it through sampling_locations as the reference does
        return False
    return bool(_lib.lib().msda3d_fused_supported(value.shape[3], n_levels, n_points)) and reference_points.shape[0] in (1, value.shape[0])


def ms_deform_attn_forward_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits):
    """-> output [N, Lq, M*C] from raw offsets [N,Lq,M,L,P,3], logits [N,Lq,M,L,P] and reference points [1|N, Lq, L, 3]."""
    _check([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
            ("reference_points", reference_points), ("sampling_offsets", sampling_offsets), ("attn_logits", attn_logits)], 1)
    N, S, M, C, L, Lq, P = _dims(value, spatial_shapes, sampling_offsets)
    out = torch.empty((N, Lq, M * C), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _timed("fwd", (N, S, M, C, L, Lq, P, 4)):
        rc = _lib.lib().msda3d_forward_fused(
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(value), _p(spatial_shapes), _p(level_start_index),
            _p(reference_points), reference_points.shape[0], _p(sampling_offsets), _p(attn_logits), N, S, M, C, L, Lq, P, _p(out))
    _lib.check(rc, "ms_deform_attn_forward_fused")
    return out


def ms_deform_attn_backward_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits, grad_output):
    """-> [grad_value, grad_sampling_offsets, grad_attn_logits]."""
    _check([("value", value), ("reference_points", reference_points), ("sampling_offsets", sampling_offsets), ("attn_logits", attn_logits),
            ("grad_output", grad_output)], 1)
    N, S, M, C, L, Lq, P = _dims(value, spatial_shapes, sampling_offsets)
    grad_value = torch.empty_like(value)                                         # zero-filled by the library
    grad_off, grad_logit = torch.empty_like(sampling_offsets), torch.empty_like(attn_logits)
    with torch.cuda.device(value.device), _timed("bwd", (N, S, M, C, L, Lq, P, 4)):
        rc = _lib.lib().msda3d_backward_fused(
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(grad_output), _p(value), _p(spatial_shapes), _p(level_start_index),
            _p(reference_points), reference_points.shape[0], _p(sampling_offsets), _p(attn_logits), N, S, M, C, L, Lq, P,
            _p(grad_value), _p(grad_off), _p(grad_logit))
    _lib.check(rc, "ms_deform_attn_backward_fused")
    return [grad_value, grad_off, grad_logit]


def ms_deform_attn_forward_merged(value, spatial_shapes, level_start_index, reference_points, merged, n_levels, n_points):
    """Fused prologue with offsets and logits in one tensor ``merged`` [N, Lq, ld]: columns [0, 3*M*L*P) offsets, then M*L*P logits."""
    _check([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
            ("reference_points", reference_points), ("merged", merged)], 1)
    N, S, M, C = value.shape
    Lq, ld = merged.shape[1], merged.shape[2]
    out = torch.empty((N, Lq, M * C), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _timed("fwd", (N, S, M, C, n_levels, Lq, n_points, 4)):
        rc = _lib.lib().msda3d_forward_fused_ld(
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(value), _p(spatial_shapes), _p(level_start_index),
            _p(reference_points), reference_points.shape[0], _p(merged), None, ld, N, S, M, C, n_levels, Lq, n_points, _p(out))
    _lib.check(rc, "ms_deform_attn_forward_merged")
    return out


def ms_deform_attn_backward_merged(value, spatial_shapes, level_start_index, reference_points, merged, grad_output, n_levels, n_points):
    """-> [grad_value, grad_merged]."""
    _check([("value", value), ("reference_points", reference_points), ("merged", merged), ("grad_output", grad_output)], 1)
    N, S, M, C = value.shape
    Lq, ld = merged.shape[1], merged.shape[2]
    grad_value = torch.empty_like(value)
    grad_merged = torch.empty_like(merged) if ld == 4 * M * n_levels * n_points else torch.zeros_like(merged)
    with torch.cuda.device(value.device), _timed("bwd", (N, S, M, C, n_levels, Lq, n_points, 4)):
        rc = _lib.lib().msda3d_backward_fused_ld(
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(grad_output), _p(value), _p(spatial_shapes), _p(level_start_index),
            _p(reference_points), reference_points.shape[0], _p(merged), None, ld, N, S, M, C, n_levels, Lq, n_points,
            _p(grad_value), _p(grad_merged), None)
    _lib.check(rc, "ms_deform_attn_backward_merged")
    return [grad_value, grad_merged]
