"""``MSDeformAttn`` -- mirror of transoar/models/ops/modules/ms_deform_attn.py:30-141.

Same constructor, parameter names (``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` -- so
reference checkpoints load unchanged), initialisation and forward semantics; the sampling itself always runs on the
sm_100a kernels behind ``MSDeformAttnFunction``.  ``use_cuda=False`` selected the pure-PyTorch debug route in the
reference (ms_deform_attn.py:137-138); this package has no such route and raises instead of silently falling back.
"""
import warnings
import weakref

import torch
import torch.nn.functional as F
from torch import nn

from ... import MultiScaleDeformableAttention as MSDA
from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnMergedFunction
from ...linear import TCLinear, linear, tc_eligible


def _power_of_two(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return n != 0 and (n & (n - 1)) == 0


def _direction_table(n_heads):
    """Unit steps the offset bias is initialised with (ms_deform_attn.py:63-73): 6 face or 26 neighbour directions."""
    cube = torch.cartesian_prod(*([torch.tensor([-1, 0, 1])] * 3)).float()
    l1 = cube.abs().sum(1)
    if n_heads == 26:
        return cube[l1 > 0]
    if n_heads == 6:
        return cube[l1 == 1]
    raise ValueError("Only nheads of value 26 or 6 are supported.")           # ms_deform_attn.py:72-73


_COVER_CHECKED = {}


def _assert_levels_cover(spatial_shapes, S):
    """The reference's ``assert (shapes.prod(1)).sum() == S`` (ms_deform_attn.py:107) reads a device tensor on the host in every forward
    of every layer.  Here the answer is remembered per shapes tensor (identity + version counter), so a model that reuses its shapes
    tensor -- as ``DecoderDefAttnBlock`` does -- synchronises once, and the forward can be captured into a CUDA graph afterwards."""
    hit = _COVER_CHECKED.get(id(spatial_shapes))
    if hit is not None and hit[0]() is spatial_shapes and hit[1] == spatial_shapes._version and hit[2] == S:
        return
    if spatial_shapes.is_cuda and torch.cuda.is_current_stream_capturing():
        raise RuntimeError("MSDeformAttn: run one eager forward with this spatial_shapes tensor before capturing a CUDA graph")
    assert int(spatial_shapes.prod(1).sum()) == S
    if len(_COVER_CHECKED) > 256:
        _COVER_CHECKED.clear()
    _COVER_CHECKED[id(spatial_shapes)] = (weakref.ref(spatial_shapes), spatial_shapes._version, S)


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, use_cuda=True):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError('d_model must be divisible by n_heads, but got {} and {}'.format(d_model, n_heads))
        if not _power_of_two(d_model // n_heads):
            warnings.warn("d_model // n_heads is not a power of two: the op falls back from the vector kernels to the "
                          "generic (slower) CUDA kernels.")
        self.im2col_step = 64                                                  # ms_deform_attn.py:48
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.use_cuda = use_cuda
        self.fuse_prologue = True                                              # False: always the reference's three-step prologue + plain op
        # nn.Linear subclasses (same parameter names): TF32 tcgen05 GEMMs when TF32 is the requested matmul precision
        self.sampling_offsets = TCLinear(d_model, n_heads * n_levels * n_points * 3)
        self.attention_weights = TCLinear(d_model, n_heads * n_levels * n_points)
        self.value_proj = TCLinear(d_model, d_model)
        self.output_proj = TCLinear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        """ms_deform_attn.py:63-91: zero offset/attention weights, directional offset bias scaled by (p+1), xavier projections."""
        nn.init.zeros_(self.sampling_offsets.weight)
        steps = torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
        bias = _direction_table(self.n_heads).view(self.n_heads, 1, 1, 3) * steps          # [M,1,P,3]
        bias = bias.expand(self.n_heads, self.n_levels, self.n_points, 3).reshape(-1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(bias.clone())            # stays trainable, as in the reference (:80-82)
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)
        for proj in (self.value_proj, self.output_proj):
            nn.init.xavier_uniform_(proj.weight)
            nn.init.zeros_(proj.bias)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """query [N,Lq,C]; reference_points [N|1,Lq,L,3] in (x,y,z); input_flatten [N,S,C]; shapes [L,3]=(D,H,W) -> [N,Lq,C]."""
        N, Lq, _ = query.shape
        _, S, _ = input_flatten.shape
        _assert_levels_cover(input_spatial_shapes, S)                          # ms_deform_attn.py:107
        M, L, P = self.n_heads, self.n_levels, self.n_points

        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, S, M, self.d_model // M)
        if reference_points.shape[-1] != 3:
            raise ValueError('Last dim of reference_points must be 3, but get {} instead.'.format(reference_points.shape[-1]))
        if (self.use_cuda and self.fuse_prologue and tc_eligible(query, self.sampling_offsets.weight)
                and MSDA.fused_supported(value, reference_points, L, P)):
            # both small projections as ONE GEMM over the concatenated weights (N = 4*M*L*P = 384 columns); its output feeds the op
            # directly and the op's backward returns one gradient tensor for one pair of gradient GEMMs
            w_cat = torch.cat((self.sampling_offsets.weight, self.attention_weights.weight), 0)
            b_cat = torch.cat((self.sampling_offsets.bias, self.attention_weights.bias), 0)
            merged = linear(query, w_cat, b_cat)
            sampled = MSDeformAttnMergedFunction.apply(value, input_spatial_shapes, input_level_start_index, reference_points, merged, L, P)
            return self.output_proj(sampled)
        offsets = self.sampling_offsets(query).view(N, Lq, M, L, P, 3)
        logits = self.attention_weights(query).view(N, Lq, M, L, P)
        if self.use_cuda and self.fuse_prologue and MSDA.fused_supported(value, reference_points, L, P):
            # softmax over the unit's L*P logits and ref + offset / (W, H, D) happen inside the kernels (include/msda3d.h, *_fused)
            sampled = MSDeformAttnFusedFunction.apply(value, input_spatial_shapes, input_level_start_index, reference_points, offsets, logits)
            return self.output_proj(sampled)
        weights = F.softmax(logits.view(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
        normalizer = input_spatial_shapes.flip(-1)                             # (W,H,D): x,y,z order, ms_deform_attn.py:123-126
        locations = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
        if not self.use_cuda:
            raise RuntimeError("transoar_b200.MSDeformAttn only implements the use_cuda=True route; the reference's "
                               "pure-PyTorch debug path (ms_deform_attn_core_pytorch) is not part of this package.")
        sampled = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                             locations.contiguous(), weights, self.im2col_step)
        return self.output_proj(sampled)
