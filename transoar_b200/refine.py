"""Deformable FPN refinement -- mirrors of transoar/models/backbones/decoder_blocks.py:
``DecoderDefAttnBlock`` (:12-97), ``DefAttnTransformer`` (:100-141), ``DefAttnLayer`` (:143-177).

Attribute names equal the reference's, so a reference ``state_dict`` loads unchanged
(``level_embed``, ``refine_def_attn.layers.<i>.self_attn.{sampling_offsets,attention_weights,value_proj,output_proj}``,
``linear1/2``, ``norm1/2``).  The sampling runs on the sm_100a kernels through ``MSDeformAttn``; projections, LayerNorm and
the FFN stay ATen (cuBLAS) calls, as in the reference.  Reference points depend only on the level shapes and are cached."""
import copy

import torch
import torch.nn.functional as F
from torch import nn

from .fused_ln import add_dropout_layer_norm
from .linear import TCLinear, ffn
from .ops.modules import MSDeformAttn


def _activation(name):
    try:
        return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]
    except KeyError:
        raise RuntimeError(f"activation should be relu/gelu/glu, not {name}.")


class DefAttnLayer(nn.Module):
    """decoder_blocks.py:143-177: deformable self-attention + FFN, post-norm."""

    def __init__(self, d_model, d_ffn, dropout, activation, n_levels, n_heads, n_points, use_cuda):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, use_cuda)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = TCLinear(d_model, d_ffn)
        self.activation = _activation(activation)
        self._fuse_relu = activation == "relu"                                  # ReLU runs in the GEMM epilogue
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = TCLinear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        query = src if pos is None else src + pos
        attn = self.self_attn(query, reference_points, src, spatial_shapes, level_start_index, padding_mask)
        src = add_dropout_layer_norm(src, attn, self.norm1, self.dropout1.p, self.training)       # norm1(src + dropout1(attn)), one kernel
        if self._fuse_relu:      # linear2(dropout2(relu(linear1(src)))): bias + ReLU + dropout live in the first GEMM's epilogue
            out = ffn(src, self.linear1, self.linear2, self.dropout2.p, self.training)
        else:
            out = self.linear2(self.dropout2(self.activation(self.linear1(src))))
        return add_dropout_layer_norm(src, out, self.norm2, self.dropout3.p, self.training)       # norm2(src + dropout3(ffn))


class DefAttnTransformer(nn.Module):
    """decoder_blocks.py:100-141."""

    def __init__(self, layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self._ref_cache = {}

    @staticmethod
    def get_reference_points(spatial_shapes, device):
        """Voxel centres of every level, normalised, (x,y,z) order, repeated over levels: [1, S, L, 3] (:107-131)."""
        pts = []
        for D_, H_, W_ in spatial_shapes.tolist():
            z = torch.linspace(0.5, D_ - 0.5, D_, dtype=torch.float32, device=device) / D_
            y = torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device) / H_
            x = torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device) / W_
            zz, yy, xx = torch.meshgrid(z, y, x, indexing="ij")
            pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)), -1))
        ref = torch.cat(pts, 0)[None, :, None, :]
        return ref.expand(1, -1, spatial_shapes.size(0), 3).contiguous()

    def forward(self, src, spatial_shapes, level_start_index, pos=None, shapes_key=None):
        key = (shapes_key if shapes_key is not None else tuple(map(tuple, spatial_shapes.tolist())), src.device)
        ref = self._ref_cache.get(key)
        if ref is None:
            ref = self._ref_cache[key] = self.get_reference_points(spatial_shapes, src.device)
        out = src
        for layer in self.layers:
            out = layer(out, pos, ref, spatial_shapes, level_start_index)
        return out


class DecoderDefAttnBlock(nn.Module):
    """decoder_blocks.py:12-97: flatten + concat the FPN levels, level embedding, N deformable layers, split back."""

    def __init__(self, d_model, nhead, num_layers, dim_feedforward, dropout, feature_levels, n_points, use_cuda=True,
                 activation="relu"):
        super().__init__()
        self.d_model, self.nhead, self.feature_levels = d_model, nhead, feature_levels
        n_levels = len(feature_levels)
        layer = DefAttnLayer(d_model, dim_feedforward, dropout, activation, n_levels, nhead, n_points, use_cuda)
        self.refine_def_attn = DefAttnTransformer(layer, num_layers)
        self.level_embed = nn.Parameter(torch.Tensor(n_levels, d_model))
        self._level_cache = {}
        self._reset_parameters()

    def _reset_parameters(self):
        """:40-48 -- xavier on every >1-D parameter, THEN each MSDeformAttn re-initialises itself, then the level embedding."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    def forward(self, fmaps, pos_embeds):
        shapes_py = tuple(tuple(f.shape[2:]) for f in fmaps)
        dev = fmaps[0].device
        cached = self._level_cache.get((shapes_py, dev))
        if cached is None:    # the reference rebuilds these two device tensors (and syncs on .tolist()) every forward (:83-95)
            ss = torch.as_tensor(shapes_py, dtype=torch.long, device=dev)
            starts = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
            cached = self._level_cache[(shapes_py, dev)] = (ss, starts)
        spatial_shapes, level_start_index = cached
        src = torch.cat([f.flatten(2).transpose(1, 2) for f in fmaps], 1)                                  # [N, S, C]
        # the sine encodings are the same for every sample of the batch: one copy [1, S, C], broadcast in `src + pos` (the reference carries
        # N identical copies through the level-embedding addition and the concatenation, decoder_blocks.py:69-95)
        pos = torch.cat([p[:1].flatten(2).transpose(1, 2) + self.level_embed[l].view(1, 1, -1)
                         for l, p in enumerate(pos_embeds)], 1)
        memory = self.refine_def_attn(src, spatial_shapes, level_start_index, pos, shapes_key=shapes_py)
        bs, c = fmaps[0].shape[:2]
        sizes = [d * h * w for d, h, w in shapes_py]
        return [m.transpose(-1, -2).reshape(bs, c, *shp) for m, shp in zip(torch.split(memory, sizes, dim=1), shapes_py)]
