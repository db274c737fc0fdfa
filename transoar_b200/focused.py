"""Focused Decoder -- mirrors of transoar/models/necks/focused_decoder.py with the RoI-masked cross-attention running
on the fused sm_100a kernel (include/roi_attn.h) instead of a dense masked [B, heads, Nq, Nkv] score tensor.

``FocusedAttn`` (:192-262), ``FocusedDecoderLayer`` (:82-189), ``FocusedDecoderModel`` (:61-80), ``FocusedDecoder``
(:12-59).  Parameter names equal the reference's (``q_proj/k_proj/v_proj/proj``, ``cross_attn``, ``self_attn``,
``linear1/2``, ``norm1/2/3``), so reference checkpoints load.  Behaviour kept on purpose:
  * the QUERY is projected with ``k_proj`` (:235); ``q_proj`` is a dead parameter that never gets a gradient;
  * dropout only on the output projection (attn_drop = 0 in every reference config).
Differences, all documented: masks are represented by their boxes (the reference only ever builds box masks,
generate_attn_masks :138-159) and derived from the feature-map shape instead of the hard-coded 160x160x256 / 256x256x128
table (:99-117); the dense attention-weight tensor (1.77 GB per sample at VISCERAL) is only materialised
when ``materialize_weights`` is set (scripts/test.py:78-80 reads it for visualisation; the decoder loop discards it, :72).
"""
import copy
import ctypes

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .fused_ln import add_dropout_layer_norm
from .linear import TCLinear, ffn

_QPG = 32     # query rows per CTA in the kernel (roiattn::TQ)


def boxes_from_bbox_props(bbox_props, num_queries, input_shape, restrict_attn=True, padding=0):
    """int32 [num_queries, 6] voxel boxes (x1,y1,z1,x2,y2,z2), restating generate_attn_masks (:138-159)."""
    shape = torch.as_tensor(list(input_shape), dtype=torch.float32)
    areas = torch.stack([torch.as_tensor(p["attn_area"], dtype=torch.float32) for p in bbox_props.values()])
    per_organ = num_queries // len(areas)
    vol = torch.repeat_interleave(areas, per_organ, dim=0)                              # :144
    vol = (vol * shape.repeat(2) - padding)
    vol = torch.maximum(vol, torch.zeros(6)).minimum(shape.repeat(2))                    # clamp(0, shape) :147
    vol[:, :3] = torch.floor(vol[:, :3])
    vol[:, 3:] = torch.ceil(vol[:, 3:])
    boxes = vol.int()
    if not restrict_attn:                                                                # :159 -> nothing masked
        boxes = torch.tensor([0, 0, 0, *map(int, input_shape)], dtype=torch.int32).repeat(num_queries, 1)
    return boxes


def boxes_from_mask(attn_mask, grid_shape):
    """Recover the boxes from a reference-style boolean mask [Nq, X*Y*Z] (True = masked).  Raises if a row is not a box."""
    Nq = attn_mask.shape[0]
    X, Y, Z = (int(s) for s in grid_shape)
    free = ~attn_mask.reshape(Nq, X, Y, Z).cpu()
    boxes = torch.zeros(Nq, 6, dtype=torch.int32)
    for q in range(Nq):
        idx = free[q].nonzero()
        if idx.numel() == 0:
            continue                                                                     # empty box: all six stay 0
        lo, hi = idx.min(0).values, idx.max(0).values + 1
        if int((hi - lo).prod()) != idx.shape[0]:
            raise ValueError(f"attn_mask row {q} is not an axis-aligned box; the fused FocusedAttn only supports box masks")
        boxes[q] = torch.cat((lo, hi)).int()
    return boxes


def groups_from_boxes(boxes):
    """int32 [G, 8] = {q0, nq, box}: runs of consecutive queries with identical boxes, split into chunks of <= 32."""
    boxes = boxes.cpu().int()
    out, q0 = [], 0
    n = boxes.shape[0]
    while q0 < n:
        q1 = q0 + 1
        while q1 < n and q1 - q0 < _QPG and torch.equal(boxes[q1], boxes[q0]):
            q1 += 1
        out.append([q0, q1 - q0, *boxes[q0].tolist()])
        q0 = q1
    return torch.tensor(out, dtype=torch.int32)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class RoIAttentionFunction(Function):
    """out[B,Nq,H*HD] = softmax_{tokens in box(q)}(q k^T) v  -- the core of focused_decoder.py:238-254."""

    @staticmethod
    def forward(ctx, q, k, v, groups, grid_yz, tf32=None):
        if not (q.is_cuda and k.is_cuda and v.is_cuda and groups.is_cuda):
            raise RuntimeError("RoI attention: Not implemented on the CPU")
        # TF32 tensor-core kernels exactly where torch itself would use TF32 for the reference's q @ k^T / attn @ v (focused_decoder.py:238,254)
        ctx.tf32 = torch.backends.cuda.matmul.allow_tf32 if tf32 is None else bool(tf32)
        fwd = _lib.lib().roi_attn_forward_tf32 if ctx.tf32 else _lib.lib().roi_attn_forward
        q, k, v = q.float().contiguous(), k.float().contiguous(), v.float().contiguous()
        B, Nq, H, HD = q.shape
        Nkv = k.shape[1]
        out = torch.empty(B, Nq, H * HD, dtype=torch.float32, device=q.device)
        lse = torch.empty(B, H, Nq, dtype=torch.float32, device=q.device)
        ws_n = _lib.lib().roi_attn_workspace_floats(groups.shape[0], B, Nq, H, HD)
        ws = torch.empty(max(ws_n, 1), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            rc = fwd(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(q), _p(k), _p(v), _p(groups),
                                             groups.shape[0], B, Nq, Nkv, H, HD, grid_yz[0], grid_yz[1], _p(out), _p(lse),
                                             _p(ws), ws_n)
        _lib.check(rc, "roi_attn_forward")
        ctx.save_for_backward(q, k, v, groups, out, lse)
        ctx.grid_yz = grid_yz
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        q, k, v, groups, out, lse = ctx.saved_tensors
        B, Nq, H, HD = q.shape
        Nkv = k.shape[1]
        dout = dout.float().contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        with torch.cuda.device(q.device):
            bwd = _lib.lib().roi_attn_backward_tf32 if ctx.tf32 else _lib.lib().roi_attn_backward
            rc = bwd(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _p(q), _p(k), _p(v), _p(groups),
                                              groups.shape[0], B, Nq, Nkv, H, HD, ctx.grid_yz[0], ctx.grid_yz[1], _p(out), _p(dout),
                                              _p(lse), _p(dq), _p(dk), _p(dv))
        _lib.check(rc, "roi_attn_backward")
        return dq, dk, dv, None, None, None


class FocusedAttn(nn.Module):
    """focused_decoder.py:192-262.  ``attn_mask``: the reference's boolean mask [Nq, Nkv] (needs ``grid_shape``) or int32 boxes [Nq, 6]."""

    def __init__(self, dim, num_heads, attn_mask, qkv_bias=None, qk_scale=None, attn_drop=0, proj_drop=0, use_pos_bias=False,
                 return_weights=True, grid_shape=None):
        super().__init__()
        if use_pos_bias:
            raise NotImplementedError("use_pos_bias is never enabled by the reference (focused_decoder.py:121) and is not fused")
        if attn_drop:
            raise NotImplementedError("attention dropout is 0 in every reference config and is not fused")
        if grid_shape is None:
            raise ValueError("FocusedAttn needs grid_shape=(X, Y, Z) of the key/value feature map")
        self.dim, self.num_heads, self.ret_weights = dim, num_heads, return_weights
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.q_proj = TCLinear(dim, dim, bias=bool(qkv_bias))       # dead parameter, kept for checkpoint compatibility (D10)
        self.k_proj = TCLinear(dim, dim, bias=bool(qkv_bias))
        self.v_proj = TCLinear(dim, dim, bias=bool(qkv_bias))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = TCLinear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.pos_bias = None
        self.grid_shape = tuple(int(s) for s in grid_shape)
        boxes = attn_mask if attn_mask.dtype == torch.int32 and attn_mask.shape[-1] == 6 else boxes_from_mask(attn_mask, grid_shape)
        self.register_buffer("boxes", boxes.int(), persistent=False)
        self.register_buffer("groups", groups_from_boxes(boxes), persistent=False)
        self.materialize_weights = False

    def dense_mask(self):
        """The reference's boolean mask [Nq, Nkv] (True = masked), rebuilt from the boxes on demand."""
        X, Y, Z = self.grid_shape
        m = torch.ones(self.boxes.shape[0], X, Y, Z, dtype=torch.bool, device=self.boxes.device)
        for qi, (x1, y1, z1, x2, y2, z2) in enumerate(self.boxes.tolist()):
            m[qi, x1:x2, y1:y2, z1:z2] = False
        return m.flatten(1)

    def forward(self, q, k, v, mask=None):
        B, Nkv, C = k.shape
        Nq = q.shape[1]
        H = self.num_heads
        if Nkv != self.grid_shape[0] * self.grid_shape[1] * self.grid_shape[2]:
            # the reference fails here too: its [Nq, X*Y*Z] mask does not broadcast against [B, H, Nq, Nkv] (focused_decoder.py:243-245)
            raise RuntimeError(f"FocusedAttn: {Nkv} key/value tokens, but the RoI boxes were built for a {self.grid_shape} feature map "
                               f"({self.grid_shape[0] * self.grid_shape[1] * self.grid_shape[2]} tokens); pass the feature-map shape "
                               "(config['neck_input_shape'] / FocusedDecoderLayer(input_shape=...))")
        kp = self.k_proj(k).reshape(B, Nkv, H, C // H)
        vp = self.v_proj(v).reshape(B, Nkv, H, C // H)
        qp = self.k_proj(q).reshape(B, Nq, H, C // H) * self.scale              # :235-236 (k_proj on the query: reference quirk)
        x = RoIAttentionFunction.apply(qp, kp, vp, self.groups, self.grid_shape[1:])
        weights = None
        if self.ret_weights and self.materialize_weights:
            with torch.no_grad():
                s = torch.matmul(qp.permute(0, 2, 1, 3), kp.permute(0, 2, 3, 1))
                s = s.masked_fill(self.dense_mask()[None, None], float("-inf"))
                weights = s.softmax(-1)
        x = self.proj_drop(self.proj(x.to(q.dtype)))
        return (x, weights) if self.ret_weights else x


def _activation(name):
    try:
        return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]
    except KeyError:
        raise RuntimeError(f"activation should be relu/gelu, not {name}.")


class FocusedDecoderLayer(nn.Module):
    """focused_decoder.py:82-189: MHA self-attention over the queries -> RoI cross-attention -> FFN (post-norm)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_heads=8, config=None, bbox_props=None,
                 input_shape=None):
        super().__init__()
        self.config, self.bbox_props = config, bbox_props
        self.num_queries_per_organ = int(config["num_queries"] / config["num_organs"])
        assert self.num_queries_per_organ in [1, 7, 27, 54]                                   # :96
        if input_shape is None:                                                             # the reference's table (:99-117)
            table = {20: [160, 160, 256]}.get(config["num_organs"], [256, 256, 128])
            input_shape = [s // 2 ** int(config["input_levels"][1]) for s in table]
        self.input_shape = tuple(int(s) for s in input_shape)
        boxes = boxes_from_bbox_props(bbox_props, config["num_queries"], self.input_shape, config["restrict_attn"])
        self.cross_attn = FocusedAttn(d_model, n_heads, boxes, proj_drop=0.1, grid_shape=self.input_shape)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = TCLinear(d_model, d_ffn)
        self.activation = _activation(activation)
        self._fuse_relu = activation == "relu"
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = TCLinear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt, query_pos, src_pos, src, src_k=None):
        """``src_k``: ``src + src_pos`` computed by the caller (it is the same tensor for every layer of the decoder)."""
        qk = tgt if query_pos is None else tgt + query_pos
        sa = self.self_attn(qk.transpose(0, 1), qk.transpose(0, 1), tgt.transpose(0, 1))[0].transpose(0, 1)
        tgt = add_dropout_layer_norm(tgt, sa, self.norm2, self.dropout2.p, self.training)
        q = tgt if query_pos is None else tgt + query_pos
        k = src_k if src_k is not None else (src if src_pos is None else src + src_pos)
        ca, weights = self.cross_attn(q, k, src)
        tgt = add_dropout_layer_norm(tgt, ca, self.norm1, self.dropout1.p, self.training)
        if self._fuse_relu:
            out = ffn(tgt, self.linear1, self.linear2, self.dropout3.p, self.training)
        else:
            out = self.linear2(self.dropout3(self.activation(self.linear1(tgt))))
        return add_dropout_layer_norm(tgt, out, self.norm3, self.dropout4.p, self.training), weights


class FocusedDecoderModel(nn.Module):
    """focused_decoder.py:61-80."""

    def __init__(self, decoder_layer, num_layers, return_intermediate=False):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)])
        self.num_layers, self.return_intermediate = num_layers, return_intermediate

    def forward(self, tgt, src, src_pos, query_pos=None):
        output, inter = tgt, []
        src_k = src if src_pos is None else src + src_pos               # focused_decoder.py:173 forms this sum in every layer
        for layer in self.layers:
            output, _ = layer(output, query_pos, src_pos, src, src_k=src_k)
            if self.return_intermediate:
                inter.append(output)
        return torch.stack(inter) if self.return_intermediate else output


class FocusedDecoder(nn.Module):
    """focused_decoder.py:12-59."""

    def __init__(self, d_model=256, nhead=8, num_decoder_layers=6, dim_feedforward=1024, dropout=0.1, activation="relu",
                 return_intermediate_dec=False, bbox_props=None, config=None, input_shape=None):
        super().__init__()
        self.bbox_props, self.config, self.d_model, self.nhead = bbox_props, config, d_model, nhead
        layer = FocusedDecoderLayer(d_model, dim_feedforward, dropout, activation, nhead, config, bbox_props, input_shape)
        self.decoder = FocusedDecoderModel(layer, num_decoder_layers, return_intermediate_dec)
        for p in self.parameters():                                                        # :39-42
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, src, query_embed, pos):
        assert query_embed is not None
        src = src.flatten(2).transpose(1, 2).contiguous()          # a view + one plain copy when the backbone is channels-last
        pos = pos[:1].flatten(2).transpose(1, 2)                    # the sine encoding is batch-independent: one copy, broadcast in src + pos
        bs, _, c = src.shape
        query_pos, tgt = torch.split(query_embed, c, dim=1)
        return self.decoder(tgt.unsqueeze(0).expand(bs, -1, -1), src, pos, query_pos.unsqueeze(0).expand(bs, -1, -1))
